"""Host-side mirror of the CompressAI operator interface the reference builds on
(``compressai.layers.GDN``, ``compressai.entropy_models.{EntropyBottleneck,GaussianConditional}``,
``compressai.layers.ResidualBlock*``, ``compressai.models.MeanScaleHyperprior``; imported by the reference at
``LHBDC/model/layers.py:6-17``): same class names, attribute names, constructor arguments and state-dict keys
(so reference checkpoints load unchanged), but every hot operator calls the sm_100a kernels in libb200vc.so.

Inference only (the reference's eval path: ``LHBDC/test/testing.py``, ``encode_B.py``, ``decode_B.py``);
training-mode noise quantisation raises.  There is no CPU path: parameters may be *constructed* on the CPU (to
build / load / move a model) but ``forward`` requires CUDA tensors.

The ``*_forward`` functions are written against duck-typed attributes so that ``b200vc.patch`` can bind them to
the modules of a model built from the *reference's own* classes (CompressAI installed) without touching its
module tree.
"""
import math

import torch
import torch.nn as nn

from . import ops

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64


def get_scale_table(lo=SCALES_MIN, hi=SCALES_MAX, levels=SCALES_LEVELS):
    """exp(linspace(ln 0.11, ln 256, 64)) -- restated in-repo by the reference at ICIP2023/src/model/elic.py:21-27."""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


def _eval_only(mod, training):
    if training if training is not None else mod.training:
        raise NotImplementedError(
            "b200vc implements the reference's inference path only (quantize mode 'dequantize'); "
            "call .eval() -- training-mode noise quantisation is out of scope")


def _cached(mod, name, key, build):
    slot = mod.__dict__.get(name)
    if slot is None or slot[0] != key:
        slot = (key, build())
        mod.__dict__[name] = slot
    return slot[1]


def _pkey(*tensors):
    return tuple((t.data_ptr(), t._version) for t in tensors)


def invalidate_caches(module):
    """Drop every derived-parameter cache (re-parametrised GDN operands, packed factorised-prior networks, CDF tables,
    masked context weights) below ``module``.  The caches are keyed on (storage pointer, tensor version), which
    ``param.copy_()`` / ``load_state_dict`` / optimiser steps bump -- but edits through ``param.data`` (the reference
    style: ``weight.data.fill_(0.0)``, Flex-Rate.../b_model/layers.py:125-126) do not.  Called automatically after
    ``load_state_dict`` and by ``update()``; call it yourself after a ``.data`` edit."""
    for m in module.modules():
        for k in [k for k in m.__dict__ if k.startswith("_b200vc_")]:
            del m.__dict__[k]
    return module


def _invalidate_after_load(module, incompatible_keys=None):
    invalidate_caches(module)


# ------------------------------------------------------------------------------------------ GDN
def gdn_params(mod):
    """Re-parametrised beta/gamma (+ operand images), recomputed only when the stored parameters change
    (the reference recomputes them on every call: 6 tiny kernels per GDN)."""
    return _cached(
        mod, "_b200vc_gdn", _pkey(mod.beta, mod.gamma),
        lambda: ops.gdn_prepare(mod.beta, mod.gamma, mod.beta_reparam.lower_bound.bound.item(),
                                mod.gamma_reparam.lower_bound.bound.item(), mod.beta_reparam.pedestal.item()))


def gdn_forward(mod, x, addend=None, inplace=None):
    return ops.gdn(x, gdn_params(mod), inverse=bool(mod.inverse), addend=addend, inplace=inplace)


class LowerBound(nn.Module):
    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))


class NonNegativeParametrizer(nn.Module):
    def __init__(self, minimum=0.0, reparam_offset=2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        self.register_buffer("pedestal", torch.Tensor([self.reparam_offset ** 2]))
        self.lower_bound = LowerBound((self.minimum + self.reparam_offset ** 2) ** 0.5)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))


class GDN(nn.Module):
    """compressai.layers.GDN: y_i = x_i * rsqrt(beta_i + sum_j gamma_ij x_j^2); ``inverse`` -> sqrt."""

    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=float(beta_min))
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(in_channels)))
        self.register_load_state_dict_post_hook(_invalidate_after_load)

    forward = gdn_forward


def conv3x3(in_ch, out_ch, stride=1):
    return nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1)


def conv1x1(in_ch, out_ch, stride=1):
    return nn.Conv2d(in_ch, out_ch, kernel_size=1, stride=stride)


def subpel_conv3x3(in_ch, out_ch, r=1):
    return nn.Sequential(nn.Conv2d(in_ch, out_ch * r ** 2, kernel_size=3, padding=1), nn.PixelShuffle(r))


def res_stride_forward(mod, x):
    """ResidualBlockWithStride.forward with ``out += identity`` folded into the GDN kernel's epilogue."""
    out = mod.conv2(mod.leaky_relu(mod.conv1(x)))
    if mod.skip is not None:
        return gdn_forward(mod.gdn, out, addend=mod.skip(x))      # the skip tensor is ours: accumulate into it
    # stride 1, in_ch == out_ch: the identity IS the caller's input -- CompressAI's `out += identity` never writes x
    return gdn_forward(mod.gdn, out, addend=x, inplace=False)


def res_upsample_forward(mod, x):
    """ResidualBlockUpsample.forward with the skip add folded into the IGDN kernel's epilogue."""
    out = mod.conv(mod.leaky_relu(mod.subpel_conv(x)))
    return gdn_forward(mod.igdn, out, addend=mod.upsample(x))


class ResidualBlockWithStride(nn.Module):
    def __init__(self, in_ch, out_ch, stride=2):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch, stride=stride)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.gdn = GDN(out_ch)
        self.skip = conv1x1(in_ch, out_ch, stride=stride) if (stride != 1 or in_ch != out_ch) else None

    forward = res_stride_forward


class ResidualBlockUpsample(nn.Module):
    def __init__(self, in_ch, out_ch, upsample=2):
        super().__init__()
        self.subpel_conv = subpel_conv3x3(in_ch, out_ch, upsample)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv = conv3x3(out_ch, out_ch)
        self.igdn = GDN(out_ch, inverse=True)
        self.upsample = subpel_conv3x3(in_ch, out_ch, upsample)

    forward = res_upsample_forward


class ResidualBlock(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.skip = conv1x1(in_ch, out_ch) if in_ch != out_ch else None

    def forward(self, x):
        out = self.leaky_relu(self.conv2(self.leaky_relu(self.conv1(x))))
        return out + (self.skip(x) if self.skip is not None else x)


# ------------------------------------------------------------------------------- entropy models
class EntropyModel(nn.Module):
    def __init__(self, likelihood_bound=1e-9, entropy_coder=None, entropy_coder_precision=16):
        super().__init__()
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self.register_load_state_dict_post_hook(_invalidate_after_load)


def _lik_bound(mod):
    return _cached(mod, "_b200vc_lb", _pkey(mod.likelihood_lower_bound.bound),
                   lambda: mod.likelihood_lower_bound.bound.item()) if mod.use_likelihood_bound else 0.0


def eb_packed(mod):
    mats = [getattr(mod, f"_matrix{i}") for i in range(5)]
    bias = [getattr(mod, f"_bias{i}") for i in range(5)]
    facs = [getattr(mod, f"_factor{i}") for i in range(4)]
    if tuple(mod.filters) != (3, 3, 3, 3):
        raise NotImplementedError("b200vc EntropyBottleneck kernel supports filters=(3,3,3,3) (the CompressAI default)")
    return _cached(mod, "_b200vc_eb", _pkey(*mats, *bias, *facs, mod.quantiles),
                   lambda: ops.eb_prepare(mats, bias, facs, mod.quantiles))


def eb_forward(mod, x, training=None):
    """EntropyBottleneck.forward (eval): (z_hat, likelihood)."""
    _eval_only(mod, training)
    r = ops.entropy_bottleneck(x, eb_packed(mod), lik_bound=_lik_bound(mod), want_bits=False)
    return r["z_hat"], r["lik"]


def eb_tables(mod):
    """Quantised CDF rows of the factorised prior (EntropyBottleneck.update), cached per weight version."""
    from . import coding
    mats = [getattr(mod, f"_matrix{i}") for i in range(5)]
    bias = [getattr(mod, f"_bias{i}") for i in range(5)]
    facs = [getattr(mod, f"_factor{i}") for i in range(4)]
    return _cached(mod, "_b200vc_eb_tables", _pkey(*mats, *bias, *facs, mod.quantiles),
                   lambda: coding.bottleneck_tables(eb_packed(mod), mod.quantiles, mod.quantiles.device))


def _channel_indexes(shape, device):
    C = shape[1]
    return torch.arange(C, dtype=torch.int32, device=device).view(1, C, 1, 1).expand(1, C, shape[2], shape[3])


def eb_compress(mod, x):
    """EntropyBottleneck.compress: one byte string per batch item (round(x - median) coded with the channel's CDF)."""
    from . import coding
    r = ops.entropy_bottleneck(x, eb_packed(mod), want_z_hat=False, want_lik=False, want_bits=False, want_symbols=True)
    idx = _channel_indexes(x.shape, x.device)
    tables = eb_tables(mod)
    return [coding.rans_encode(r["symbols"][i], idx, tables) for i in range(x.shape[0])]


def eb_decompress(mod, strings, size):
    """EntropyBottleneck.decompress: byte strings + spatial size -> z_hat [N, C, *size]."""
    from . import coding
    C = mod.quantiles.shape[0]
    dev = mod.quantiles.device
    shape = (1, C, int(size[0]), int(size[1]))
    idx = _channel_indexes(shape, dev)
    tables = eb_tables(mod)
    med = mod.quantiles[:, 0, 1].detach().view(1, C, 1, 1)
    out = [coding.rans_decode(s, idx, tables).view(shape).float() + med for s in strings]
    return torch.cat(out, 0)


class EntropyBottleneck(EntropyModel):
    def __init__(self, channels, *args, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        filt = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = math.log(math.expm1(1 / scale / filt[i + 1]))
            self.register_parameter(f"_matrix{i:d}",
                                    nn.Parameter(torch.full((self.channels, filt[i + 1], filt[i]), init)))
            bias = torch.empty(self.channels, filt[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))
            if i < len(self.filters):
                self.register_parameter(f"_factor{i:d}", nn.Parameter(torch.zeros(self.channels, filt[i + 1], 1)))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles = nn.Parameter(init.repeat(self.channels, 1, 1))
        target = math.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    forward = eb_forward
    compress = eb_compress
    decompress = eb_decompress

    def update(self, force=False):
        """CompressAI rebuilds the quantised CDFs here; b200vc builds them lazily on the device at the first
        ``compress`` / ``decompress`` (keyed on the parameter versions), so this only reports success."""
        return True


def gc_forward(mod, inputs, scales, means=None, training=None):
    """GaussianConditional.forward (eval): (y_hat, likelihood)."""
    _eval_only(mod, training)
    if means is None:
        means = torch.zeros_like(inputs)
    r = ops.gauss_cond(inputs, scales, means, scale_bound=_scale_bound(mod), lik_bound=_lik_bound(mod),
                       want_bits=False)
    return r["y_hat"], r["lik"]


def _scale_bound(mod):
    return _cached(mod, "_b200vc_sb", _pkey(mod.lower_bound_scale.bound), lambda: mod.lower_bound_scale.bound.item())


def gc_build_indexes(mod, scales):
    """GaussianConditional.build_indexes: one kernel instead of a 63-step compare loop."""
    z = torch.zeros_like(scales)
    r = ops.gauss_cond(z, scales, z, scale_bound=_scale_bound(mod), lik_bound=0.0, want_y_hat=False,
                       want_lik=False, want_bits=False, want_symbols=True, scale_table=mod.scale_table)
    return r["indexes"]


def gc_quantize(mod, inputs, mode, means=None):
    """EntropyModel.quantize for modes 'dequantize' and 'symbols' (inference)."""
    if mode not in ("dequantize", "symbols"):
        raise NotImplementedError(f"b200vc quantize: mode {mode!r} is not part of the inference path")
    if means is None:
        means = torch.zeros_like(inputs)
    ones = torch.ones_like(inputs)
    if mode == "dequantize":
        return ops.gauss_cond(inputs, ones, means, want_lik=False, want_bits=False)["y_hat"]
    tab = mod.scale_table if mod.scale_table.numel() >= 2 else get_scale_table().to(inputs.device)
    return ops.gauss_cond(inputs, ones, means, want_y_hat=False, want_lik=False, want_bits=False,
                          want_symbols=True, scale_table=tab)["symbols"]


def gc_tables(mod):
    from . import coding
    if mod.scale_table.numel() < 2:
        raise RuntimeError("GaussianConditional: call update(force=True) before compress / decompress")
    return _cached(mod, "_b200vc_gc_tables", _pkey(mod.scale_table),
                   lambda: coding.gaussian_tables(mod.scale_table.detach().cpu(), mod.scale_table.device,
                                                  tail_mass=mod.tail_mass, scale_bound=_scale_bound(mod)))


def gc_compress(mod, inputs, indexes, means=None):
    """GaussianConditional.compress: symbols = round(inputs - means) coded with the CDF row ``indexes``."""
    from . import coding
    sym = gc_quantize(mod, inputs, "symbols", means)
    tables = gc_tables(mod)
    return [coding.rans_encode(sym[i], indexes[i], tables) for i in range(inputs.shape[0])]


def gc_decompress(mod, strings, indexes, dtype=torch.float, means=None):
    """GaussianConditional.decompress: byte strings + indexes (+ means) -> dequantised values."""
    from . import coding
    tables = gc_tables(mod)
    out = torch.stack([coding.rans_decode(s, indexes[i], tables).view(indexes[i].shape)
                       for i, s in enumerate(strings)], 0).to(dtype)
    return out + means if means is not None else out


class GaussianConditional(EntropyModel):
    def __init__(self, scale_table, *args, scale_bound=0.11, tail_mass=1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            self.lower_bound_scale = LowerBound(scale_table[0])
        elif scale_bound > 0:
            self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer(
            "scale_table", torch.Tensor(tuple(float(s) for s in scale_table)) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)

    def update_scale_table(self, scale_table, force=False):
        self.scale_table = torch.as_tensor(scale_table, dtype=torch.float32, device=self.scale_table.device)
        return True

    forward = gc_forward
    build_indexes = gc_build_indexes
    quantize = gc_quantize
    compress = gc_compress
    decompress = gc_decompress


# ----------------------------------------------------------------------------------- hyperprior
def hyperprior_forward(mod, x):
    """MeanScaleHyperprior.forward (compressai/models/priors.py) -- API-compatible result dict."""
    y = mod.g_a(x)
    z = mod.h_a(y)
    z_hat, z_lik = mod.entropy_bottleneck(z)
    scales_hat, means_hat = mod.h_s(z_hat).chunk(2, 1)
    y_hat, y_lik = mod.gaussian_conditional(y, scales_hat, means=means_hat)
    return {"x_hat": mod.g_s(y_hat), "likelihoods": {"y": y_lik, "z": z_lik}}


def hyperprior_forward_bits(mod, x, want_symbols=False):
    """Same computation, but the likelihood tensors are never written to HBM: the quantise->CDF-difference->
    -log2 pass reduces straight to per-sample bit totals.  Returns (x_hat, bits_y[N], bits_z[N]) (float64).
    ``want_symbols=True`` (the encoder's view of the same pass: LHBDC/model/layers.py:93-104) additionally returns a
    dict with the int32 y symbols + CDF indexes, the z symbols and the latents they were taken from."""
    eb, gc = mod.entropy_bottleneck, mod.gaussian_conditional
    y = mod.g_a(x)
    z = mod.h_a(y)
    rz = ops.entropy_bottleneck(z, eb_packed(eb), lik_bound=_lik_bound(eb), want_lik=False, want_symbols=want_symbols)
    scales_hat, means_hat = mod.h_s(rz["z_hat"]).chunk(2, 1)
    ry = ops.gauss_cond(y, scales_hat, means_hat, scale_bound=_scale_bound(gc), lik_bound=_lik_bound(gc),
                        want_lik=False, want_symbols=want_symbols,
                        scale_table=gc.scale_table if want_symbols else None)
    x_hat = mod.g_s(ry["y_hat"])
    if want_symbols:
        return x_hat, ry["bits"], rz["bits"], {
            "y_symbols": ry["symbols"], "y_indexes": ry["indexes"], "z_symbols": rz["symbols"], "shape": z.size()[-2:],
            "y": y, "z": z, "scales_hat": scales_hat, "means_hat": means_hat, "y_hat": ry["y_hat"], "z_hat": rz["z_hat"]}
    return x_hat, ry["bits"], rz["bits"]


def hyperprior_symbols(mod, x):
    """Tensor half of ``compress`` (LHBDC/model/layers.py:93-104): everything the rANS coder is fed --
    int32 y symbols + CDF indexes and z symbols -- in two kernels (the reference: 63-step index loop,
    separate quantise, EB round trip through the CPU coder)."""
    eb, gc = mod.entropy_bottleneck, mod.gaussian_conditional
    y = mod.g_a(x)
    z = mod.h_a(y)
    rz = ops.entropy_bottleneck(z, eb_packed(eb), lik_bound=_lik_bound(eb), want_lik=False, want_bits=False,
                                want_symbols=True)
    scales_hat, means_hat = mod.h_s(rz["z_hat"]).chunk(2, 1)
    ry = ops.gauss_cond(y, scales_hat, means_hat, scale_bound=_scale_bound(gc), lik_bound=_lik_bound(gc),
                        want_y_hat=False, want_lik=False, want_bits=False, want_symbols=True,
                        scale_table=gc.scale_table)
    return {"y_symbols": ry["symbols"], "y_indexes": ry["indexes"], "z_symbols": rz["symbols"],
            "shape": z.size()[-2:]}


def hyperprior_compress(mod, x):
    """``compress`` of the reference's hyperprior compressors (LHBDC/model/layers.py:93-104)."""
    y = mod.g_a(x)
    z = mod.h_a(y)
    z_strings = mod.entropy_bottleneck.compress(z)
    z_hat = mod.entropy_bottleneck.decompress(z_strings, z.size()[-2:])
    scales_hat, means_hat = mod.h_s(z_hat).chunk(2, 1)
    indexes = mod.gaussian_conditional.build_indexes(scales_hat)
    y_strings = mod.gaussian_conditional.compress(y, indexes, means=means_hat)
    return {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}


def hyperprior_decompress(mod, strings, shape):
    """``decompress`` (LHBDC/model/layers.py:106-117)."""
    assert isinstance(strings, list) and len(strings) == 2
    z_hat = mod.entropy_bottleneck.decompress(strings[1], shape)
    scales_hat, means_hat = mod.h_s(z_hat).chunk(2, 1)
    indexes = mod.gaussian_conditional.build_indexes(scales_hat)
    y_hat = mod.gaussian_conditional.decompress(strings[0], indexes, means=means_hat)
    return {"x_hat": mod.g_s(y_hat)}


class MeanScaleHyperprior(nn.Module):
    """Container with CompressAI's member names; sub-networks are supplied by the subclasses."""

    def __init__(self, N, M):
        super().__init__()
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.gaussian_conditional = GaussianConditional(None)
        self.N, self.M = int(N), int(M)

    forward = hyperprior_forward
    forward_bits = hyperprior_forward_bits
    symbols = hyperprior_symbols
    compress = hyperprior_compress
    decompress = hyperprior_decompress

    def update(self, scale_table=None, force=False):
        """Installs the Gaussian scale table (``model.mv_compressor.update(force=True)``, LHBDC/encode_B.py:34).
        The quantised CDF rows are built lazily on the device at the first compress / decompress; every derived
        cache below this module is dropped (the reference calls ``update`` after loading / editing weights)."""
        invalidate_caches(self)
        if scale_table is None:
            scale_table = get_scale_table()
        return self.gaussian_conditional.update_scale_table(scale_table, force=force)


# ------------------------------------------------------------------- joint autoregressive container (ICIP codecs)
def _conv5(i, o, stride=2):
    return nn.Conv2d(i, o, kernel_size=5, stride=stride, padding=2)


def _deconv5(i, o, stride=2):
    return nn.ConvTranspose2d(i, o, kernel_size=5, stride=stride, output_padding=stride - 1, padding=2)


# CompressAI's `mbt2018-mean` configurations (compressai/zoo/image.py: cfgs["mbt2018-mean"]): quality -> (N, M)
MBT2018_MEAN_CFG = {1: (128, 192), 2: (128, 192), 3: (128, 192), 4: (128, 192),
                    5: (192, 320), 6: (192, 320), 7: (192, 320), 8: (192, 320)}


class Mbt2018Mean(MeanScaleHyperprior):
    """``compressai.models.MeanScaleHyperprior`` with its own sub-networks -- the I-frame codec the reference's
    evaluation loops build with ``mbt2018_mean(args.i_qual, "mse", pretrained=True)`` (LHBDC/test/testing.py:209,
    Flex-Rate.../test/testing.py) and call through ``image_compress`` (testing.py:78-86).  Same member names and
    state-dict keys as CompressAI's class, so ``compressai.zoo`` checkpoints load unchanged; GDN / IGDN (C = N: 128 or
    192, both on the tcgen05 kernel), the factorised prior and the Gaussian conditional run on the b200vc kernels."""

    def __init__(self, N=192, M=320):
        super().__init__(N, M)
        self.g_a = nn.Sequential(_conv5(3, N), GDN(N), _conv5(N, N), GDN(N), _conv5(N, N), GDN(N), _conv5(N, M))
        self.g_s = nn.Sequential(_deconv5(M, N), GDN(N, inverse=True), _deconv5(N, N), GDN(N, inverse=True),
                                 _deconv5(N, N), GDN(N, inverse=True), _deconv5(N, 3))
        self.h_a = nn.Sequential(nn.Conv2d(M, N, 3, 1, 1), nn.LeakyReLU(inplace=True), _conv5(N, N),
                                 nn.LeakyReLU(inplace=True), _conv5(N, N))
        self.h_s = nn.Sequential(_deconv5(N, M), nn.LeakyReLU(inplace=True), _deconv5(M, M * 3 // 2),
                                 nn.LeakyReLU(inplace=True), nn.Conv2d(M * 3 // 2, M * 2, 3, 1, 1))


def mbt2018_mean(quality, metric="mse", pretrained=False):
    """``compressai.zoo.mbt2018_mean`` without the download: the architecture for ``quality`` with fresh weights
    (load a CompressAI checkpoint with ``load_state_dict``; there is no network access here for ``pretrained``)."""
    if metric != "mse":
        raise ValueError(f"mbt2018_mean: unknown metric {metric!r}")
    if quality not in MBT2018_MEAN_CFG:
        raise ValueError(f'Invalid quality "{quality}", should be between (1, 8)')
    if pretrained:
        raise RuntimeError("mbt2018_mean: pretrained weights cannot be downloaded here; build the model and "
                           "load_state_dict() a compressai.zoo checkpoint")
    return Mbt2018Mean(*MBT2018_MEAN_CFG[quality])


def image_compress(im, compressor):
    """LHBDC/test/testing.py:78-86 -- ``(decoded image, size in bits)`` of an I-frame codec; here the likelihood
    tensors are never written (``forward_bits``), the size comes back as a float64 tensor [N] without a host sync."""
    x_hat, bits_y, bits_z = compressor.forward_bits(im)
    return x_hat, bits_y + bits_z


class MaskedConv2d(nn.Conv2d):
    """compressai.layers.MaskedConv2d (causal PixelCNN mask).  Present in the ICIP checkpoints as
    ``context_prediction`` (created by the base class below), never called by those models."""

    def __init__(self, *args, mask_type="A", **kwargs):
        super().__init__(*args, **kwargs)
        self.register_buffer("mask", torch.ones_like(self.weight.data))
        _, _, h, w = self.mask.size()
        self.mask[:, :, h // 2, w // 2 + (mask_type == "B"):] = 0
        self.mask[:, :, h // 2 + 1:] = 0

    def forward(self, x):
        self.weight.data *= self.mask
        return super().forward(x)


class CheckerboardContext(nn.Conv2d):
    """ICIP2024/src/model/layers.py:6-29: 5x5 context convolution restricted to the anchor checkerboard.  The
    reference re-multiplies the weight by the mask on every call (quirk B.9); masking is idempotent, so here it is
    applied once per weight version."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.register_buffer("mask", torch.zeros_like(self.weight.data))
        self.mask[:, :, 0::2, 1::2] = 1
        self.mask[:, :, 1::2, 0::2] = 1
        self.register_load_state_dict_post_hook(_invalidate_after_load)

    def forward(self, x):
        key = _pkey(self.weight)
        if self.__dict__.get("_b200vc_masked") != key:
            with torch.no_grad():
                self.weight.mul_(self.mask)
            self.__dict__["_b200vc_masked"] = _pkey(self.weight)
        return super().forward(x)


class JointAutoregressiveHierarchicalPriors(MeanScaleHyperprior):
    """compressai.models.JointAutoregressiveHierarchicalPriors: the base class of the ICIP codecs' ``Offset_ELIC`` /
    ``Res_ELIC`` (ICIP2024/src/model/compression_bottlenecks.py:72,313).  The subclasses replace ``h_a``, ``h_s`` and
    ``entropy_parameters`` and never call ``g_a``, ``g_s`` or ``context_prediction``; the members are built as
    CompressAI builds them so that reference checkpoints load with ``strict=True``."""

    def __init__(self, N=192, M=192, **kwargs):
        super().__init__(N=N, M=M)
        self.g_a = nn.Sequential(_conv5(3, N), GDN(N), _conv5(N, N), GDN(N), _conv5(N, N), GDN(N), _conv5(N, M))
        self.g_s = nn.Sequential(_deconv5(M, N), GDN(N, inverse=True), _deconv5(N, N), GDN(N, inverse=True),
                                 _deconv5(N, N), GDN(N, inverse=True), _deconv5(N, 3))
        self.h_a = nn.Sequential(nn.Conv2d(M, N, 3, 1, 1), nn.LeakyReLU(inplace=True), _conv5(N, N),
                                 nn.LeakyReLU(inplace=True), _conv5(N, N))
        self.h_s = nn.Sequential(_deconv5(N, M), nn.LeakyReLU(inplace=True), _deconv5(M, M * 3 // 2),
                                 nn.LeakyReLU(inplace=True), nn.Conv2d(M * 3 // 2, M * 2, 3, 1, 1))
        self.entropy_parameters = nn.Sequential(
            nn.Conv2d(M * 12 // 3, M * 10 // 3, 1), nn.LeakyReLU(inplace=True),
            nn.Conv2d(M * 10 // 3, M * 8 // 3, 1), nn.LeakyReLU(inplace=True), nn.Conv2d(M * 8 // 3, M * 6 // 3, 1))
        self.context_prediction = MaskedConv2d(M, 2 * M, kernel_size=5, padding=2, stride=1)
        self.gaussian_conditional = GaussianConditional(None)
