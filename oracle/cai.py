"""TEST INFRASTRUCTURE (oracle) -- restatement of the CompressAI 1.1.x operators the
reference instantiates.  parity unpinned: ``compressai==1.1.8`` is pinned at
``LHBDC/environment.yml:142`` but its source is not under ``/root/reference`` and it is
not installable offline, so this file restates the library's published algorithm
(InterDigitalInc/CompressAI 1.1.x: ``compressai/layers/gdn.py``,
``compressai/ops/parametrizers.py``, ``compressai/entropy_models/entropy_models.py``,
``compressai/layers/layers.py``, ``compressai/models/priors.py``) with identical attribute
names and state-dict keys, anchored on the reference call sites:

* ``LHBDC/model/layers.py:6-17``  (imports: MeanScaleHyperprior, EntropyBottleneck,
  GaussianConditional, GDN, ResidualBlock, ResidualBlockUpsample, ResidualBlockWithStride,
  conv3x3, subpel_conv3x3)
* ``LHBDC/model/layers.py:93-117`` (``compress``/``decompress``: ``build_indexes``)
* ``Flex-Rate-Hier-Bidir-Video-Compression/b_model/layers.py:135-152`` (forward with gains)
* ``ICIP2023/src/model/elic.py:21-27`` (the scale table, restated in-repo by the reference)

Everything is plain eager torch (fp32) and runs on CPU or CUDA.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64


def get_scale_table(lo=SCALES_MIN, hi=SCALES_MAX, levels=SCALES_LEVELS):
    """``exp(linspace(ln lo, ln hi, levels))`` -- ICIP2023/src/model/elic.py:21-27."""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


# --------------------------------------------------------------------------- ops
class LowerBound(nn.Module):
    """max(x, bound) with a buffer named ``bound`` (compressai/ops/bound_ops.py)."""

    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return torch.max(x, self.bound)


class NonNegativeParametrizer(nn.Module):
    """compressai/ops/parametrizers.py: stored = sqrt(max(x+ped, ped)); used = max(stored, b)^2 - ped."""

    def __init__(self, minimum=0.0, reparam_offset=2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        pedestal = self.reparam_offset ** 2
        self.register_buffer("pedestal", torch.Tensor([pedestal]))
        bound = (self.minimum + self.reparam_offset ** 2) ** 0.5
        self.lower_bound = LowerBound(bound)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x):
        out = self.lower_bound(x)
        return out ** 2 - self.pedestal


class GDN(nn.Module):
    """y_i = x_i * rsqrt(beta_i + sum_j gamma_ij x_j^2)  (IGDN: * sqrt).  SURVEY A.1."""

    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        beta_min = float(beta_min)
        gamma_init = float(gamma_init)
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=beta_min)
        beta = torch.ones(in_channels)
        self.beta = nn.Parameter(self.beta_reparam.init(beta))
        self.gamma_reparam = NonNegativeParametrizer()
        gamma = gamma_init * torch.eye(in_channels)
        self.gamma = nn.Parameter(self.gamma_reparam.init(gamma))

    def forward(self, x):
        _, C, _, _ = x.size()
        beta = self.beta_reparam(self.beta)
        gamma = self.gamma_reparam(self.gamma).reshape(C, C, 1, 1)
        norm = F.conv2d(x ** 2, gamma, beta)
        norm = torch.sqrt(norm) if self.inverse else torch.rsqrt(norm)
        return x * norm


def conv3x3(in_ch, out_ch, stride=1):
    return nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1)


def conv1x1(in_ch, out_ch, stride=1):
    return nn.Conv2d(in_ch, out_ch, kernel_size=1, stride=stride)


def subpel_conv3x3(in_ch, out_ch, r=1):
    return nn.Sequential(nn.Conv2d(in_ch, out_ch * r ** 2, kernel_size=3, padding=1), nn.PixelShuffle(r))


class ResidualBlockWithStride(nn.Module):
    def __init__(self, in_ch, out_ch, stride=2):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch, stride=stride)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.gdn = GDN(out_ch)
        self.skip = conv1x1(in_ch, out_ch, stride=stride) if (stride != 1 or in_ch != out_ch) else None

    def forward(self, x):
        identity = x
        out = self.gdn(self.conv2(self.leaky_relu(self.conv1(x))))
        if self.skip is not None:
            identity = self.skip(x)
        out += identity
        return out


class ResidualBlockUpsample(nn.Module):
    def __init__(self, in_ch, out_ch, upsample=2):
        super().__init__()
        self.subpel_conv = subpel_conv3x3(in_ch, out_ch, upsample)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv = conv3x3(out_ch, out_ch)
        self.igdn = GDN(out_ch, inverse=True)
        self.upsample = subpel_conv3x3(in_ch, out_ch, upsample)

    def forward(self, x):
        out = self.igdn(self.conv(self.leaky_relu(self.subpel_conv(x))))
        out += self.upsample(x)
        return out


class ResidualBlock(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.skip = conv1x1(in_ch, out_ch) if in_ch != out_ch else None

    def forward(self, x):
        identity = x
        out = self.leaky_relu(self.conv2(self.leaky_relu(self.conv1(x))))
        if self.skip is not None:
            identity = self.skip(x)
        return out + identity


class AttentionBlock(nn.Module):  # imported (unused) at LHBDC/model/layers.py:10
    def __init__(self, N):
        super().__init__()
        raise NotImplementedError("AttentionBlock is imported but never instantiated by the reference")


# ---------------------------------------------------------------- entropy models
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

    def quantize(self, inputs, mode, means=None):
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            return inputs + torch.empty_like(inputs).uniform_(-0.5, 0.5)
        outputs = inputs.clone()
        if means is not None:
            outputs -= means
        outputs = torch.round(outputs)
        if mode == "dequantize":
            if means is not None:
                outputs += means
            return outputs
        return outputs.int()

    @staticmethod
    def dequantize(inputs, means=None, dtype=torch.float):
        if means is not None:
            outputs = inputs.type_as(means)
            outputs += means
        else:
            outputs = inputs.type(dtype)
        return outputs


class EntropyBottleneck(EntropyModel):
    """Factorised prior (Balle 2018), SURVEY A.3."""

    def __init__(self, channels, *args, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        filt = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        channels = self.channels
        for i in range(len(self.filters) + 1):
            init = math.log(math.expm1(1 / scale / filt[i + 1]))
            matrix = torch.Tensor(channels, filt[i + 1], filt[i])
            matrix.data.fill_(init)
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(matrix))
            bias = torch.Tensor(channels, filt[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))
            if i < len(self.filters):
                factor = torch.Tensor(channels, filt[i + 1], 1)
                nn.init.zeros_(factor)
                self.register_parameter(f"_factor{i:d}", nn.Parameter(factor))
        self.quantiles = nn.Parameter(torch.Tensor(channels, 1, 3))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles.data = init.repeat(self.quantiles.size(0), 1, 1)
        target = math.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def _get_medians(self):
        return self.quantiles[:, :, 1:2].detach()

    def _logits_cumulative(self, inputs, stop_gradient=False):
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = getattr(self, f"_matrix{i:d}")
            logits = torch.matmul(F.softplus(matrix), logits)
            logits = logits + getattr(self, f"_bias{i:d}")
            if i < len(self.filters):
                factor = getattr(self, f"_factor{i:d}")
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def _likelihood(self, inputs):
        lower = self._logits_cumulative(inputs - 0.5)
        upper = self._logits_cumulative(inputs + 0.5)
        sign = -torch.sign(lower + upper).detach()
        return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))

    def forward(self, x, training=None):
        if training is None:
            training = self.training
        perm = list(range(x.dim()))
        perm[0], perm[1] = perm[1], perm[0]
        x = x.permute(*perm).contiguous()  # C, B, ...
        shape = x.size()
        values = x.reshape(x.size(0), 1, -1)
        outputs = self.quantize(values, "noise" if training else "dequantize", self._get_medians())
        likelihood = self._likelihood(outputs)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        outputs = outputs.reshape(shape).permute(*perm).contiguous()
        likelihood = likelihood.reshape(shape).permute(*perm).contiguous()
        return outputs, likelihood


class GaussianConditional(EntropyModel):
    """SURVEY A.2."""

    def __init__(self, scale_table, *args, scale_bound=0.11, tail_mass=1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            self.lower_bound_scale = LowerBound(scale_table[0])
        elif scale_bound > 0:
            self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer(
            "scale_table", torch.Tensor(tuple(float(s) for s in scale_table)) if scale_table else torch.Tensor()
        )
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)

    def update_scale_table(self, scale_table, force=False):
        """Installs the table (the CDF build that real rANS coding needs is row f-1, not done here)."""
        self.scale_table = torch.as_tensor(scale_table, dtype=torch.float32, device=self.scale_table.device)
        return True

    @staticmethod
    def _standardized_cumulative(inputs):
        half = float(0.5)
        const = float(-(2 ** -0.5))
        return half * torch.erfc(const * inputs)

    def _likelihood(self, inputs, scales, means=None):
        half = float(0.5)
        values = inputs - means if means is not None else inputs
        scales = self.lower_bound_scale(scales)
        values = torch.abs(values)
        upper = self._standardized_cumulative((half - values) / scales)
        lower = self._standardized_cumulative((-half - values) / scales)
        return upper - lower

    def forward(self, inputs, scales, means=None, training=None):
        if training is None:
            training = self.training
        outputs = self.quantize(inputs, "noise" if training else "dequantize", means)
        likelihood = self._likelihood(outputs, scales, means)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        return outputs, likelihood

    def build_indexes(self, scales):
        scales = self.lower_bound_scale(scales)
        indexes = scales.new_full(scales.size(), len(self.scale_table) - 1).int()
        for s in self.scale_table[:-1]:
            indexes -= (scales <= s).int()
        return indexes


# ------------------------------------------------------------------------ models
class CompressionModel(nn.Module):
    def __init__(self, entropy_bottleneck_channels, init_weights=True):
        super().__init__()
        self.entropy_bottleneck = EntropyBottleneck(entropy_bottleneck_channels)
        if init_weights:
            for m in self.modules():
                if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                    nn.init.kaiming_normal_(m.weight)
                    if m.bias is not None:
                        nn.init.zeros_(m.bias)


def _conv5(i, o, stride=2):
    return nn.Conv2d(i, o, kernel_size=5, stride=stride, padding=2)


def _deconv5(i, o, stride=2):
    return nn.ConvTranspose2d(i, o, kernel_size=5, stride=stride, output_padding=stride - 1, padding=2)


class ScaleHyperprior(CompressionModel):
    def __init__(self, N, M, **kwargs):
        super().__init__(entropy_bottleneck_channels=N, **kwargs)
        self.g_a = nn.Sequential(_conv5(3, N), GDN(N), _conv5(N, N), GDN(N), _conv5(N, N), GDN(N), _conv5(N, M))
        self.g_s = nn.Sequential(
            _deconv5(M, N), GDN(N, inverse=True), _deconv5(N, N), GDN(N, inverse=True),
            _deconv5(N, N), GDN(N, inverse=True), _deconv5(N, 3),
        )
        self.h_a = nn.Sequential(
            nn.Conv2d(M, N, 3, 1, 1), nn.ReLU(inplace=True), _conv5(N, N), nn.ReLU(inplace=True), _conv5(N, N)
        )
        self.h_s = nn.Sequential(
            _deconv5(N, N), nn.ReLU(inplace=True), _deconv5(N, N), nn.ReLU(inplace=True),
            nn.Conv2d(N, M, 3, 1, 1), nn.ReLU(inplace=True),
        )
        self.gaussian_conditional = GaussianConditional(None)
        self.N = int(N)
        self.M = int(M)


class MeanScaleHyperprior(ScaleHyperprior):
    """SURVEY A.4; this is also the topology of ``compressai.zoo.mbt2018_mean`` (I-frame anchors,
    LHBDC/test/testing.py:209)."""

    def __init__(self, N, M, **kwargs):
        super().__init__(N, M, **kwargs)
        self.h_a = nn.Sequential(
            nn.Conv2d(M, N, 3, 1, 1), nn.LeakyReLU(inplace=True), _conv5(N, N), nn.LeakyReLU(inplace=True),
            _conv5(N, N),
        )
        self.h_s = nn.Sequential(
            _deconv5(N, M), nn.LeakyReLU(inplace=True), _deconv5(M, M * 3 // 2), nn.LeakyReLU(inplace=True),
            nn.Conv2d(M * 3 // 2, M * 2, 3, 1, 1),
        )

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(y)
        z_hat, z_likelihoods = self.entropy_bottleneck(z)
        gaussian_params = self.h_s(z_hat)
        scales_hat, means_hat = gaussian_params.chunk(2, 1)
        y_hat, y_likelihoods = self.gaussian_conditional(y, scales_hat, means=means_hat)
        x_hat = self.g_s(y_hat)
        out = {"x_hat": x_hat, "likelihoods": {"y": y_likelihoods, "z": z_likelihoods}}
        if getattr(self, "keep_latents", False):  # test hook (not CompressAI): what the entropy coder would see
            out["latents"] = {"y_hat": y_hat, "z_hat": z_hat, "scales_hat": scales_hat, "means_hat": means_hat}
        return out

    def update(self, scale_table=None, force=False):
        if scale_table is None:
            scale_table = get_scale_table()
        return self.gaussian_conditional.update_scale_table(scale_table, force=force)


class MaskedConv2d(nn.Conv2d):
    """compressai.layers.MaskedConv2d (PixelCNN-style causal mask; type 'A' hides the centre too).  The ICIP models
    inherit one as ``context_prediction`` from the joint-autoregressive container and never call it."""

    def __init__(self, *args, mask_type="A", **kwargs):
        super().__init__(*args, **kwargs)
        if mask_type not in ("A", "B"):
            raise ValueError(f'Invalid "mask_type" value "{mask_type}"')
        self.register_buffer("mask", torch.ones_like(self.weight.data))
        _, _, h, w = self.mask.size()
        self.mask[:, :, h // 2, w // 2 + (mask_type == "B"):] = 0
        self.mask[:, :, h // 2 + 1:] = 0

    def forward(self, x):
        self.weight.data *= self.mask
        return super().forward(x)


class JointAutoregressiveHierarchicalPriors(MeanScaleHyperprior):
    """compressai.models.JointAutoregressiveHierarchicalPriors (Minnen 2018), the base class of the ICIP codecs'
    ``Offset_ELIC`` / ``Res_ELIC`` / ``ELIC`` (ICIP2024/src/model/compression_bottlenecks.py:72, :313).  Those
    subclasses replace ``h_a``, ``h_s`` and ``entropy_parameters`` and never call ``g_a``, ``g_s`` or
    ``context_prediction`` -- but the members exist (state-dict keys of the reference checkpoints), so they are
    built here as CompressAI builds them."""

    def __init__(self, N=192, M=192, **kwargs):
        super().__init__(N=N, M=M, **kwargs)
        self.g_a = nn.Sequential(_conv5(3, N), GDN(N), _conv5(N, N), GDN(N), _conv5(N, N), GDN(N), _conv5(N, M))
        self.g_s = nn.Sequential(
            _deconv5(M, N), GDN(N, inverse=True), _deconv5(N, N), GDN(N, inverse=True),
            _deconv5(N, N), GDN(N, inverse=True), _deconv5(N, 3),
        )
        self.h_a = nn.Sequential(
            nn.Conv2d(M, N, 3, 1, 1), nn.LeakyReLU(inplace=True), _conv5(N, N), nn.LeakyReLU(inplace=True),
            _conv5(N, N),
        )
        self.h_s = nn.Sequential(
            _deconv5(N, M), nn.LeakyReLU(inplace=True), _deconv5(M, M * 3 // 2), nn.LeakyReLU(inplace=True),
            nn.Conv2d(M * 3 // 2, M * 2, 3, 1, 1),
        )
        self.entropy_parameters = nn.Sequential(
            nn.Conv2d(M * 12 // 3, M * 10 // 3, 1), nn.LeakyReLU(inplace=True),
            nn.Conv2d(M * 10 // 3, M * 8 // 3, 1), nn.LeakyReLU(inplace=True),
            nn.Conv2d(M * 8 // 3, M * 6 // 3, 1),
        )
        self.context_prediction = MaskedConv2d(M, 2 * M, kernel_size=5, padding=2, stride=1)
        self.gaussian_conditional = GaussianConditional(None)
        self.N = int(N)
        self.M = int(M)

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(y)
        z_hat, z_likelihoods = self.entropy_bottleneck(z)
        params = self.h_s(z_hat)
        y_hat = self.gaussian_conditional.quantize(y, "noise" if self.training else "dequantize")
        ctx_params = self.context_prediction(y_hat)
        gaussian_params = self.entropy_parameters(torch.cat((params, ctx_params), dim=1))
        scales_hat, means_hat = gaussian_params.chunk(2, 1)
        _, y_likelihoods = self.gaussian_conditional(y, scales_hat, means=means_hat)
        x_hat = self.g_s(y_hat)
        return {"x_hat": x_hat, "likelihoods": {"y": y_likelihoods, "z": z_likelihoods}}


class Cheng2020Anchor(JointAutoregressiveHierarchicalPriors):
    """Imported (never instantiated) at ICIP2024/src/model/compression_bottlenecks.py:7."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("Cheng2020Anchor is imported but never instantiated by the reference")
