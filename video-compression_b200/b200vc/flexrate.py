"""Host-side mirror of the Flex-Rate hierarchical bidirectional codec interface (reference:
``Flex-Rate-Hier-Bidir-Video-Compression/b_model/b_model.py`` ``BidirFlowRef``, ``b_model/layers.py``
``Gain_Module`` / ``FlowCompressor`` / ``ResidualCompressor``, ``b_model/unet.py`` ``UNet``).  Same names, call
signatures (``forward(x_before, x_current, x_after, n=[int], l=float, train=False) -> {"x_hat","size","rate"}``) and
state-dict keys; the hot path runs in the sm_100a kernels:

* ``backwarp`` x4 -> K-WARP, FLEX variant (zeros padding, half-pixel grid: b_model.py:99-112)
* sigmoid + normalised blend + residual (b_model.py:68-73) -> K-BLEND, NORMW form (one pass)
* every GDN / IGDN (+ residual add) -> K-GDN
* ``hyper_gain_unit`` -> EntropyBottleneck -> ``hyper_inv_gain_unit`` (layers.py:140-143) -> K-EB with gain prologue and
  inverse-gain epilogue; GaussianConditional -> ``inv_gain_unit`` (layers.py:145-146) -> K-GC with inverse-gain epilogue
* per-sample bit sums (b_model.py:80-90) -> fp64 partials inside the two entropy kernels

U-Nets and conv stacks stay on cuDNN (out of scope).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import modules as M
from . import ops
from .lhbdc import _Compressor


class UNetConvBlock(nn.Module):
    def __init__(self, in_size, out_size, padding):
        super().__init__()
        pad = int(padding)
        self.block = nn.Sequential(nn.Conv2d(in_size, out_size, kernel_size=3, padding=pad), nn.LeakyReLU(0.1),
                                   nn.Conv2d(out_size, out_size, kernel_size=3, padding=pad), nn.LeakyReLU(0.1))

    def forward(self, x):
        return self.block(x)


class UNetUpBlock(nn.Module):
    def __init__(self, in_size, out_size, padding):
        super().__init__()
        self.up = nn.Sequential(nn.Upsample(mode="bilinear", scale_factor=2),
                                nn.Conv2d(in_size, out_size, kernel_size=3, padding=1))
        self.conv_block = UNetConvBlock(in_size, out_size, padding)

    def forward(self, x, bridge):
        up = self.up(x)
        th, tw = up.shape[2:]
        dy, dx = (bridge.shape[2] - th) // 2, (bridge.shape[3] - tw) // 2
        return self.conv_block(torch.cat((up, bridge[:, :, dy:dy + th, dx:dx + tw]), 1))


class UNet(nn.Module):
    """b_model/unet.py (tunable U-Net): depth x [conv3-lrelu-conv3-lrelu], avg-pool down, bilinear up."""

    def __init__(self, in_channels=1, n_classes=2, depth=5, wf=5, padding=True):
        super().__init__()
        self.padding, self.depth = padding, depth
        widths = [2 ** (wf + i) for i in range(depth)]
        self.down_path = nn.ModuleList(
            UNetConvBlock(i, o, padding) for i, o in zip([in_channels] + widths[:-1], widths))
        self.midconv = nn.Conv2d(widths[-1], widths[-1], kernel_size=3, padding=1)
        self.up_path = nn.ModuleList(
            UNetUpBlock(widths[i + 1], widths[i], padding) for i in reversed(range(depth - 1)))
        self.last = nn.Conv2d(widths[0], n_classes, kernel_size=3, padding=1)

    def forward(self, x):
        skips = []
        for i, down in enumerate(self.down_path):
            x = down(x)
            if i + 1 < len(self.down_path):
                skips.append(x)
                x = F.avg_pool2d(x, 2)
        x = F.leaky_relu(self.midconv(x), negative_slope=0.1)
        for up, skip in zip(self.up_path, reversed(skips)):
            x = up(x, skip)
        return self.last(x)


class Gain_Module(nn.Module):
    """layers.py:40-73.  ``gain(n, l)`` is the per-channel vector; the multiply itself is fused into the entropy
    kernels wherever the scaled tensor is not needed by a convolution."""

    def __init__(self, n=6, N=128, bias=False, inv=False):
        super().__init__()
        self.gain_matrix = nn.Parameter(torch.ones(n, N))
        if bias:
            raise NotImplementedError("the reference instantiates Gain_Module with bias=False only")
        self.bias = False

    def gain(self, n, l):
        if l != 1:
            return (torch.abs(self.gain_matrix[n]) ** l) * (torch.abs(self.gain_matrix[[n[0] + 1]]) ** (1 - l))
        return torch.abs(self.gain_matrix[n])

    def forward(self, x, n=None, l=1):
        return self.gain(n, l).unsqueeze(2).unsqueeze(3) * x


class _GainedCompressor(_Compressor):
    def __init__(self, n, in_ch, out_ch, N):
        super().__init__(in_ch, N)
        if out_ch != in_ch:
            self.g_s[-1] = M.subpel_conv3x3(N, out_ch, 2)
        self.gain_unit = Gain_Module(n=n, N=N)
        self.inv_gain_unit = Gain_Module(n=n, N=N, inv=True)
        self.hyper_gain_unit = Gain_Module(n=n, N=N)
        self.hyper_inv_gain_unit = Gain_Module(n=n, N=N, inv=True)

    def _stages(self, x, n, l, want_lik):
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        y = self.g_a(x)
        scaled_y = self.gain_unit(y, n, l)          # needed dense by h_a (a convolution)
        z = self.h_a(scaled_y)
        rz = ops.entropy_bottleneck(z, M.eb_packed(eb), lik_bound=M._lik_bound(eb),
                                    gain=self.hyper_gain_unit.gain(n, l), inv_gain=self.hyper_inv_gain_unit.gain(n, l),
                                    want_lik=want_lik)
        scales_hat, means_hat = self.h_s(rz["z_hat"]).chunk(2, 1)
        ry = ops.gauss_cond(scaled_y, scales_hat, means_hat, scale_bound=M._scale_bound(gc),
                            lik_bound=M._lik_bound(gc), inv_gain=self.inv_gain_unit.gain(n, l), want_lik=want_lik)
        return self.g_s(ry["y_hat"]), ry, rz

    def update(self, scale_table=None, force=False):
        return super().update(scale_table, force)

    def forward(self, x, n=None, l=None, train=False):
        """layers.py:135-152 -- API-compatible dict with materialised likelihoods."""
        if train:
            raise NotImplementedError("b200vc mirrors the inference path (train=False) only")
        x_hat, ry, rz = self._stages(x, n, l, want_lik=True)
        return {"x_hat": x_hat, "likelihoods": {"y": ry["lik"], "z": rz["lik"]}}

    def forward_bits(self, x, n, l):
        x_hat, ry, rz = self._stages(x, n, l, want_lik=False)
        return x_hat, ry["bits"], rz["bits"]

    def compress(self, x, n, l):
        """layers.py:154-172.  Quirk B.4 is kept: the *unscaled* y is entropy-coded while forward() rates the
        scaled one."""
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        y = self.g_a(x)
        scaled_y = self.gain_unit(y, n, l)
        z = self.h_a(scaled_y)
        scaled_z = self.hyper_gain_unit(z, n, l)
        z_strings = eb.compress(scaled_z)
        z_hat = eb.decompress(z_strings, z.size()[-2:])
        scales_hat, means_hat = self.h_s(self.hyper_inv_gain_unit(z_hat, n, l)).chunk(2, 1)
        indexes = gc.build_indexes(scales_hat)
        y_strings = gc.compress(y, indexes, means=means_hat)
        return {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}

    def decompress(self, strings, shape, n, l):
        """layers.py:174-189."""
        assert isinstance(strings, list) and len(strings) == 2
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        z_hat = eb.decompress(strings[1], shape)
        scales_hat, means_hat = self.h_s(self.hyper_inv_gain_unit(z_hat, n, l)).chunk(2, 1)
        indexes = gc.build_indexes(scales_hat)
        y_hat = gc.decompress(strings[0], indexes, means=means_hat)
        return {"x_hat": self.g_s(self.inv_gain_unit(y_hat, n, l)).clamp_(0, 1)}


class FlowCompressor(_GainedCompressor):
    def __init__(self, n=6, in_ch=19, out_ch=5, N=128, bias=False, **kwargs):
        super().__init__(n, in_ch, out_ch, N)
        self.g_s[-1][0].weight.data.fill_(0.0)  # layers.py:125-126
        self.g_s[-1][0].bias.data.fill_(0.0)


class ResidualCompressor(_GainedCompressor):
    def __init__(self, n=6, in_ch=3, N=128, bias=False, **kwargs):
        super().__init__(n, in_ch, in_ch, N)


class BidirFlowRef(nn.Module):
    """Bidirectional compression with flow refinement (b_model.py:21-112)."""

    def __init__(self, n=6, N=128):
        super().__init__()
        self.flow_predictor = UNet(6, 4, 5)
        self.Mask = UNet(16, 2, 4)
        self.flow_compressor = FlowCompressor(n=n, in_ch=19, out_ch=4, N=N, bias=False)
        self.residual_compressor = ResidualCompressor(n=n, in_ch=3, N=N, bias=False)

    def backwarp(self, img, flow):
        return ops.backwarp(img, flow, "flex")

    def process(self, x0, x1, t=0.5):
        """b_model.py:34-45: flow predictor, then K-WARP2 (Flex form): linear-motion glue + both warps + concat."""
        flow = self.flow_predictor(torch.cat((x0, x1), 1))
        conc = ops.warp2_flex(x0, x1, flow[:, 0:2], flow[:, 2:4], "linear", t)   # cat(ft0, ft1, x0, x1, xt1, xt2)
        return conc[:, 0:2], conc[:, 2:4], conc

    def forward_device(self, x_before, x_current, x_after, n, l):
        mv_before, mv_after, x_conc = self.process(x_before, x_after)
        flow_hat, fy, fz = self.flow_compressor.forward_bits(torch.cat((x_conc, x_current), 1), n, l)
        # mv refinement + both warps + the mask net's 16-channel input (b_model.py:58-66), one kernel
        temp = ops.warp2_flex(x_before, x_after, x_conc[:, 0:4], flow_hat[:, 0:4], "refine")
        logits = self.Mask(temp)
        x_comp, residual, _ = ops.blend_residual("normw", logits, temp[:, 10:13], temp[:, 13:16], x_current)
        res_hat, ry, rz = self.residual_compressor.forward_bits(residual, n, l)
        return x_comp + res_hat, (fy + fz) + (ry + rz)

    def forward(self, x_before, x_current, x_after, n=None, l=1, train=False):
        """b_model.py:49-96: ``size`` per sample (bits), ``rate`` = size / (H*W)."""
        if train:
            raise NotImplementedError("b200vc mirrors the inference path (train=False) only")
        x_hat, bits = self.forward_device(x_before, x_current, x_after, n, l)
        size = bits.float()
        return {"x_hat": x_hat, "size": size, "rate": size / (x_current.shape[2] * x_current.shape[3])}
