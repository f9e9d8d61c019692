"""TEST INFRASTRUCTURE (oracle).  Torch restatement of the Flex-Rate hierarchical bidirectional codec
(reference: ``Flex-Rate-Hier-Bidir-Video-Compression/b_model/{b_model,layers,unet}.py``) on top of ``oracle/cai.py``;
device-agnostic, same module tree / state-dict keys / construction order as the reference (checked against the
reference's own classes, imported through ``oracle/shim.py``, by ``oracle/make_golden.py``).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import cai, warp
from .lhbdc import _analysis, _hyper_analysis, _hyper_synthesis, _synthesis


# ------------------------------------------------------------------ U-Net (b_model/unet.py)
class UNetConvBlock(nn.Module):
    def __init__(self, in_size, out_size, padding):
        super().__init__()
        self.block = nn.Sequential(
            nn.Conv2d(in_size, out_size, kernel_size=3, padding=int(padding)), nn.LeakyReLU(0.1),
            nn.Conv2d(out_size, out_size, kernel_size=3, padding=int(padding)), nn.LeakyReLU(0.1),
        )

    def forward(self, x):
        return self.block(x)


class UNetUpBlock(nn.Module):
    def __init__(self, in_size, out_size, padding):
        super().__init__()
        self.up = nn.Sequential(nn.Upsample(mode="bilinear", scale_factor=2),
                                nn.Conv2d(in_size, out_size, kernel_size=3, padding=1))
        self.conv_block = UNetConvBlock(in_size, out_size, padding)

    @staticmethod
    def center_crop(layer, target_size):
        _, _, h, w = layer.size()
        dy, dx = (h - target_size[0]) // 2, (w - target_size[1]) // 2
        return layer[:, :, dy:dy + target_size[0], dx:dx + target_size[1]]

    def forward(self, x, bridge):
        up = self.up(x)
        return self.conv_block(torch.cat((up, self.center_crop(bridge, up.shape[2:])), 1))


class UNet(nn.Module):
    def __init__(self, in_channels=1, n_classes=2, depth=5, wf=5, padding=True):
        super().__init__()
        self.padding, self.depth = padding, depth
        prev = in_channels
        self.down_path = nn.ModuleList()
        for i in range(depth):
            self.down_path.append(UNetConvBlock(prev, 2 ** (wf + i), padding))
            prev = 2 ** (wf + i)
        self.midconv = nn.Conv2d(prev, prev, kernel_size=3, padding=1)
        self.up_path = nn.ModuleList()
        for i in reversed(range(depth - 1)):
            self.up_path.append(UNetUpBlock(prev, 2 ** (wf + i), padding))
            prev = 2 ** (wf + i)
        self.last = nn.Conv2d(prev, n_classes, kernel_size=3, padding=1)

    def forward(self, x):
        blocks = []
        for i, down in enumerate(self.down_path):
            x = down(x)
            if i != len(self.down_path) - 1:
                blocks.append(x)
                x = F.avg_pool2d(x, 2)
        x = F.leaky_relu(self.midconv(x), negative_slope=0.1)
        for i, up in enumerate(self.up_path):
            x = up(x, blocks[-i - 1])
        return self.last(x)


# ------------------------------------------------------------------ gained hyperpriors (b_model/layers.py)
class Gain_Module(nn.Module):
    """layers.py:40-73: per-channel gain, geometric interpolation between levels n and n+1 when l != 1."""

    def __init__(self, n=6, N=128, bias=False, inv=False):
        super().__init__()
        self.gain_matrix = nn.Parameter(torch.ones(n, N))
        self.bias = bias
        if bias:
            self.bias = nn.Parameter(torch.ones(N))

    def gain(self, n, l):
        if l != 1:
            g1 = self.gain_matrix[n]
            g2 = self.gain_matrix[[n[0] + 1]]
            return (torch.abs(g1) ** l) * (torch.abs(g2) ** (1 - l))
        return torch.abs(self.gain_matrix[n])

    def forward(self, x, n=None, l=1):
        out = self.gain(n, l).unsqueeze(2).unsqueeze(3) * x
        if self.bias:
            out += self.bias[n]
        return out


class _GainedHyperprior(cai.MeanScaleHyperprior):
    def __init__(self, n, in_ch, out_ch, N, bias, zero_last):
        super().__init__(N=N, M=N)
        self.g_a = _analysis(in_ch, N)
        self.h_a = _hyper_analysis(N)
        self.h_s = _hyper_synthesis(N)
        self.g_s = _synthesis(N, out_ch)
        if zero_last:  # layers.py:125-126
            self.g_s[-1][0].weight.data.fill_(0.0)
            self.g_s[-1][0].bias.data.fill_(0.0)
        self.gain_unit = Gain_Module(n=n, N=N, bias=bias, inv=False)
        self.inv_gain_unit = Gain_Module(n=n, N=N, bias=bias, inv=True)
        self.hyper_gain_unit = Gain_Module(n=n, N=N, bias=bias, inv=False)
        self.hyper_inv_gain_unit = Gain_Module(n=n, N=N, bias=bias, inv=True)

    def forward(self, x, n=None, l=None, train=False):
        """layers.py:135-152 / 249-266."""
        self.training = train
        y = self.g_a(x)
        scaled_y = self.gain_unit(y, n, l)
        z = self.h_a(scaled_y)
        scaled_z = self.hyper_gain_unit(z, n, l)
        z_hat, z_lik = self.entropy_bottleneck(scaled_z)
        scaled_z_hat = self.hyper_inv_gain_unit(z_hat, n, l)
        scales_hat, means_hat = self.h_s(scaled_z_hat).chunk(2, 1)
        y_hat, y_lik = self.gaussian_conditional(scaled_y, scales_hat, means=means_hat)
        scaled_y_hat = self.inv_gain_unit(y_hat, n, l)
        return {"x_hat": self.g_s(scaled_y_hat), "likelihoods": {"y": y_lik, "z": z_lik}}


class FlowCompressor(_GainedHyperprior):
    def __init__(self, n=6, in_ch=19, out_ch=5, N=128, bias=False):
        super().__init__(n, in_ch, out_ch, N, bias, zero_last=True)


class ResidualCompressor(_GainedHyperprior):
    def __init__(self, n=6, in_ch=3, N=128, bias=False):
        super().__init__(n, in_ch, in_ch, N, bias, zero_last=False)


# ------------------------------------------------------------------ BidirFlowRef (b_model/b_model.py)
class BidirFlowRef(nn.Module):
    def __init__(self, n=6, N=128):
        super().__init__()
        self.flow_predictor = UNet(6, 4, 5)
        self.Mask = UNet(16, 2, 4)
        self.flow_compressor = FlowCompressor(n=n, in_ch=19, out_ch=4, N=N, bias=False)
        self.residual_compressor = ResidualCompressor(n=n, in_ch=3, N=N, bias=False)

    backwarp = staticmethod(warp.backwarp_flex)

    def process(self, x0, x1, t=0.5):
        """b_model.py:35-45: linear-motion flows to the middle frame + the two warps."""
        x = torch.cat((x0, x1), 1)
        flow = self.flow_predictor(x)
        f01, f10 = flow[:, :2], flow[:, 2:4]
        ft0 = -(1 - t) * t * f01 + t * t * f10
        ft1 = (1 - t) * (1 - t) * f01 - t * (1 - t) * f10
        xt1, xt2 = self.backwarp(x0, ft0), self.backwarp(x1, ft1)
        return ft0, ft1, torch.cat((ft0, ft1, x, xt1, xt2), 1)

    def forward(self, x_before, x_current, x_after, n=None, l=1, train=False):
        """b_model.py:49-96."""
        x = torch.cat((x_before, x_after), 1)
        _, _, H, W = x_current.shape
        num_pixels = H * W
        mv_before, mv_after, x_conc = self.process(x_before, x_after)
        flow_result = self.flow_compressor(torch.cat((x_conc, x_current), 1), n, l, train)
        flow_hat = flow_result["x_hat"]
        mv_b = mv_before + flow_hat[:, :2]
        mv_a = mv_after + flow_hat[:, 2:4]
        x_b, x_a = self.backwarp(x_before, mv_b), self.backwarp(x_after, mv_a)
        logits = self.Mask(torch.cat((mv_b, mv_a, x, x_b, x_a), 1))
        x_comp, residual = warp.blend_residual_flex(logits, x_b, x_a, x_current)
        residual_result = self.residual_compressor(residual, n, l, train)
        x_hat = x_comp + residual_result["x_hat"]
        size = lambda res: sum(torch.log(v).sum(dim=(1, 2, 3)) / (-math.log(2)) for v in res["likelihoods"].values())
        size_flow, size_res = size(flow_result), size(residual_result)
        return {"x_hat": x_hat, "size": size_flow + size_res, "rate": size_flow / num_pixels + size_res / num_pixels}
