"""TEST INFRASTRUCTURE (oracle).  Torch restatement of the LHBDC B-frame codec (the reference's
``LHBDC/model/{m,flow,layers}.py``) on top of ``oracle/cai.py``; device-agnostic, same module tree and
state-dict keys, same parameter construction order (so ``torch.manual_seed(s); Model()`` yields the same
weights as the reference built through ``oracle/shim.py`` -- asserted by ``oracle/make_golden.py``).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import cai, warp


# ------------------------------------------------------------------ SPyNet (LHBDC/model/flow.py:28-101)
class _Preprocess(nn.Module):
    """flow.py:36-45: ImageNet normalisation with channels 0<->2 swapped."""

    def forward(self, x):
        b = (x[:, 0:1] - 0.406) / 0.225
        g = (x[:, 1:2] - 0.456) / 0.224
        r = (x[:, 2:3] - 0.485) / 0.229
        return torch.cat([r, g, b], 1)


class _Basic(nn.Module):
    """flow.py:48-67: five 7x7 convs 8->32->64->32->16->2."""

    def __init__(self):
        super().__init__()
        chans = [8, 32, 64, 32, 16, 2]
        layers = []
        for i in range(5):
            layers.append(nn.Conv2d(chans[i], chans[i + 1], kernel_size=7, stride=1, padding=3))
            if i < 4:
                layers.append(nn.ReLU(inplace=False))
        self.netBasic = nn.Sequential(*layers)

    def forward(self, x):
        return self.netBasic(x)


class Network(nn.Module):
    def __init__(self, backwarp=warp.backwarp_lhbdc):
        super().__init__()
        self.netPreprocess = _Preprocess()
        self.netBasic = nn.ModuleList([_Basic() for _ in range(6)])
        self._backwarp = backwarp

    def forward(self, first, second):
        """flow.py:77-101."""
        first = [self.netPreprocess(first)]
        second = [self.netPreprocess(second)]
        for _ in range(5):
            if first[0].shape[2] > 32 or first[0].shape[3] > 32:
                first.insert(0, F.avg_pool2d(first[0], kernel_size=2, stride=2, count_include_pad=False))
                second.insert(0, F.avg_pool2d(second[0], kernel_size=2, stride=2, count_include_pad=False))
        flow = first[0].new_zeros(
            [first[0].shape[0], 2, int(math.floor(first[0].shape[2] / 2.0)), int(math.floor(first[0].shape[3] / 2.0))]
        )
        for lvl in range(len(first)):
            up = F.interpolate(flow, scale_factor=2, mode="bilinear", align_corners=True) * 2.0
            if up.shape[2] != first[lvl].shape[2]:
                up = F.pad(up, [0, 0, 0, 1], mode="replicate")
            if up.shape[3] != first[lvl].shape[3]:
                up = F.pad(up, [0, 1, 0, 0], mode="replicate")
            flow = self.netBasic[lvl](torch.cat([first[lvl], self._backwarp(second[lvl], up), up], 1)) + up
        return flow


# ------------------------------------------------------- hyperprior compressors (LHBDC/model/layers.py)
def _analysis(in_ch, N):
    return nn.Sequential(
        cai.ResidualBlockWithStride(in_ch, N, stride=2), cai.ResidualBlock(N, N),
        cai.ResidualBlockWithStride(N, N, stride=2), cai.ResidualBlock(N, N),
        cai.ResidualBlockWithStride(N, N, stride=2), cai.ResidualBlock(N, N),
        cai.conv3x3(N, N, stride=2),
    )


def _hyper_analysis(N):
    return nn.Sequential(
        cai.conv3x3(N, N), nn.LeakyReLU(inplace=True), cai.conv3x3(N, N), nn.LeakyReLU(inplace=True),
        cai.conv3x3(N, N, stride=2), nn.LeakyReLU(inplace=True), cai.conv3x3(N, N), nn.LeakyReLU(inplace=True),
        cai.conv3x3(N, N, stride=2),
    )


def _hyper_synthesis(N):
    return nn.Sequential(
        cai.conv3x3(N, N), nn.LeakyReLU(inplace=True), cai.subpel_conv3x3(N, N, 2), nn.LeakyReLU(inplace=True),
        cai.conv3x3(N, N * 3 // 2), nn.LeakyReLU(inplace=True),
        cai.subpel_conv3x3(N * 3 // 2, N * 3 // 2, 2), nn.LeakyReLU(inplace=True),
        cai.conv3x3(N * 3 // 2, N * 2),
    )


def _synthesis(N, out_ch):
    return nn.Sequential(
        cai.ResidualBlock(N, N), cai.ResidualBlockUpsample(N, N, 2), cai.ResidualBlock(N, N),
        cai.ResidualBlockUpsample(N, N, 2), cai.ResidualBlock(N, N), cai.ResidualBlockUpsample(N, N, 2),
        cai.ResidualBlock(N, N), cai.subpel_conv3x3(N, out_ch, 2),
    )


class _Hyperprior(cai.MeanScaleHyperprior):
    """layers.py:43-117 / 119-190: creation order g_a, h_a, h_s, g_s."""

    def __init__(self, ch, N=128):
        super().__init__(N=N, M=N)
        self.g_a = _analysis(ch, N)
        self.h_a = _hyper_analysis(N)
        self.h_s = _hyper_synthesis(N)
        self.g_s = _synthesis(N, ch)

    def symbols(self, x):
        """The tensor half of ``compress`` (layers.py:93-104): what the rANS coder would be fed.
        EB round-trip is replaced by its fixed point ``round(z - m) + m``."""
        y = self.g_a(x)
        z = self.h_a(y)
        med = self.entropy_bottleneck._get_medians().reshape(1, -1, 1, 1)
        z_sym = torch.round(z - med).int()
        z_hat = z_sym.float() + med
        scales_hat, means_hat = self.h_s(z_hat).chunk(2, 1)
        idx = self.gaussian_conditional.build_indexes(scales_hat)
        y_sym = self.gaussian_conditional.quantize(y, "symbols", means_hat)
        return {"y_symbols": y_sym, "y_indexes": idx, "z_symbols": z_sym, "shape": z.size()[-2:]}


class MVCompressor(_Hyperprior):
    def __init__(self, N=128):
        super().__init__(4, N)


class ResidualCompressor(_Hyperprior):
    def __init__(self, N=128):
        super().__init__(3, N)


def _same_conv(i, o, k):
    return nn.Conv2d(i, o, kernel_size=k, stride=1, padding=k // 2)


class Mask(nn.Module):
    """layers.py:193-249: 3-level U-Net producing a 1-channel sigmoid mask."""

    def __init__(self, ch=32):
        super().__init__()
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.conv1 = _same_conv(6, ch, 5)
        self.conv2 = _same_conv(ch, ch * 2, 5)
        self.conv3 = _same_conv(ch * 2, ch * 4, 3)
        self.bottleneck = _same_conv(ch * 4, ch * 4, 3)
        self.deconv1 = _same_conv(ch * 8, ch * 4, 3)
        self.deconv2 = _same_conv(ch * 4 + ch * 2, ch * 2, 5)
        self.deconv3 = _same_conv(ch * 2 + ch, ch, 5)
        self.conv4 = _same_conv(ch, 1, 5)

    def forward(self, x):
        up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
        c1 = F.relu(self.conv1(x))
        c2 = F.relu(self.conv2(self.pool(c1)))
        c3 = F.relu(self.conv3(self.pool(c2)))
        x = F.relu(self.bottleneck(self.pool(c3)))
        x = F.relu(self.deconv1(torch.cat([up(x), c3], dim=1)))
        x = F.relu(self.deconv2(torch.cat([up(x), c2], dim=1)))
        x = F.relu(self.deconv3(torch.cat([up(x), c1], dim=1)))
        return torch.sigmoid(self.conv4(x))


# ------------------------------------------------------------------------ Model (LHBDC/model/m.py)
class Model(nn.Module):
    def __init__(self):
        super().__init__()
        self.FlowNet = Network()
        self.mv_compressor = MVCompressor()
        self.residual_compressor = ResidualCompressor()
        self.masknet = Mask()
        self.upsample_flow = nn.Upsample(scale_factor=4, mode="bilinear")

    pad = staticmethod(warp.reflect_pad64)
    backwarp = staticmethod(warp.backwarp_lhbdc)

    def motion(self, x_before, x_current, x_after):
        """m.py:38-53: four SPyNet runs, /2 on the anchor-to-anchor pair, 4x4 mean pool, pad, difference."""
        flow_ba = F.avg_pool2d(self.FlowNet(x_before, x_after) / 2.0, 4)
        flow_ab = F.avg_pool2d(self.FlowNet(x_after, x_before) / 2.0, 4)
        hh, ww = flow_ab.shape[2], flow_ab.shape[3]
        flow_ba, flow_ab = self.pad(flow_ba), self.pad(flow_ab)
        flow_cb = self.pad(F.avg_pool2d(self.FlowNet(x_current, x_before), 4))
        flow_ca = self.pad(F.avg_pool2d(self.FlowNet(x_current, x_after), 4))
        diff = torch.cat([flow_cb - flow_ab, flow_ca - flow_ba], dim=1)
        return diff, flow_ab, flow_ba, hh, ww

    def forward(self, x_before, x_current, x_after, train, return_parts=False):
        N, _, H, W = x_current.size()
        num_pixels = N * H * W
        diff, flow_ab, flow_ba, hh, ww = self.motion(x_before, x_current, x_after)
        flow_result = self.mv_compressor(diff)
        flow_cb_hat, flow_ca_hat = warp.lhbdc_flow_glue(flow_result["x_hat"], flow_ab, flow_ba, hh, ww)
        fw, bw = self.backwarp(x_before, flow_cb_hat), self.backwarp(x_after, flow_ca_hat)
        mask1 = self.masknet(torch.cat([fw, bw], dim=1))
        pred, residual = warp.blend_residual_lhbdc(mask1, fw, bw, x_current)
        residual_result = self.residual_compressor(residual)
        x_hat = residual_result["x_hat"] + pred

        def rate(res):
            return sum(torch.log(l).sum() / (-math.log(2) * num_pixels) for l in res["likelihoods"].values())

        def size(res):
            return sum(torch.log(l).sum() / (-math.log(2)) for l in res["likelihoods"].values())

        r = (rate(flow_result) + rate(residual_result)) / 2.0
        if train:
            return x_hat, r
        s = size(flow_result).item() + size(residual_result).item()
        if return_parts:
            parts = dict(
                diff_flow=diff, flow_cb_hat=flow_cb_hat, flow_ca_hat=flow_ca_hat, fw=fw, bw=bw, mask=mask1,
                pred=pred, residual=residual, flow_result=flow_result, residual_result=residual_result,
                size64=sum(warp.bits_fp64(l) for l in flow_result["likelihoods"].values()).item()
                + sum(warp.bits_fp64(l) for l in residual_result["likelihoods"].values()).item(),
            )
            return x_hat, r, s, parts
        return x_hat, r, s


def encode_B_symbols(model, x_after, x_current, x_before):
    """Tensor half of ``encode_B`` (LHBDC/encode_B.py:71-105) with rANS replaced by symbols+indexes
    (SURVEY 8d config 1).  Keeps quirk B.1: both anchor flows become pad(flow_ab)."""
    flow_ab = F.avg_pool2d(model.FlowNet(x_after, x_before) / 2.0, 4)
    hh, ww = flow_ab.shape[2], flow_ab.shape[3]
    flow_ba = model.pad(flow_ab)
    flow_ab = model.pad(flow_ba)
    flow_cb = model.pad(F.avg_pool2d(model.FlowNet(x_current, x_before), 4))
    flow_ca = model.pad(F.avg_pool2d(model.FlowNet(x_current, x_after), 4))
    diff = torch.cat([flow_cb - flow_ab, flow_ca - flow_ba], dim=1)
    flow_result = model.mv_compressor(diff)
    cb, ca = warp.lhbdc_flow_glue(flow_result["x_hat"], flow_ab, flow_ba, hh, ww)
    mv_syms = model.mv_compressor.symbols(diff)
    fw, bw = model.backwarp(x_before, cb), model.backwarp(x_after, ca)
    mask1 = model.masknet(torch.cat([fw, bw], dim=1))
    _, res = warp.blend_residual_lhbdc(mask1, fw, bw, x_current)
    return mv_syms, model.residual_compressor.symbols(res)
