"""Host-side mirror of the LHBDC B-frame codec interface (reference: ``LHBDC/model/m.py`` ``Model``,
``LHBDC/model/flow.py`` ``Network``, ``LHBDC/model/layers.py`` ``MVCompressor`` / ``ResidualCompressor`` / ``Mask``;
entry points ``LHBDC/encode_B.py:71-105``).  Same class / attribute names, call signatures and state-dict keys,
so ``compression_<lambda>.pth`` checkpoints load unchanged; the hot path (both backward warps + flow glue,
blend/residual, every GDN/IGDN, both entropy models and the bit sums) runs in the sm_100a kernels.

Convolutions (SPyNet, mask U-Net, analysis/synthesis transforms) stay on cuDNN through torch: SURVEY.md 8
scopes them out.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import modules as M
from . import ops


def reflect_pad64(im):
    """``Model.pad`` (m.py:101-108): bottom/right reflection padding to a multiple of 64."""
    H, W = im.shape[-2:]
    return F.pad(im, (0, (64 - W % 64) % 64, 0, (64 - H % 64) % 64), mode="reflect")


class _Preprocess(nn.Module):
    def forward(self, t):  # flow.py:39-44 (channel swap + ImageNet statistics)
        return torch.cat([(t[:, 2:3] - 0.485) / 0.229, (t[:, 1:2] - 0.456) / 0.224, (t[:, 0:1] - 0.406) / 0.225], 1)


class _Basic(nn.Module):
    def __init__(self):
        super().__init__()
        widths = (8, 32, 64, 32, 16, 2)
        seq = []
        for a, b in zip(widths[:-1], widths[1:]):
            seq += [nn.Conv2d(a, b, kernel_size=7, stride=1, padding=3), nn.ReLU(inplace=False)]
        self.netBasic = nn.Sequential(*seq[:-1])

    def forward(self, t):
        return self.netBasic(t)


class Network(nn.Module):
    """SPyNet (flow.py:28-101); its per-level ``backwarp`` (flow.py:98) is the K-WARP kernel."""

    def __init__(self):
        super().__init__()
        self.netPreprocess = _Preprocess()
        self.netBasic = nn.ModuleList(_Basic() for _ in range(6))

    def forward(self, tenFirst, tenSecond):
        # flow.py:78-101.  Preprocess + avg-pool pyramid: one launch per image (K-SPYPYR); per level the x2 flow
        # upsample, replicate pad, warp and 8-channel concat: one launch (K-SPYLEVEL).  Convolutions stay on cuDNN.
        pyr1, pyr2 = ops.spynet_pyramid(tenFirst), ops.spynet_pyramid(tenSecond)
        flow = None  # the reference starts from new_zeros([N, 2, h // 2, w // 2]) (flow.py:90)
        for lvl, (a, b) in enumerate(zip(pyr1, pyr2)):
            feat = ops.spynet_level(a, b, flow)
            flow = self.netBasic[lvl](feat) + feat[:, 6:8]
        return flow


class _Compressor(M.MeanScaleHyperprior):
    """layers.py:43-117 / 119-190 (MVCompressor and ResidualCompressor differ only in the I/O channel count)."""

    def __init__(self, io_ch, N=128):
        super().__init__(N=N, M=N)
        rs, rb, ru = M.ResidualBlockWithStride, M.ResidualBlock, M.ResidualBlockUpsample
        self.g_a = nn.Sequential(rs(io_ch, N, stride=2), rb(N, N), rs(N, N, stride=2), rb(N, N),
                                 rs(N, N, stride=2), rb(N, N), M.conv3x3(N, N, stride=2))
        act = lambda: nn.LeakyReLU(inplace=True)
        self.h_a = nn.Sequential(M.conv3x3(N, N), act(), M.conv3x3(N, N), act(), M.conv3x3(N, N, stride=2), act(),
                                 M.conv3x3(N, N), act(), M.conv3x3(N, N, stride=2))
        self.h_s = nn.Sequential(M.conv3x3(N, N), act(), M.subpel_conv3x3(N, N, 2), act(),
                                 M.conv3x3(N, N * 3 // 2), act(), M.subpel_conv3x3(N * 3 // 2, N * 3 // 2, 2), act(),
                                 M.conv3x3(N * 3 // 2, N * 2))
        self.g_s = nn.Sequential(rb(N, N), ru(N, N, 2), rb(N, N), ru(N, N, 2), rb(N, N), ru(N, N, 2), rb(N, N),
                                 M.subpel_conv3x3(N, io_ch, 2))


class MVCompressor(_Compressor):
    def __init__(self, N=128, **kwargs):
        super().__init__(4, N)


class ResidualCompressor(_Compressor):
    def __init__(self, N=128, **kwargs):
        super().__init__(3, N)


class Mask(nn.Module):
    """layers.py:193-249."""

    def __init__(self, ch=32):
        super().__init__()
        c = lambda i, o, k: nn.Conv2d(i, o, kernel_size=k, stride=1, padding=k // 2)
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.conv1, self.conv2, self.conv3 = c(6, ch, 5), c(ch, ch * 2, 5), c(ch * 2, ch * 4, 3)
        self.bottleneck = c(ch * 4, ch * 4, 3)
        self.deconv1, self.deconv2, self.deconv3 = c(ch * 8, ch * 4, 3), c(ch * 6, ch * 2, 5), c(ch * 3, ch, 5)
        self.conv4 = c(ch, 1, 5)

    def forward(self, x):
        up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
        s1 = F.relu(self.conv1(x))
        s2 = F.relu(self.conv2(self.pool(s1)))
        s3 = F.relu(self.conv3(self.pool(s2)))
        t = F.relu(self.bottleneck(self.pool(s3)))
        t = F.relu(self.deconv1(torch.cat([up(t), s3], dim=1)))
        t = F.relu(self.deconv2(torch.cat([up(t), s2], dim=1)))
        t = F.relu(self.deconv3(torch.cat([up(t), s1], dim=1)))
        return torch.sigmoid(self.conv4(t))


class Model(nn.Module):
    """LHBDC bidirectional B-frame codec (m.py:19-126)."""

    def __init__(self):
        super().__init__()
        self.FlowNet = Network()
        self.mv_compressor = MVCompressor()
        self.residual_compressor = ResidualCompressor()
        self.masknet = Mask()
        self.upsample_flow = nn.Upsample(scale_factor=4, mode="bilinear")

    def pad(self, im):
        return reflect_pad64(im)

    def backwarp(self, tenInput, tenFlow):
        """m.py:111-126 -- K-WARP (no per-call CPU grid build / 16.7 MB H2D)."""
        return ops.backwarp(tenInput, tenFlow, "lhbdc")

    # -- motion estimation front end (m.py:38-53): SPyNet x4, pooling, padding, flow difference
    def motion(self, x_before, x_current, x_after):
        flow_ba = F.avg_pool2d(self.FlowNet(x_before, x_after) / 2., 4)
        flow_ab = F.avg_pool2d(self.FlowNet(x_after, x_before) / 2., 4)
        hh, ww = flow_ab.shape[-2:]
        flow_ba, flow_ab = self.pad(flow_ba), self.pad(flow_ab)
        flow_cb = self.pad(F.avg_pool2d(self.FlowNet(x_current, x_before), 4))
        flow_ca = self.pad(F.avg_pool2d(self.FlowNet(x_current, x_after), 4))
        return torch.cat([flow_cb - flow_ab, flow_ca - flow_ba], dim=1), flow_ab, flow_ba, hh, ww

    def forward_device(self, x_before, x_current, x_after, return_parts=False):
        """The whole B-frame step with no host synchronisation (CUDA-graph capturable).
        Returns (x_hat, bits[N] float64 = size_flow + size_residual, parts dict).  ``return_parts=True`` keeps every
        intermediate of m.py:38-93 plus both compressors' symbols / CDF indexes in ``parts`` (the acceptance tests
        compare them stage by stage with the reference path)."""
        diff, flow_ab, flow_ba, hh, ww = self.motion(x_before, x_current, x_after)
        mv = self.mv_compressor.forward_bits(diff, want_symbols=return_parts)
        flow_hat, fb_y, fb_z = mv[:3]
        H, W = x_current.shape[-2:]
        if (hh * 4, ww * 4) != (H, W):
            raise RuntimeError(f"LHBDC needs H, W divisible by 4 (got {H}x{W})")
        warped = ops.warp2_lhbdc(x_before, x_after, flow_hat, flow_ab, flow_ba)  # [N,6,H,W] = cat(fw, bw)
        mask = self.masknet(warped)
        pred, residual, _ = ops.blend_residual("mask", mask, warped[:, 0:3], warped[:, 3:6], x_current)
        rs = self.residual_compressor.forward_bits(residual, want_symbols=return_parts)
        res_hat, rb_y, rb_z = rs[:3]
        x_hat = res_hat + pred
        bits_flow, bits_res = fb_y + fb_z, rb_y + rb_z
        parts = {"bits_flow": bits_flow, "bits_residual": bits_res}
        if return_parts:
            parts.update(diff_flow=diff, flow_ab=flow_ab, flow_ba=flow_ba, flow_hat=flow_hat, warped=warped, mask=mask,
                         pred=pred, residual=residual, res_hat=res_hat, mv=mv[3], res=rs[3],
                         bits_flow_y=fb_y, bits_flow_z=fb_z, bits_res_y=rb_y, bits_res_z=rb_z)
        return x_hat, bits_flow + bits_res, parts

    def forward(self, x_before, x_current, x_after, train):
        """m.py:32-98.  ``rate`` keeps the reference's definition ((rate_flow + rate_residual)/2 over the padded
        pixel count); ``size`` is the python float the reference obtains through ``.item()``."""
        if train:
            raise NotImplementedError("b200vc mirrors the inference path (train=False) only")
        N, _, H, W = x_current.size()
        x_hat, bits, _ = self.forward_device(x_before, x_current, x_after)
        total = bits.sum()
        rate = (total / (N * H * W) / 2.0).float()
        return x_hat, rate, total.item()


def _anchor_flows(model, x_before, x_after):
    """The scripts' version of the anchor-to-anchor prior (encode_B.py:74-79 == decode_B.py:65-70), quirk B.1
    included: flow_ba is overwritten by pad(flow_ab), so both priors are pad(flow_ab)."""
    flow_ab = F.avg_pool2d(model.FlowNet(x_after, x_before) / 2., 4)
    hh, ww = flow_ab.shape[-2:]
    flow_ba = model.pad(flow_ab)
    flow_ab = model.pad(flow_ba)
    return flow_ab, flow_ba, hh, ww


def encode_B(model, x_after, x_current, x_before):
    """``encode_B`` of LHBDC/encode_B.py:71-105: returns (mv_bits, res_bits), each
    {"strings": [y_strings, z_strings], "shape": z spatial size} with real rANS byte strings."""
    flow_ab, flow_ba, hh, ww = _anchor_flows(model, x_before, x_after)
    flow_cb = model.pad(F.avg_pool2d(model.FlowNet(x_current, x_before), 4))
    flow_ca = model.pad(F.avg_pool2d(model.FlowNet(x_current, x_after), 4))
    diff_flow = torch.cat([flow_cb - flow_ab, flow_ca - flow_ba], dim=1)
    flow_hat, _, _ = model.mv_compressor.forward_bits(diff_flow)
    mv_bits = model.mv_compressor.compress(diff_flow)
    warped = ops.warp2_lhbdc(x_before, x_after, flow_hat, flow_ab, flow_ba)
    mask = model.masknet(warped)
    _, res, _ = ops.blend_residual("mask", mask, warped[:, 0:3], warped[:, 3:6], x_current, want_pred=False)
    return mv_bits, model.residual_compressor.compress(res)


def decode_B(x_before, x_after, model, string_flow, string_res, shape_flow, shape_res):
    """``decode_B`` of LHBDC/decode_B.py:63-86: the decoded (padded) frame."""
    flow_ab, flow_ba, hh, ww = _anchor_flows(model, x_before, x_after)
    flow_hat = model.mv_compressor.decompress(string_flow, shape_flow)["x_hat"]
    warped = ops.warp2_lhbdc(x_before, x_after, flow_hat, flow_ab, flow_ba)
    mask = model.masknet(warped)
    pred, _, _ = ops.blend_residual("mask", mask, warped[:, 0:3], warped[:, 3:6], x_before, want_res=False)
    return model.residual_compressor.decompress(string_res, shape_res)["x_hat"] + pred


def encode_B_symbols(model, x_after, x_current, x_before):
    """Tensor half of ``encode_B`` (LHBDC/encode_B.py:71-105): everything up to the rANS calls, i.e. the
    int32 symbols + CDF indexes of both latents.  Keeps the script's quirk: both anchor flows are
    pad(flow_ab)."""
    flow_ab = F.avg_pool2d(model.FlowNet(x_after, x_before) / 2., 4)
    hh, ww = flow_ab.shape[-2:]
    flow_ba = model.pad(flow_ab)
    flow_ab = model.pad(flow_ba)
    flow_cb = model.pad(F.avg_pool2d(model.FlowNet(x_current, x_before), 4))
    flow_ca = model.pad(F.avg_pool2d(model.FlowNet(x_current, x_after), 4))
    diff_flow = torch.cat([flow_cb - flow_ab, flow_ca - flow_ba], dim=1)
    flow_hat, _, _ = model.mv_compressor.forward_bits(diff_flow)
    mv_syms = model.mv_compressor.symbols(diff_flow)
    warped = ops.warp2_lhbdc(x_before, x_after, flow_hat, flow_ab, flow_ba)
    mask = model.masknet(warped)
    _, res, _ = ops.blend_residual("mask", mask, warped[:, 0:3], warped[:, 3:6], x_current, want_pred=False)
    return mv_syms, model.residual_compressor.symbols(res)
