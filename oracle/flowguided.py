"""TEST INFRASTRUCTURE (oracle).  Torch restatement of the ICIP2024 flow-guided deformable B-frame codec:
``ICIP2024/src/model/m.py`` (``FlowGuidedB``), ``helpers.py`` (feature pyramid, flow net, temporal conditioners,
``OffsetDiversity``, reconstructor), ``compression_bottlenecks.py`` (``Offset_ELIC`` / ``Res_ELIC``: gain-modulated
ELIC-style entropy models with the checkerboard x channel-group context loop), ``elic.py:69-84``
(``ResidualBottleneckBlock``), ``layers.py`` (``CheckerboardContext``); and of the evaluation loop around it:
``ICIP2024/src/opt_helpers.py:23-51`` (down-ratio search), ``utils.py:154-240`` (reference selection, coding order,
temporal scales), ``test.py:37-93`` (per-sequence loop).

Same module tree and state-dict keys as the reference (checked key for key by ``oracle/make_golden_flowguided.py``,
which also asserts that every function here equals the reference's own code bit for bit on the CPU and stores
``tests/golden/flowguided_reference.npz``).  Device-agnostic; the hot operators are the plain torch / torchvision
calls the reference makes (``grid_sample``, ``torchvision.ops.deform_conv2d``, elementwise chains).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import cai, icip
from .warp import warp_ac1

DOWN_RATIOS = (1, 2, 4, 8, 16)            # opt_helpers.py:43
LEVELS = 5                                # compression_bottlenecks.py:204


def _conv(i, o, k=5, s=2):
    return nn.Conv2d(i, o, kernel_size=k, stride=s, padding=k // 2)


def _deconv(i, o, k=5, s=2):
    return nn.ConvTranspose2d(i, o, kernel_size=k, stride=s, output_padding=s - 1, padding=k // 2)


class ResidualBottleneckBlock(nn.Module):
    """elic.py:69-84: 1x1 -> ReLU -> 3x3 -> ReLU -> 1x1, plus identity."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.BottleneckBlock = nn.Sequential(
            nn.Conv2d(in_ch, out_ch, 1), nn.ReLU(inplace=True), nn.Conv2d(out_ch, out_ch, 3, padding=1),
            nn.ReLU(inplace=True), nn.Conv2d(out_ch, out_ch, 1))

    def forward(self, x):
        return self.BottleneckBlock(x) + x


def _stage(head, ch, blocks=3, tail=None):
    """head -> `blocks` bottleneck blocks -> tail().  ``tail`` is a factory so that parameters are created in the
    reference's order (same seed => same weights as the reference classes)."""
    mods = ([head] if head is not None else []) + [ResidualBottleneckBlock(ch, ch) for _ in range(blocks)]
    return nn.Sequential(*(mods + ([tail()] if tail is not None else [])))


class CheckerboardContext(nn.Conv2d):
    """layers.py:6-29: 5x5 conv whose weight is multiplied (in place, every call) by the anchor checkerboard."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.register_buffer("mask", torch.zeros_like(self.weight.data))
        self.mask[:, :, 0::2, 1::2] = 1
        self.mask[:, :, 1::2, 0::2] = 1

    def forward(self, x):
        self.weight.data *= self.mask
        return super().forward(x)


class MS_Feature(nn.Module):
    """helpers.py:72-108: three-level feature pyramid 64 / 96 / 128 channels at 1/2, 1/4, 1/8."""

    def __init__(self):
        super().__init__()
        self.layer1 = _stage(_conv(3, 64, 3, 2), 64)
        self.layer2 = _stage(_conv(64, 96, 3, 2), 96)
        self.layer3 = _stage(_conv(96, 128, 3, 2), 128)

    def forward(self, x):
        l1 = self.layer1(x)
        l2 = self.layer2(l1)
        return l1, l2, self.layer3(l2)


class FlowNET(nn.Module):
    """helpers.py:111-166: 4-level U-net on the concatenated reference pair -> 4 flow channels (2 -> 1, 1 -> 2)."""

    def __init__(self):
        super().__init__()
        w = (32, 64, 128, 192)
        self.down0 = _stage(_conv(6, w[0], 3, 2), w[0], 2)
        self.down1 = _stage(_conv(w[0], w[1], 3, 2), w[1], 2)
        self.down2 = _stage(_conv(w[1], w[2], 3, 2), w[2], 2)
        self.down3 = _stage(_conv(w[2], w[3], 3, 2), w[3], 2)
        self.up0 = _stage(None, w[3], 2, lambda: cai.subpel_conv3x3(w[3], w[2], 2))
        self.up1 = _stage(_conv(2 * w[2], w[2], 1, 1), w[2], 2, lambda: cai.subpel_conv3x3(w[2], w[1], 2))
        self.up2 = _stage(_conv(2 * w[1], w[1], 1, 1), w[1], 2, lambda: cai.subpel_conv3x3(w[1], w[0], 2))
        self.up3 = _stage(_conv(2 * w[0], w[0], 1, 1), w[0], 2, lambda: cai.subpel_conv3x3(w[0], 4, 2))

    def forward(self, inp):
        s0 = self.down0(inp)
        s1 = self.down1(s0)
        s2 = self.down2(s1)
        x = self.up0(self.down3(s2))
        x = self.up1(torch.cat((x, s2), 1))
        x = self.up2(torch.cat((x, s1), 1))
        return self.up3(torch.cat((x, s0), 1))


class TemporalEnc(nn.Module):
    """helpers.py:169-222 (OffsetTemproalEnc: widths x4; ResidualTemproalEnc: x1): pyramid -> 1/16 condition."""

    def __init__(self, mult, N=128, M=128):
        super().__init__()
        self.g_a1 = _stage(_conv(64 * mult, N), N)
        self.g_a2 = _stage(_conv(N + 96 * mult, N), N)
        self.g_a3 = _stage(_conv(N + 128 * mult, M), M)

    def forward(self, l1, l2, l3):
        y = self.g_a1(l1)
        y = self.g_a2(torch.cat([y, l2], dim=1))
        return self.g_a3(torch.cat([y, l3], dim=1))


class Reconstuctor(nn.Module):
    """helpers.py:226-261 (the reference's spelling is kept: it is a state-dict-visible class only by attribute)."""

    def __init__(self):
        super().__init__()
        self.layer3 = _stage(None, 128, 3, lambda: cai.subpel_conv3x3(128, 128, 2))
        self.layer2 = _stage(_conv(128 + 96, 96, 1, 1), 96, 3, lambda: cai.subpel_conv3x3(96, 96, 2))
        self.layer1 = _stage(_conv(96 + 64, 64, 1, 1), 64, 3, lambda: cai.subpel_conv3x3(64, 3, 2))

    def forward(self, c1, c2, c3):
        l3 = self.layer3(c3)
        l2 = self.layer2(torch.cat([c2, l3], dim=1))
        return self.layer1(torch.cat([c1, l2], dim=1))


class OffsetDiversity(nn.Module):
    """helpers.py:35-69 on torchvision's operator (through oracle.icip.offset_diversity_forward)."""

    def __init__(self, in_channel, magnitude):
        super().__init__()
        from torchvision.ops import DeformConv2d
        self.in_channel, self.magnitude = in_channel, magnitude
        self.fusion = DeformConv2d(in_channel * 2, in_channel, kernel_size=3, padding=1, groups=2 * 8)

    def forward(self, x1, offset1, flow1, x2, offset2, flow2):
        from torchvision.ops import deform_conv2d
        o1, m1 = icip.offset_diversity_prep(offset1, flow1, self.magnitude)
        o2, m2 = icip.offset_diversity_prep(offset2, flow2, self.magnitude)
        return deform_conv2d(torch.cat((x1, x2), 1), torch.cat((o1, o2), 1), self.fusion.weight, self.fusion.bias,
                             padding=(1, 1), mask=torch.cat((m1, m2), 1))


class _GainedELIC(cai.JointAutoregressiveHierarchicalPriors):
    """compression_bottlenecks.py:72-311 (Offset_ELIC) / :313-551 (Res_ELIC).  ``mult`` = how many feature maps per
    pyramid level enter the analysis (5 / 2), ``dmult`` = how many re-enter the synthesis (4 / 1), ``outs`` = output
    channels of the three heads (27*8*2 offsets+masks each / the residual feature widths)."""

    def __init__(self, mult, dmult, outs, names, N=128, M=128):
        super().__init__(N, M)
        self.names = names
        self.g_a1 = _stage(_conv(64 * mult, N), N)
        self.g_a2 = _stage(_conv(N + 96 * mult, N), N)
        self.g_a3 = _stage(_conv(N + 128 * mult, M), M)
        self.g_s3 = _stage(None, M, 3, lambda: _deconv(M, N))
        self.g_o3 = _stage(_conv(N + 128 * dmult, N, 3, 1), N, 3, lambda: _conv(N, outs[2], 3, 1))
        self.g_s2 = _stage(_conv(N + 128 * dmult, N, 1, 1), N, 3, lambda: _deconv(N, N))
        self.g_o2 = _stage(_conv(N + 96 * dmult, N, 3, 1), N, 3, lambda: _conv(N, outs[1], 3, 1))
        self.g_s1 = _stage(_conv(N + 96 * dmult, N, 1, 1), N, 3, lambda: _deconv(N, N))
        self.g_o1 = _stage(_conv(N + 64 * dmult, N, 3, 1), N, 3, lambda: _conv(N, outs[0], 3, 1))
        relu = lambda: nn.ReLU(inplace=True)
        self.h_a = nn.Sequential(_conv(M, N, 3, 1), relu(), _conv(N, N), relu(), _conv(N, N))
        self.h_s = nn.Sequential(_deconv(N, M), relu(), _deconv(M, M), relu(), _conv(M, M, 3, 1))
        self.prior_fusion = _stage(_conv(2 * M, 2 * M, 3, 1), 2 * M, 3, lambda: _conv(2 * M, 2 * M, 3, 1))
        lrelu = lambda: nn.LeakyReLU(inplace=True)
        groups = (6, 6, 12, 24, M - 48)
        self.entropy_parameters = nn.ModuleList(
            nn.Sequential(nn.Conv2d(M * (4 if g == 0 else 6), M * 10 // 3, 1), lrelu(),
                          nn.Conv2d(M * 10 // 3, M * 8 // 3, 1), lrelu(), nn.Conv2d(M * 8 // 3, 2 * c, 1))
            for g, c in enumerate(groups))
        self.channel_context_models = nn.ModuleList(
            nn.Sequential(_conv(c, N, 5, 1), relu(), _conv(N, N, 5, 1), relu(), _conv(N, 2 * M, 5, 1))
            for c in (6, 12, 24, 48))
        self.context_prediction_models = nn.ModuleList(
            CheckerboardContext(in_channels=c, out_channels=2 * M, kernel_size=5, stride=1, padding=2) for c in groups)
        self.levels = LEVELS
        for nm, width in (("Gain", M), ("InverseGain", M), ("HyperGain", N), ("InverseHyperGain", N)):
            setattr(self, nm, nn.Parameter(torch.ones(LEVELS, width)))

    def interpolate_gain(self, s):
        """:285-311: |G[s]| at integer levels, geometric interpolation |G[up]|^(1-l) * |G[lo]|^l between them."""
        s = max(min(s, self.levels - 1), 0)
        up, lo = int(min(math.ceil(s), self.levels - 1)), int(max(math.floor(s), 0))
        out = []
        for nm in ("Gain", "HyperGain", "InverseHyperGain", "InverseGain"):
            g = getattr(self, nm)
            if up == lo:
                out.append(torch.abs(g[int(s)]))
            else:
                l = up - s
                out.append(torch.abs(g[up]) ** (1 - l) * torch.abs(g[lo]) ** l)
        return out  # gain, hypergain, invhypergain, invgain

    def analysis(self, f, fd):
        raise NotImplementedError

    def forward(self, f1, f2, f3, f1d, f2d, f3d, temp, s):
        gain, hypergain, invhypergain, invgain = self.interpolate_gain(s)
        bc = lambda v: v.unsqueeze(0).unsqueeze(2).unsqueeze(3)
        y = self.analysis((f1, f2, f3), (f1d, f2d, f3d)) * bc(gain)
        z = self.h_a(y) * bc(hypergain)
        _, z_lik = self.entropy_bottleneck(z)
        z_hat = icip.ste_round(z) * bc(invhypergain)           # quirk B.8: no median in the reconstruction path
        hyper = self.prior_fusion(torch.cat([self.h_s(z_hat), temp], dim=1))
        liks, y_hat = icip.elic_context_likelihoods(y, hyper, self.context_prediction_models,
                                                    self.channel_context_models, self.entropy_parameters,
                                                    self.gaussian_conditional, inv_gain=invgain)
        lik = {"z": z_lik}
        lik.update(liks)
        inp3 = torch.cat([self.g_s3(y_hat), f3d], dim=1)
        inp2 = torch.cat([self.g_s2(inp3), f2d], dim=1)
        inp1 = torch.cat([self.g_s1(inp2), f1d], dim=1)
        n = self.names
        return {n + "3": self.g_o3(inp3), n + "2": self.g_o2(inp2), n + "1": self.g_o1(inp1), "likelihoods": lik}


class Offset_ELIC(_GainedELIC):
    def __init__(self, N=128, M=128):
        super().__init__(5, 4, (27 * 8 * 2,) * 3, "offset", N, M)

    def analysis(self, f, fd):
        y = self.g_a1(f[0])
        y = self.g_a2(torch.cat([y, f[1]], dim=1))
        return self.g_a3(torch.cat([y, f[2]], dim=1))


class Res_ELIC(_GainedELIC):
    def __init__(self, N=128, M=128):
        super().__init__(2, 1, (64, 96, 128), "res", N, M)

    def analysis(self, f, fd):
        y = self.g_a1(torch.cat([f[0], fd[0]], dim=1))
        y = self.g_a2(torch.cat([y, f[1], fd[1]], dim=1))
        return self.g_a3(torch.cat([y, f[2], fd[2]], dim=1))


def _bits(result):
    return sum(torch.log(l).sum() / (-math.log(2)) for l in result["likelihoods"].values())


class FlowGuidedB(nn.Module):
    """m.py:31-282."""

    def __init__(self):
        super().__init__()
        self.feature_extractor = MS_Feature()
        self.flow_estimator = FlowNET()
        self.offset_temporal_conditioner = TemporalEnc(4)
        self.offset_compressor = Offset_ELIC()
        self.offset_diversity_l3 = OffsetDiversity(128, 10)
        self.offset_diversity_l2 = OffsetDiversity(96, 20)
        self.offset_diversity_l1 = OffsetDiversity(64, 40)
        self.residue_temporal_conditioner = TemporalEnc(1)
        self.residual_compressor = Res_ELIC()
        self.reconstructor = Reconstuctor()

    convert_scales = staticmethod(icip.convert_scales)
    warp = staticmethod(warp_ac1)

    @staticmethod
    def pad_flow(t):
        """m.py:51-58: zero-pad bottom / right to a multiple of 16."""
        h, w = t.shape[2], t.shape[3]
        return F.pad(t, (0, (16 - w % 16) % 16, 0, (16 - h % 16) % 16))

    def estimate_flow(self, xref1, xref2, down_ratio):
        """m.py:84-102: flow on a (2*down_ratio)-times pooled pair, up-sampled by down_ratio -> half resolution."""
        a, b = F.avg_pool2d(xref1, down_ratio * 2), F.avg_pool2d(xref2, down_ratio * 2)
        h, w = a.shape[2], a.shape[3]
        flow = self.flow_estimator(torch.cat((self.pad_flow(a), self.pad_flow(b)), dim=1))[:, :, :h, :w]
        return F.interpolate(flow, scale_factor=down_ratio, mode="bilinear", align_corners=False) * down_ratio

    def forward(self, xref1, xref2, scale1, scale2, xcur, s, down_ratio):
        B, _, H, W = xcur.shape
        scale1, scale2 = self.convert_scales(scale1, scale2, xcur)
        flow = self.estimate_flow(xref1, xref2, down_ratio)
        fref1, fref2, fcur = (self.feature_extractor(t) for t in (xref1, xref2, xcur))
        flows, wrefs = [], []
        for lvl in range(3):                                   # m.py:104-118, three pyramid levels
            f21, f12 = torch.chunk(flow, 2, dim=1)
            c1, c2 = f21 * scale1, f12 * scale2
            flows.append((c1, c2))
            wrefs.append((self.warp(fref1[lvl], c1), self.warp(fref2[lvl], c2)))
            flow = F.interpolate(flow, scale_factor=0.5, mode="bilinear", align_corners=False) * 0.5
        cond = [torch.cat((wrefs[l][0], wrefs[l][1], fref1[l], fref2[l]), dim=1) for l in range(3)]
        inp = [torch.cat((cond[l], fcur[l]), dim=1) for l in range(3)]
        off = self.offset_compressor(inp[0], inp[1], inp[2], cond[0], cond[1], cond[2],
                                     self.offset_temporal_conditioner(*cond), s)
        comp = []
        for lvl, od in ((2, self.offset_diversity_l3), (1, self.offset_diversity_l2), (0, self.offset_diversity_l1)):
            o1, o2 = torch.chunk(off["offset" + str(lvl + 1)], 2, dim=1)
            comp.append(od(fref1[lvl], o1, flows[lvl][0], fref2[lvl], o2, flows[lvl][1]))
        c3, c2, c1 = comp
        res = self.residual_compressor(fcur[0], fcur[1], fcur[2], c1, c2, c3,
                                       self.residue_temporal_conditioner(c1, c2, c3), s)
        x_hat = self.reconstructor(c1 + res["res1"], c2 + res["res2"], c3 + res["res3"])
        size_offset, size_residual = _bits(off), _bits(res)
        n = H * W * B
        return {"x_hat": x_hat, "size": size_offset + size_residual, "rate": size_offset / n + size_residual / n}


# ------------------------------------------------------------------------------------- evaluation loop
def get_best_down_ratio_prediction(model, xref1, xref2, scale1, scale2, xcur):
    """opt_helpers.py:41-51 -> (best_down_ratio, best_pred_psnr): first strictly best flow-only prediction PSNR."""
    best_psnr, best = 0, None
    for r in DOWN_RATIOS:
        pred = icip.prediction_flowonly(model, xcur, xref1, xref2, scale1, scale2, r)
        psnr = 10 * torch.log10(1.0 / torch.mean((torch.clamp(pred, 0, 1) - xcur) ** 2))
        if psnr > best_psnr:
            best_psnr, best = psnr, r
    return best, best_psnr


def code_frame(model, xref1, xref2, xcur, scale1, scale2, s):
    """test.py:60-76: down-ratio search, then the model at the winning ratio."""
    ratio, pred_psnr = get_best_down_ratio_prediction(model, xref1, xref2, scale1, scale2, xcur)
    out = model(xref1=xref1, xref2=xref2, xcur=xcur, scale1=scale1, scale2=scale2, s=s, down_ratio=ratio)
    return {"x_hat": out["x_hat"], "bits": out["size"].item(), "down_ratio": ratio, "pred_psnr": float(pred_psnr)}


def get_order_typ_list(intra_size, frame_number):
    """utils.py:190-222: hierarchical coding order inside each 16-frame period (I at multiples of ``intra_size`` and at
    the last frame), with the reference's special tails: a generic descending tail, and the literal orders it hard-codes
    for 300 and 600 frames."""
    order = [16, 8, 4, 12, 2, 14, 6, 10, 1, 15, 3, 13, 5, 11, 7, 9]
    o = [0] + [order[i % 16] + (i // 16) * 16 for i in range(frame_number - 1)]
    ff = (frame_number - 1) % intra_size
    if ff != 0:
        m = max(o[:-ff])
        o[-ff:] = [m + ff - i for i in range(ff)]
    typ = ["I" if i % intra_size == 0 else "B" for i in range(frame_number)]
    typ[-1] = "I"
    if frame_number == 300:
        o[-11:] = [299, 293, 290, 296, 289, 291, 292, 294, 295, 297, 298]
    if frame_number == 600:
        o[-7:] = [599, 595, 593, 597, 594, 596, 598]
    return o, typ


def select_references(order, buffer_order):
    """utils.py:154-177 -> (index of ref1, index of ref2) into the buffer: the two decoded frames nearest in time
    (``torch.topk(largest=False)`` tie order), the earlier one first; a single buffered frame serves as both."""
    d = torch.tensor([abs(i - order) for i in buffer_order])
    k = 1 if len(buffer_order) == 1 else 2
    ind = torch.topk(d, k, largest=False).indices.tolist()
    if k == 1:
        return ind[0], ind[0]
    lo, hi = (ind[0], ind[1]) if buffer_order[ind[0]] < buffer_order[ind[1]] else (ind[1], ind[0])
    return lo, hi


def get_scales(order, order1, order2):
    """utils.py:225-243: temporal position of the current frame between its references."""
    if order2 - order1 == 0:
        return 0, 0
    return (order - order1) / (order2 - order1), (order - order2) / (order1 - order2)


@torch.no_grad()
def code_sequence(model, frames, s, crop, intra_size=16, buffer_len=32):
    """test.py:37-93 for one rate level ``s``: frames [T,3,H,W] (padded); I-frames pass through uncoded (the ELIC
    intra codec is outside the B-frame path); the reference buffer holds clamp(dec, 0, 1).
    Returns (bits[T], sse_u8[T], down_ratio[T]) as python lists."""
    T = frames.shape[0]
    order_list, typ = get_order_typ_list(intra_size, T)
    h, w = crop
    bits, sse, ratios = [0.0] * T, [0.0] * T, [0] * T
    buf, buf_order = [], []
    for order in order_list:
        x = frames[order:order + 1]
        if typ[order] == "I":
            dec = x
        else:
            i1, i2 = select_references(order, buf_order)
            s1, s2 = get_scales(order, buf_order[i1], buf_order[i2])
            out = code_frame(model, buf[i1], buf[i2], x, s1, s2, s)
            dec, bits[order], ratios[order] = out["x_hat"], out["bits"], out["down_ratio"]
        u8 = lambda t: torch.round(torch.clamp(t[0, :, :h, :w], 0.0, 1.0) * 255.0)
        sse[order] = ((u8(dec) - u8(x)).double() ** 2).sum().item()
        buf, buf_order = buf + [torch.clamp(dec, 0, 1)], buf_order + [order]
        if len(buf) > buffer_len:
            buf, buf_order = buf[1:], buf_order[1:]
    return bits, sse, ratios
