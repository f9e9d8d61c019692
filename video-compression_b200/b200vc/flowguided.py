"""Host-side mirror of the ICIP2024 flow-guided deformable B-frame codec (reference: ``ICIP2024/src/model/m.py``
``FlowGuidedB``; ``helpers.py`` ``MS_Feature`` / ``FlowNET`` / ``OffsetTemproalEnc`` / ``ResidualTemproalEnc`` /
``OffsetDiversity`` / ``Reconstuctor``; ``compression_bottlenecks.py`` ``Offset_ELIC`` / ``Res_ELIC``; evaluation loop
``ICIP2024/src/test.py:37-93``, ``utils.py:154-243``, ``opt_helpers.py:23-51``).  Same class / attribute names, call
signatures and state-dict keys, so the reference's checkpoints load unchanged; ICIP2023's ``DeformB`` shares the same
operators (``ICIP2023/src/model/m.py``).

What runs in the sm_100a kernels: the three feature-pyramid backward warps per reference (K-WARP, align_corners=True,
64 / 96 / 128 channels), the three flow-guided modulated deformable fusions (K-DCN), both factorised priors with their
hyper-gains, the checkerboard x channel-group context loops (K-CHK glue + K-GC likelihoods, bits reduced in-kernel:
no likelihood tensor is written), the 5-ratio flow-only search (warp + blend + clamp + SSE fused, one host sync per
frame) and the uint8 PSNR.  Convolutions stay on cuDNN (out of scope).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import icip, ops
from . import modules as M

LEVELS = 5  # compression_bottlenecks.py:204


def _conv(i, o, k=5, s=2):
    return nn.Conv2d(i, o, kernel_size=k, stride=s, padding=k // 2)


def _deconv(i, o, k=5, s=2):
    return nn.ConvTranspose2d(i, o, kernel_size=k, stride=s, output_padding=s - 1, padding=k // 2)


class ResidualBottleneckBlock(nn.Module):
    """ICIP2024/src/model/elic.py:69-84."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.BottleneckBlock = nn.Sequential(
            nn.Conv2d(in_ch, out_ch, 1), nn.ReLU(inplace=True), nn.Conv2d(out_ch, out_ch, 3, padding=1),
            nn.ReLU(inplace=True), nn.Conv2d(out_ch, out_ch, 1))

    def forward(self, x):
        return self.BottleneckBlock(x) + x


def _seq(ch, blocks, head=None, tail=None):
    layers = [] if head is None else [head]
    layers += [ResidualBottleneckBlock(ch, ch) for _ in range(blocks)]
    if tail is not None:
        layers.append(tail)
    return nn.Sequential(*layers)


class MS_Feature(nn.Module):
    def __init__(self):
        super().__init__()
        self.layer1 = _seq(64, 3, _conv(3, 64, 3, 2))
        self.layer2 = _seq(96, 3, _conv(64, 96, 3, 2))
        self.layer3 = _seq(128, 3, _conv(96, 128, 3, 2))

    def forward(self, x):
        l1 = self.layer1(x)
        l2 = self.layer2(l1)
        return l1, l2, self.layer3(l2)


class FlowNET(nn.Module):
    def __init__(self):
        super().__init__()
        c = (32, 64, 128, 192)
        self.down0 = _seq(c[0], 2, _conv(6, c[0], 3, 2))
        self.down1 = _seq(c[1], 2, _conv(c[0], c[1], 3, 2))
        self.down2 = _seq(c[2], 2, _conv(c[1], c[2], 3, 2))
        self.down3 = _seq(c[3], 2, _conv(c[2], c[3], 3, 2))
        self.up0 = _seq(c[3], 2, None, M.subpel_conv3x3(c[3], c[2], 2))
        self.up1 = _seq(c[2], 2, _conv(2 * c[2], c[2], 1, 1), M.subpel_conv3x3(c[2], c[1], 2))
        self.up2 = _seq(c[1], 2, _conv(2 * c[1], c[1], 1, 1), M.subpel_conv3x3(c[1], c[0], 2))
        self.up3 = _seq(c[0], 2, _conv(2 * c[0], c[0], 1, 1), M.subpel_conv3x3(c[0], 4, 2))

    def forward(self, inp):
        s0 = self.down0(inp)
        s1 = self.down1(s0)
        s2 = self.down2(s1)
        x = self.up0(self.down3(s2))
        x = self.up1(torch.cat((x, s2), 1))
        x = self.up2(torch.cat((x, s1), 1))
        return self.up3(torch.cat((x, s0), 1))


class _TemporalEnc(nn.Module):
    def __init__(self, mult, N=128, M_=128):
        super().__init__()
        self.g_a1 = _seq(N, 3, _conv(64 * mult, N))
        self.g_a2 = _seq(N, 3, _conv(N + 96 * mult, N))
        self.g_a3 = _seq(M_, 3, _conv(N + 128 * mult, M_))

    def forward(self, l1, l2, l3):
        y = self.g_a1(l1)
        y = self.g_a2(torch.cat([y, l2], dim=1))
        return self.g_a3(torch.cat([y, l3], dim=1))


class OffsetTemproalEnc(_TemporalEnc):
    def __init__(self, N=128, M=128):
        super().__init__(4, N, M)


class ResidualTemproalEnc(_TemporalEnc):
    def __init__(self, N=128, M=128):
        super().__init__(1, N, M)


class Reconstuctor(nn.Module):
    def __init__(self):
        super().__init__()
        self.layer3 = _seq(128, 3, None, M.subpel_conv3x3(128, 128, 2))
        self.layer2 = _seq(96, 3, _conv(128 + 96, 96, 1, 1), M.subpel_conv3x3(96, 96, 2))
        self.layer1 = _seq(64, 3, _conv(96 + 64, 64, 1, 1), M.subpel_conv3x3(64, 3, 2))

    def forward(self, c1, c2, c3):
        l3 = self.layer3(c3)
        l2 = self.layer2(torch.cat([c2, l3], dim=1))
        return self.layer1(torch.cat([c1, l2], dim=1))


class _GainedELIC(M.JointAutoregressiveHierarchicalPriors):
    """Offset_ELIC / Res_ELIC (compression_bottlenecks.py:72-311, :313-551): three-level analysis, gain-modulated
    latent and hyper-latent (5 rate levels, geometric interpolation between them), ELIC entropy model, three-level
    synthesis with the conditioning features re-injected."""

    def __init__(self, mult, dmult, outs, prefix, N=128, M_=128):
        super().__init__(N, M_)
        self.prefix = prefix
        self.g_a1 = _seq(N, 3, _conv(64 * mult, N))
        self.g_a2 = _seq(N, 3, _conv(N + 96 * mult, N))
        self.g_a3 = _seq(M_, 3, _conv(N + 128 * mult, M_))
        self.g_s3 = _seq(M_, 3, None, _deconv(M_, N))
        self.g_o3 = _seq(N, 3, _conv(N + 128 * dmult, N, 3, 1), _conv(N, outs[2], 3, 1))
        self.g_s2 = _seq(N, 3, _conv(N + 128 * dmult, N, 1, 1), _deconv(N, N))
        self.g_o2 = _seq(N, 3, _conv(N + 96 * dmult, N, 3, 1), _conv(N, outs[1], 3, 1))
        self.g_s1 = _seq(N, 3, _conv(N + 96 * dmult, N, 1, 1), _deconv(N, N))
        self.g_o1 = _seq(N, 3, _conv(N + 64 * dmult, N, 3, 1), _conv(N, outs[0], 3, 1))
        act = lambda: nn.ReLU(inplace=True)
        self.h_a = nn.Sequential(_conv(M_, N, 3, 1), act(), _conv(N, N), act(), _conv(N, N))
        self.h_s = nn.Sequential(_deconv(N, M_), act(), _deconv(M_, M_), act(), _conv(M_, M_, 3, 1))
        self.prior_fusion = _seq(2 * M_, 3, _conv(2 * M_, 2 * M_, 3, 1), _conv(2 * M_, 2 * M_, 3, 1))
        lact = lambda: nn.LeakyReLU(inplace=True)
        widths = (6, 6, 12, 24, M_ - 48)
        self.entropy_parameters = nn.ModuleList(
            nn.Sequential(nn.Conv2d(M_ * (4 if i == 0 else 6), M_ * 10 // 3, 1), lact(),
                          nn.Conv2d(M_ * 10 // 3, M_ * 8 // 3, 1), lact(), nn.Conv2d(M_ * 8 // 3, 2 * w, 1))
            for i, w in enumerate(widths))
        self.channel_context_models = nn.ModuleList(
            nn.Sequential(_conv(w, N, 5, 1), act(), _conv(N, N, 5, 1), act(), _conv(N, 2 * M_, 5, 1))
            for w in (6, 12, 24, 48))
        self.context_prediction_models = nn.ModuleList(
            M.CheckerboardContext(in_channels=w, out_channels=2 * M_, kernel_size=5, stride=1, padding=2)
            for w in widths)
        self.levels = LEVELS
        self.Gain = nn.Parameter(torch.ones(LEVELS, M_))
        self.InverseGain = nn.Parameter(torch.ones(LEVELS, M_))
        self.HyperGain = nn.Parameter(torch.ones(LEVELS, N))
        self.InverseHyperGain = nn.Parameter(torch.ones(LEVELS, N))

    def interpolate_gain(self, s):
        """(gain, hypergain, invhypergain, invgain) at (possibly fractional) rate level ``s``."""
        s = max(min(s, self.levels - 1), 0)
        hi, lo = int(min(math.ceil(s), self.levels - 1)), int(max(math.floor(s), 0))

        def pick(g):
            if hi == lo:
                return torch.abs(g[int(s)])
            l = hi - s
            return torch.abs(g[hi]) ** (1 - l) * torch.abs(g[lo]) ** l

        return pick(self.Gain), pick(self.HyperGain), pick(self.InverseHyperGain), pick(self.InverseGain)

    def _analysis(self, f, fd):
        raise NotImplementedError

    def _run(self, f, fd, temp, s, bits_only):
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        gain, hypergain, invhypergain, invgain = self.interpolate_gain(s)
        y = self._analysis(f, fd) * gain.view(1, -1, 1, 1)
        z = self.h_a(y)
        # factorised prior on z * hypergain: likelihood at round(. - median) + median (K-EB, gain folded in); the
        # reconstruction path uses ste_round(z * hypergain) WITHOUT the median (quirk B.8)
        rz = ops.entropy_bottleneck(z, M.eb_packed(eb), lik_bound=M._lik_bound(eb), gain=hypergain,
                                    want_z_hat=False, want_lik=not bits_only, want_bits=bits_only)
        z_hat, _ = ops.round_checker(z * hypergain.view(1, -1, 1, 1), want_half=False)
        hyper = self.prior_fusion(torch.cat([self.h_s(z_hat * invhypergain.view(1, -1, 1, 1)), temp], dim=1))
        liks, y_hat = icip.elic_context_likelihoods(y, hyper, self.context_prediction_models,
                                                    self.channel_context_models, self.entropy_parameters, gc,
                                                    inv_gain=invgain, bits_only=bits_only)
        inp3 = torch.cat([self.g_s3(y_hat), fd[2]], dim=1)
        inp2 = torch.cat([self.g_s2(inp3), fd[1]], dim=1)
        inp1 = torch.cat([self.g_s1(inp2), fd[0]], dim=1)
        p = self.prefix
        out = {p + "3": self.g_o3(inp3), p + "2": self.g_o2(inp2), p + "1": self.g_o1(inp1)}
        if bits_only:
            out["bits"] = rz["bits"] + liks
        else:
            lik = {"z": rz["lik"]}
            lik.update(liks)
            out["likelihoods"] = lik
        return out

    def forward(self, f1, f2, f3, f1d, f2d, f3d, temp, s):
        """Reference signature and result dict (likelihood tensors materialised)."""
        return self._run((f1, f2, f3), (f1d, f2d, f3d), temp, s, bits_only=False)

    def forward_bits(self, f1, f2, f3, f1d, f2d, f3d, temp, s):
        """Same computation; ``out["bits"]`` [N] float64 instead of the likelihood tensors."""
        return self._run((f1, f2, f3), (f1d, f2d, f3d), temp, s, bits_only=True)


class Offset_ELIC(_GainedELIC):
    def __init__(self, N=128, M=128, **kwargs):
        super().__init__(5, 4, (27 * 8 * 2,) * 3, "offset", N, M)

    def _analysis(self, f, fd):
        y = self.g_a1(f[0])
        y = self.g_a2(torch.cat([y, f[1]], dim=1))
        return self.g_a3(torch.cat([y, f[2]], dim=1))


class Res_ELIC(_GainedELIC):
    def __init__(self, N=128, M=128, **kwargs):
        super().__init__(2, 1, (64, 96, 128), "res", N, M)

    def _analysis(self, f, fd):
        y = self.g_a1(torch.cat([f[0], fd[0]], dim=1))
        y = self.g_a2(torch.cat([y, f[1], fd[1]], dim=1))
        return self.g_a3(torch.cat([y, f[2], fd[2]], dim=1))


class FlowGuidedB(nn.Module):
    """ICIP2024/src/model/m.py:31-282."""

    def __init__(self):
        super().__init__()
        self.feature_extractor = MS_Feature()
        self.flow_estimator = FlowNET()
        self.offset_temporal_conditioner = OffsetTemproalEnc()
        self.offset_compressor = Offset_ELIC()
        self.offset_diversity_l3 = icip.OffsetDiversity(in_channel=128, magnitude=10)
        self.offset_diversity_l2 = icip.OffsetDiversity(in_channel=96, magnitude=20)
        self.offset_diversity_l1 = icip.OffsetDiversity(in_channel=64, magnitude=40)
        self.residue_temporal_conditioner = ResidualTemproalEnc()
        self.residual_compressor = Res_ELIC()
        self.reconstructor = Reconstuctor()

    convert_scales = staticmethod(icip.convert_scales)

    def warp(self, img, flow):
        """m.py:262-282 -- K-WARP (align_corners=True, border), any channel count."""
        return ops.backwarp(img, flow, "ac1")

    @staticmethod
    def pad_flow(t):
        h, w = t.shape[2], t.shape[3]
        return F.pad(t, (0, (16 - w % 16) % 16, 0, (16 - h % 16) % 16))

    def estimate_flow(self, xref1, xref2, down_ratio):
        a, b = F.avg_pool2d(xref1, down_ratio * 2), F.avg_pool2d(xref2, down_ratio * 2)
        h, w = a.shape[2], a.shape[3]
        flow = self.flow_estimator(torch.cat((self.pad_flow(a), self.pad_flow(b)), dim=1))[:, :, :h, :w]
        return F.interpolate(flow, scale_factor=down_ratio, mode="bilinear", align_corners=False) * down_ratio

    def forward_device(self, xref1, xref2, scale1, scale2, xcur, s, down_ratio):
        """The whole B-frame step without host synchronisation -> (x_hat, bits[N] float64)."""
        scale1, scale2 = self.convert_scales(scale1, scale2, xcur)
        flow = self.estimate_flow(xref1, xref2, down_ratio)
        fref1, fref2, fcur = (self.feature_extractor(t) for t in (xref1, xref2, xcur))
        flows, cond = [], []
        for lvl in range(3):
            f21, f12 = torch.chunk(flow, 2, dim=1)
            c1, c2 = (f21 * scale1).contiguous(), (f12 * scale2).contiguous()
            flows.append((c1, c2))
            C = fref1[lvl].shape[1]
            buf = torch.empty((xcur.shape[0], 4 * C) + tuple(fref1[lvl].shape[2:]), device=xcur.device, dtype=xcur.dtype)
            ops.backwarp(fref1[lvl], c1, "ac1", out=buf[:, 0:C])          # cat(wref1, wref2, fref1, fref2) in place
            ops.backwarp(fref2[lvl], c2, "ac1", out=buf[:, C:2 * C])
            buf[:, 2 * C:3 * C], buf[:, 3 * C:] = fref1[lvl], fref2[lvl]
            cond.append(buf)
            if lvl < 2:
                flow = F.interpolate(flow, scale_factor=0.5, mode="bilinear", align_corners=False) * 0.5
        inp = [torch.cat((cond[l], fcur[l]), dim=1) for l in range(3)]
        off = self.offset_compressor.forward_bits(inp[0], inp[1], inp[2], cond[0], cond[1], cond[2],
                                                  self.offset_temporal_conditioner(*cond), s)
        comp = {}
        for lvl, od in ((2, self.offset_diversity_l3), (1, self.offset_diversity_l2), (0, self.offset_diversity_l1)):
            o1, o2 = torch.chunk(off["offset" + str(lvl + 1)], 2, dim=1)
            comp[lvl] = od(fref1[lvl], o1, flows[lvl][0], fref2[lvl], o2, flows[lvl][1])
        res = self.residual_compressor.forward_bits(fcur[0], fcur[1], fcur[2], comp[0], comp[1], comp[2],
                                                    self.residue_temporal_conditioner(comp[0], comp[1], comp[2]), s)
        x_hat = self.reconstructor(comp[0] + res["res1"], comp[1] + res["res2"], comp[2] + res["res3"])
        return x_hat, off["bits"] + res["bits"]

    def forward(self, xref1, xref2, scale1, scale2, xcur, s, down_ratio):
        """m.py:181-260: {"x_hat", "size" (bits, summed over the batch), "rate" (bits per pixel)}."""
        B, _, H, W = xcur.shape
        x_hat, bits = self.forward_device(xref1, xref2, scale1, scale2, xcur, s, down_ratio)
        size = bits.sum().float()
        return {"x_hat": x_hat, "size": size, "rate": size / (H * W * B)}


# ------------------------------------------------------------------------------------- evaluation loop
def get_order_typ_list(intra_size, frame_number):
    """ICIP2024/src/utils.py:190-222 (coding order incl. the reference's literal tails for 300 / 600 frames, I/B types)."""
    base = (16, 8, 4, 12, 2, 14, 6, 10, 1, 15, 3, 13, 5, 11, 7, 9)
    order = [0] + [base[i % 16] + 16 * (i // 16) for i in range(frame_number - 1)]
    tail = (frame_number - 1) % intra_size
    if tail:
        top = max(order[:-tail])
        order[-tail:] = [top + tail - i for i in range(tail)]
    types = ["I" if i % intra_size == 0 else "B" for i in range(frame_number)]
    types[-1] = "I"
    if frame_number == 300:
        order[-11:] = [299, 293, 290, 296, 289, 291, 292, 294, 295, 297, 298]
    if frame_number == 600:
        order[-7:] = [599, 595, 593, 597, 594, 596, 598]
    return order, types


def select_references(order, buffer_order):
    """utils.py:154-177: buffer positions of the two decoded frames nearest in time, the earlier one first."""
    dist = torch.tensor([abs(i - order) for i in buffer_order])
    if len(buffer_order) == 1:
        return 0, 0
    a, b = torch.topk(dist, 2, largest=False).indices.tolist()
    return (a, b) if buffer_order[a] < buffer_order[b] else (b, a)


def get_scales(order, order1, order2):
    """utils.py:225-243."""
    if order2 == order1:
        return 0, 0
    return (order - order1) / (order2 - order1), (order - order2) / (order1 - order2)


@torch.no_grad()
def code_frame(model, xref1, xref2, xcur, scale1, scale2, s):
    """test.py:60-76: the 5-ratio flow-only search (fused warp + blend + clamp + SSE, one host sync), then the model."""
    ratio, pred_psnr = icip.get_best_down_ratio_prediction(model, xref1, xref2, scale1, scale2, xcur)
    x_hat, bits = model.forward_device(xref1, xref2, scale1, scale2, xcur, s, ratio)
    return {"x_hat": x_hat, "bits": bits.sum(), "down_ratio": ratio, "pred_psnr": float(pred_psnr)}


class SequenceCoder:
    """test.py:37-93 for one rate level: hierarchical order, nearest-two reference selection from a 32-frame buffer of
    clamped decoded frames -- kept on the device (the reference moves every decoded frame to the CPU and back,
    test.py:61,93).  I-frames pass through uncoded (the ELIC intra codec is outside the B-frame path)."""

    def __init__(self, model, gop=16, buffer_len=32):
        self.model, self.gop, self.buffer_len = model, gop, buffer_len

    @torch.no_grad()
    def code(self, frames, crop, s, want_decoded=False):
        """frames [T,3,H,W] (padded) -> (bits[T] float64, sse_u8[T] float64, down_ratio list, decoded or None)."""
        T = frames.shape[0]
        order_list, types = get_order_typ_list(self.gop, T)
        h, w = crop
        dev = frames.device
        bits = torch.zeros(T, device=dev, dtype=torch.float64)
        sse = torch.zeros(T, device=dev, dtype=torch.float64)
        ratios = [0] * T
        decoded = [None] * T
        buf, buf_order = [], []
        for order in order_list:
            x = frames[order:order + 1]
            if types[order] == "I":
                dec = x
            else:
                i1, i2 = select_references(order, buf_order)
                s1, s2 = get_scales(order, buf_order[i1], buf_order[i2])
                out = code_frame(self.model, buf[i1], buf[i2], x, s1, s2, s)
                dec, ratios[order] = out["x_hat"], out["down_ratio"]
                bits[order] = out["bits"]
            sse[order] = ops.sse_u8(dec, x, h, w)[0]
            decoded[order] = dec
            buf, buf_order = buf + [dec.clamp(0, 1)], buf_order + [order]
            if len(buf) > self.buffer_len:
                buf, buf_order = buf[1:], buf_order[1:]
        return bits, sse, ratios, (torch.cat(decoded, 0) if want_decoded else None)


def calibrate_(model, seed=0):
    from . import synthetic
    return synthetic.calibrate_flowguided_(model, seed)
