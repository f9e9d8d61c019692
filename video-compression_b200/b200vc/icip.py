"""Host-side mirror of the motion-adaptive down-ratio search of the ICIP2024 codec (reference:
``ICIP2024/src/opt_helpers.py:23-51`` ``prediction_flowonly`` / ``get_best_down_ratio_prediction``,
``ICIP2024/src/model/m.py:71-82`` ``convert_scales``; the same loop with 32 candidates lives at
``OJSP2025/video_model.py:621-666``).  For every candidate ratio the reference estimates flow (a conv net: out of
scope, supplied by ``model.estimate_flow``), x2-upsamples and scales it, warps both references (align_corners=True,
border), blends 0.5/0.5, clamps and takes the MSE -- four full-frame tensors per candidate.  Here the warp -> blend ->
clamp -> squared-error chain is ONE kernel per candidate (``ops.warp2_half_sse``), nothing but fp64 partials is
written, and the host synchronises once for the whole search instead of once per candidate.

Also here: the deformable alignment operator of the ICIP codecs (``torchvision.ops.DeformConv2d`` at
``ICIP2023/src/model/m.py:29-34`` and ``ICIP2024/src/model/helpers.py:35-69`` ``OffsetDiversity``) on the K-DCN kernel.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

DOWN_RATIOS = (1, 2, 4, 8, 16)


def convert_scales(scale1, scale2, x):
    """m.py:71-82: python / tensor scales -> [B,1,1,1] float tensors rounded to two decimals."""
    if not torch.is_tensor(scale1):
        scale1, scale2 = torch.tensor([scale1]), torch.tensor([scale2])
    scale1 = scale1.view(-1, 1, 1, 1).to(x.device).float()
    scale2 = scale2.view(-1, 1, 1, 1).to(x.device).float()
    return torch.round(scale1 * 10 ** 2) / (10 ** 2), torch.round(scale2 * 10 ** 2) / (10 ** 2)


def _candidate_flows(model, xref1, xref2, scale1, scale2, down_ratio):
    """opt_helpers.py:24-31 (flow glue, torch: tiny 2-channel tensors)."""
    f21, f12 = model.estimate_flow(xref1, xref2, down_ratio).chunk(2, 1)
    f21 = F.interpolate(f21, scale_factor=2, mode="bilinear", align_corners=False) * 2
    f12 = F.interpolate(f12, scale_factor=2, mode="bilinear", align_corners=False) * 2
    return f21 * scale1, f12 * scale2


def prediction_flowonly(model, xcur, xref1, xref2, scale1, scale2, down_ratio):
    """opt_helpers.py:23-38: the flow-only prediction 0.5*warp(xref1) + 0.5*warp(xref2)."""
    scale1, scale2 = convert_scales(scale1, scale2, xref1)
    f1, f2 = _candidate_flows(model, xref1, xref2, scale1, scale2, down_ratio)
    _, pred = ops.warp2_half_sse(xref1, xref2, f1, f2, xcur, "ac1", want_pred=True)
    return pred


def get_best_down_ratio_prediction(model, xref1, xref2, scale1, scale2, xcur, level=None, beta=None,
                                   ratios=DOWN_RATIOS):
    """opt_helpers.py:41-51: (best_down_ratio, best_pred_psnr); ties keep the earlier candidate (strict `>`)."""
    s1, s2 = convert_scales(scale1, scale2, xref1)
    sses = []
    for r in ratios:
        f1, f2 = _candidate_flows(model, xref1, xref2, s1, s2, r)
        sse, _ = ops.warp2_half_sse(xref1, xref2, f1, f2, xcur, "ac1")
        sses.append(sse.sum())
    mse = torch.stack(sses) / xcur.numel()
    psnr = (10 * torch.log10(1.0 / mse)).cpu()          # the search's only host synchronisation
    best, best_psnr = ratios[0], torch.tensor(0.0)
    for r, p in zip(ratios, psnr):
        if p > best_psnr:
            best, best_psnr = r, p
    return best, best_psnr.float()


# ---------------------------------------------------------------------------------- deformable alignment
def deform_conv_forward(mod, input, offset, mask=None):
    """``DeformConv2d.forward`` re-bound by ``patch()`` on torchvision's own module (weights stay where they are)."""
    return ops.deform_conv2d(input, offset, mod.weight, mod.bias, stride=mod.stride, padding=mod.padding,
                             dilation=mod.dilation, mask=mask)


class DeformConv2d(nn.Module):
    """``torchvision.ops.DeformConv2d`` (constructor, parameter names / shapes / init and ``forward(input, offset,
    mask=None)``) on the K-DCN kernel; state dicts are interchangeable with torchvision's module."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True):
        super().__init__()
        if in_channels % groups != 0 or out_channels % groups != 0:
            raise ValueError("in_channels and out_channels must be divisible by groups")
        pair = lambda v: (v, v) if isinstance(v, int) else tuple(v)
        self.in_channels, self.out_channels, self.groups = in_channels, out_channels, groups
        self.kernel_size, self.stride, self.padding, self.dilation = pair(kernel_size), pair(stride), pair(padding), \
            pair(dilation)
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.weight.shape[1] * self.kernel_size[0] * self.kernel_size[1]
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, input, offset, mask=None):
        return deform_conv_forward(self, input, offset, mask)


class OffsetDiversity(nn.Module):
    """ICIP2024/src/model/helpers.py:35-69: flow-guided modulated deformable fusion of two reference features."""

    def __init__(self, in_channel, magnitude):
        super().__init__()
        self.in_channel, self.magnitude = in_channel, magnitude
        self.fusion = DeformConv2d(in_channel * 2, in_channel, kernel_size=3, padding=1, groups=2 * 8)

    def prep(self, out, flow):
        o1, o2, mask = torch.chunk(out, 3, dim=1)
        mask = torch.sigmoid(mask)
        offset = torch.tanh(torch.cat((o1, o2), dim=1)) * self.magnitude
        offset = offset + flow.flip(1).repeat(1, offset.size(1) // 2, 1, 1)
        return offset, mask

    def forward(self, x1, offset1, flow1, x2, offset2, flow2):
        offset1, mask1 = self.prep(offset1, flow1)
        offset2, mask2 = self.prep(offset2, flow2)
        return self.fusion(torch.cat((x1, x2), dim=1), torch.cat((offset1, offset2), dim=1),
                           torch.cat((mask1, mask2), dim=1))

    def warp(self, img, flow):
        return ops.backwarp(img, flow, "ac1")


# ------------------------------------------------------------------------- checkerboard context loop
ELIC_GROUPS = (6, 6, 12, 24, None)  # compression_bottlenecks.py:229-235: channel groups, the last one takes the rest


def elic_context_likelihoods(y, hyper_params, context_prediction_models, channel_context_models, entropy_parameters,
                             gaussian_conditional, group_sizes=ELIC_GROUPS, inv_gain=None, bits_only=False):
    """The channel-group x checkerboard entropy loop of ``Offset_ELIC`` / ``Res_ELIC``
    (ICIP2024/src/model/compression_bottlenecks.py:229-269, :471-511) with the reference's own sub-modules.

    Returns ``({"y_0": lik, ...}, y_hat)`` where ``y_hat = ste_round(y) * inv_gain`` (``inv_gain`` [M] or None); with
    ``bits_only=True`` the first element is ``bits[N]`` (float64, sum of -log2 of every group's likelihoods) and no
    likelihood tensor is written.
    The latent is quantised ONCE (K-CHK ``round_checker``: rounded + anchor-zeroed copies; every group and every
    "earlier groups" context input is a channel slice of those), the context convolution's output is checkerboard-
    masked straight into the entropy-parameter network's input buffer (``checker_mask``), and the likelihoods come
    from ``gaussian_conditional`` (the K-GC kernel when the module is the b200vc mirror or was ``patch()``-ed)."""
    N, M, H, W = y.shape
    sizes, start = [], 0
    for s in group_sizes:
        s = M - start if s is None else s
        if s <= 0:
            raise ValueError(f"channel groups {group_sizes} do not fit {M} channels")
        sizes.append((start, start + s))
        start += s
    if start != M:
        raise ValueError(f"channel groups {group_sizes} do not cover {M} channels")
    y_hat, y_half = ops.round_checker(y)
    likelihoods, bits = {}, None
    for i, (a, b) in enumerate(sizes):
        ctx = context_prediction_models[i](y_half[:, a:b])
        parts = [ctx.shape[1]]
        chan = None
        if i > 0:
            chan = channel_context_models[i - 1](y_hat[:, :a])       # == ste_round(cat(groups[:i]))
            parts.append(chan.shape[1])
        parts.append(hyper_params.shape[1])
        buf = torch.empty((N, sum(parts), H, W), device=y.device, dtype=y.dtype)
        ops.checker_mask(ctx, out=buf[:, :parts[0]], zero_parity=1)
        if chan is not None:
            buf[:, parts[0]:parts[0] + parts[1]] = chan
        buf[:, -parts[-1]:] = hyper_params
        scales_hat, means_hat = entropy_parameters[i](buf).chunk(2, 1)
        if bits_only:
            from . import modules as M_
            r = ops.gauss_cond(y[:, a:b].contiguous(), scales_hat, means_hat,
                               scale_bound=M_._scale_bound(gaussian_conditional),
                               lik_bound=M_._lik_bound(gaussian_conditional), want_y_hat=False, want_lik=False)
            bits = r["bits"] if bits is None else bits + r["bits"]
        else:
            _, lik = gaussian_conditional(y[:, a:b], scales_hat, means=means_hat)
            likelihoods[f"y_{i}"] = lik
    if inv_gain is not None:
        y_hat = y_hat * inv_gain.view(1, -1, 1, 1)
    return (bits if bits_only else likelihoods), y_hat
