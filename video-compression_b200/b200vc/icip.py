"""Host-side mirror of the motion-adaptive down-ratio search of the ICIP2024 codec (reference:
``ICIP2024/src/opt_helpers.py:23-51`` ``prediction_flowonly`` / ``get_best_down_ratio_prediction``,
``ICIP2024/src/model/m.py:71-82`` ``convert_scales``; the same loop with 32 candidates lives at
``OJSP2025/video_model.py:621-666``).  For every candidate ratio the reference estimates flow (a conv net: out of
scope, supplied by ``model.estimate_flow``), x2-upsamples and scales it, warps both references (align_corners=True,
border), blends 0.5/0.5, clamps and takes the MSE -- four full-frame tensors per candidate.  Here the warp -> blend ->
clamp -> squared-error chain is ONE kernel per candidate (``ops.warp2_half_sse``), nothing but fp64 partials is
written, and the host synchronises once for the whole search instead of once per candidate.
"""
import torch
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
