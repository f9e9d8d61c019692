"""Host-side mirror of the down-sampling-ratio search of the OJSP2025 codec (reference:
``OJSP2025/video_model.py:621-666`` ``DMC.optimize_down_sampling_ratio``, ``:29-30`` ``PSNR``, ``:668-676``
``DMC.warp``).  For each of 32 candidate ratios the reference down-samples the current and the reference frame
(antialiased bilinear: torch, out of scope), estimates motion on the small pair (``self.optic_flow``: a conv net,
out of scope, supplied by the model), up-samples the field, warps the reference frame with it and takes the PSNR
against the current frame -- at 2160 x 3840 that is a 100 MB warped frame written and re-read 32 times, and one
``.item()``-style host sync per candidate (the ``print``).  Here warp -> squared error is ONE kernel per candidate
(``ops.warp_sse``), only fp64 partials are written, and the host synchronises once for the whole search.
"""
import torch
import torch.nn.functional as F

from . import ops

DOWNSAMPLING_RATIOS = tuple(1 + 0.25 * i for i in range(32))  # video_model.py:622: 1, 1.25, ... 8.75


def PSNR(x, y):
    """video_model.py:29-30, through the fused kernel when ``y`` is not needed: see ``warp_psnr``."""
    return 10 * torch.log10(1 / torch.mean((x - y) ** 2))


def candidate_flow(model, x, ref_frame, ratio, p=8):
    """video_model.py:629-642: the torch part of one candidate (resampling + the model's flow estimator)."""
    x_down = F.interpolate(x, scale_factor=1 / ratio, mode="bilinear", antialias=True)
    ref_down = F.interpolate(ref_frame, scale_factor=1 / ratio, mode="bilinear", antialias=True)
    _, _, h, w = x_down.size()
    pad_bottom, pad_right = (p - h % p) % p, (p - w % p) % p
    x_pad = F.pad(x_down, (0, pad_right, 0, pad_bottom))
    ref_pad = F.pad(ref_down, (0, pad_right, 0, pad_bottom))
    mv = model.optic_flow(x_pad, ref_pad)[:, :, :h, :w]
    return F.interpolate(mv, size=(x.shape[2], x.shape[3]), mode="bilinear", antialias=True) * ratio


def warp_psnr(ref_frame, flow, x):
    """``PSNR(x, self.warp(ref_frame, flow))`` (video_model.py:643-645) as a device scalar, no warped frame written."""
    sse, _ = ops.warp_sse(ref_frame, flow, x, "ac1")
    return 10 * torch.log10(1.0 / (sse.sum() / x.numel()))


def optimize_down_sampling_ratio(model, x, dpb, ratios=DOWNSAMPLING_RATIOS, bias=0.1):
    """video_model.py:621-666 -> (best_est_mv_down, best_ratio).  Same selection rule: the running best starts at
    -inf and only a strictly greater PSNR replaces it (the earliest best candidate wins, NaN never does); if the best
    PSNR beats the previous frame's ratio (``dpb["ref_down_ratio"]``) by less than ``bias`` dB the previous ratio and its
    motion field are kept.  Only the 32 PSNR scalars live through the sweep (the reference keeps two full-resolution
    fields; keeping all 32 would be 2.1 GB at 2160 x 3840): the winning field is recomputed after the single sync."""
    ref = dpb["ref_frame"]
    prev = dpb["ref_down_ratio"]
    if prev not in ratios:
        # the reference leaves prev_ratio_psnr unbound in this case (UnboundLocalError at video_model.py:657)
        raise ValueError(f"dpb['ref_down_ratio']={prev!r} is not one of the candidate ratios")
    psnr = torch.stack([warp_psnr(ref, candidate_flow(model, x, ref, r), x) for r in ratios]).float().cpu()  # one sync
    best, best_psnr = None, float("-inf")
    for i in range(len(ratios)):
        if psnr[i] > best_psnr:
            best, best_psnr = i, psnr[i].item()
    j = list(ratios).index(prev)
    if best is None or ((best_psnr - psnr[j].item()) < bias and prev != ratios[best]):
        best = j
    return candidate_flow(model, x, ref, ratios[best]), ratios[best]
