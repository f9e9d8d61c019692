"""TEST INFRASTRUCTURE (oracle).  Torch restatement of OJSP2025/video_model.py:621-666
(DMC.optimize_down_sampling_ratio, without its prints), :29-30 (PSNR) and :668-676 (DMC.warp, == oracle.warp.warp_ac1,
pinned bit for bit against the reference's own method by oracle/make_golden.py).  The rest of DCVC-FM is absent from
/root/reference's importable set (relative imports of missing modules), so the flow estimator is supplied by the
caller (``model.optic_flow``), as in the product mirror."""
import torch
import torch.nn.functional as F

from .warp import warp_ac1


def PSNR(x, y):
    return 10 * torch.log10(1 / torch.mean((x - y) ** 2))


def _candidate(model, x, ref, ratio, multiple=8):
    """One candidate of the search: shrink both frames by `ratio` (antialiased bilinear), pad bottom / right to a
    multiple of 8, estimate motion on the small pair, crop, enlarge the field back and scale the vectors by `ratio`."""
    small = [F.interpolate(t, scale_factor=1 / ratio, mode="bilinear", antialias=True) for t in (x, ref)]
    h, w = small[0].shape[-2:]
    pad = (0, (multiple - w % multiple) % multiple, 0, (multiple - h % multiple) % multiple)
    mv = model.optic_flow(F.pad(small[0], pad), F.pad(small[1], pad))[:, :, :h, :w]
    return F.interpolate(mv, size=tuple(x.shape[-2:]), mode="bilinear", antialias=True) * ratio


def optimize_down_sampling_ratio(model, x, dpb, bias=0.1):
    """Returns (motion field, ratio, PSNR per candidate).  Candidates 1, 1.25, ..., 8.75; the first strictly best PSNR
    wins; if it beats the previous frame's ratio (dpb["ref_down_ratio"]) by less than `bias` dB the previous ratio and
    its field are kept instead."""
    ratios = [1 + 0.25 * i for i in range(32)]
    ref = dpb["ref_frame"]
    fields = [_candidate(model, x, ref, r) for r in ratios]
    scores = [PSNR(x, warp_ac1(ref, mv)) for mv in fields]
    best = 0
    for i in range(1, len(ratios)):
        if scores[i] > scores[best]:
            best = i
    keep = ratios.index(dpb["ref_down_ratio"])          # the reference fails with an unbound name if it is not listed
    if (scores[best] - scores[keep]) < bias and ratios[keep] != ratios[best]:
        best = keep
    return fields[best], ratios[best], torch.stack(scores)
