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


def optimize_down_sampling_ratio(model, x, dpb):
    downsampling_ratios = [1, 1.25, 1.5, 1.75, 2, 2.25, 2.5, 2.75, 3, 3.25, 3.5, 3.75, 4, 4.25, 4.5, 4.75, 5, 5.25,
                           5.5, 5.75, 6, 6.25, 6.5, 6.75, 7, 7.25, 7.5, 7.75, 8, 8.25, 8.5, 8.75]
    best_psnr = -float("inf")
    best_est_mv_down = None
    best_ratio = None
    psnrs = []
    for ratio in downsampling_ratios:
        x_down = F.interpolate(x, scale_factor=1 / ratio, mode="bilinear", antialias=True)
        ref_frame_down = F.interpolate(dpb["ref_frame"], scale_factor=1 / ratio, mode="bilinear", antialias=True)
        p = 8
        _, _, h, w = x_down.size()
        pad_bottom = (p - h % p) % p
        pad_right = (p - w % p) % p
        x_down_padded = F.pad(x_down, (0, pad_right, 0, pad_bottom))
        ref_frame_down_padded = F.pad(ref_frame_down, (0, pad_right, 0, pad_bottom))
        est_mv_down_padded = model.optic_flow(x_down_padded, ref_frame_down_padded)
        est_mv_down = est_mv_down_padded[:, :, :h, :w]
        est_mv_down = F.interpolate(est_mv_down, size=(x.shape[2], x.shape[3]), mode="bilinear", antialias=True) * ratio
        x_hat = warp_ac1(dpb["ref_frame"], est_mv_down)
        psnr = PSNR(x, x_hat)
        psnrs.append(psnr)
        if ratio == dpb["ref_down_ratio"]:
            prev_ratio_psnr = psnr
            prev_ratio_mv = est_mv_down
        if psnr > best_psnr:
            best_psnr = psnr
            best_est_mv_down = est_mv_down
            best_ratio = ratio
    bias = 0.1
    if (best_psnr - prev_ratio_psnr) < bias:
        if dpb["ref_down_ratio"] != best_ratio:
            best_est_mv_down = prev_ratio_mv
            best_ratio = dpb["ref_down_ratio"]
    return best_est_mv_down, best_ratio, torch.stack(psnrs)
