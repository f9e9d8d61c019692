"""TEST INFRASTRUCTURE (oracle).  Torch restatement of the reference's four backward-warp variants,
the flow glue, the blend/residual step and the bit sums.
Pinned against the reference's own functions by ``oracle/make_golden.py`` -> ``tests/golden/warp_*.npz``.

All functions are device-agnostic (the reference hard-codes ``device = torch.device("cuda")``).
"""
import math

import torch
import torch.nn.functional as F


def _pixel_centre_grid(H, W, device):
    """LHBDC/model/m.py:113-118, LHBDC/model/flow.py:17-20: linspace built on the CPU, then moved."""
    hor = torch.linspace(-1.0 + (1.0 / W), 1.0 - (1.0 / W), W).view(1, 1, 1, -1).expand(-1, -1, H, -1)
    ver = torch.linspace(-1.0 + (1.0 / H), 1.0 - (1.0 / H), H).view(1, 1, -1, 1).expand(-1, -1, -1, W)
    return torch.cat([hor, ver], 1).to(device)


def backwarp_lhbdc(img, flow):
    """W1 / W2: ``Model.backwarp`` LHBDC/model/m.py:111-126 == ``flow.backwarp`` LHBDC/model/flow.py:15-25.
    grid_sample(bilinear, border, align_corners=False) on a pixel-centre grid; flow scaled by 2/(W-1)."""
    H, W = flow.shape[2], flow.shape[3]
    grid = _pixel_centre_grid(H, W, img.device)
    nflow = torch.cat(
        [flow[:, 0:1] / ((img.shape[3] - 1.0) / 2.0), flow[:, 1:2] / ((img.shape[2] - 1.0) / 2.0)], 1
    )
    return F.grid_sample(
        input=img, grid=(grid + nflow).permute(0, 2, 3, 1), mode="bilinear", padding_mode="border",
        align_corners=False,
    )


def backwarp_flex(img, flow):
    """W3: ``BidirFlowRef.backwarp`` Flex-Rate.../b_model/b_model.py:99-112.  Integer meshgrid + flow,
    ``2*(x/W - 0.5)``, grid_sample defaults (bilinear, zeros, align_corners=False) => half-pixel shift."""
    _, _, H, W = img.size()
    gy, gx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")  # int64, like np.meshgrid(xy)
    gx = gx.to(img.device)
    gy = gy.to(img.device)
    u = flow[:, 0, :, :]
    v = flow[:, 1, :, :]
    x = gx.unsqueeze(0).expand_as(u).float() + u
    y = gy.unsqueeze(0).expand_as(v).float() + v
    normx = 2 * (x / W - 0.5)
    normy = 2 * (y / H - 0.5)
    grid = torch.stack((normx, normy), dim=3)
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


def warp_ac1(img, flow):
    """W4: ``FlowGuidedB.warp`` ICIP2024/src/model/m.py:262-282 (== helpers.py:61-69, OJSP2025/video_model.py:668-676).
    linspace(-1,1) grid, flow scaled by 2/(W-1), grid_sample(bilinear, border, align_corners=True)."""
    B, _, H, W = flow.shape
    xx = torch.linspace(-1.0, 1.0, W).view(1, 1, 1, W).expand(B, -1, H, -1)
    yy = torch.linspace(-1.0, 1.0, H).view(1, 1, H, 1).expand(B, -1, -1, W)
    grid = torch.cat([xx, yy], 1).to(img)
    nflow = torch.cat([flow[:, 0:1] / ((W - 1.0) / 2.0), flow[:, 1:2] / ((H - 1.0) / 2.0)], 1)
    return F.grid_sample(
        input=img, grid=(grid + nflow).permute(0, 2, 3, 1), mode="bilinear", padding_mode="border",
        align_corners=True,
    )


def upsample4(flow):
    """G0: ``nn.Upsample(scale_factor=4, mode='bilinear')`` LHBDC/model/m.py:30,57,59 (align_corners=False)."""
    return F.interpolate(flow, scale_factor=4, mode="bilinear", align_corners=False)


def lhbdc_flow_glue(flow_hat4, flow_ab, flow_ba, hh, ww):
    """G0: LHBDC/model/m.py:55-59.  chunk, add the linear-motion prior, crop, x4 upsample."""
    cb, ca = torch.chunk(flow_hat4, 2, dim=1)
    cb = upsample4((cb + flow_ab)[:, :, :hh, :ww])
    ca = upsample4((ca + flow_ba)[:, :, :hh, :ww])
    return cb, ca


def blend_residual_lhbdc(mask1, fw, bw, x_cur):
    """B0: LHBDC/model/m.py:63-67."""
    mask = mask1.repeat([1, 3, 1, 1])
    pred = mask * fw + (1.0 - mask) * bw
    return pred, x_cur - pred


def blend_residual_flex(mask_logits2, x_b, x_a, x_cur):
    """B0: Flex-Rate.../b_model/b_model.py:68-73 (sigmoid, 0.5 weights, normalised blend)."""
    mask = torch.sigmoid(mask_logits2)
    w1, w2 = 0.5 * mask[:, 0:1], 0.5 * mask[:, 1:2]
    pred = (w1 * x_b + w2 * x_a) / (w1 + w2 + 1e-8)
    return pred, x_cur - pred


def blend_half_mse(w1, w2, x_cur):
    """B0 (search form): ICIP2024/src/opt_helpers.py:35-36,45: 0.5/0.5 blend -> clamp -> MSE."""
    pred = 0.5 * w1 + (1 - 0.5) * w2
    return torch.mean((torch.clamp(pred, 0, 1) - x_cur) ** 2)


def bits_fp32(lik):
    """Q4, reference style: ``torch.log(lik).sum() / (-ln 2)`` (LHBDC/model/m.py:83-91)."""
    return torch.log(lik).sum() / (-math.log(2))


def bits_fp64(lik):
    """Q4, order-independent companion total."""
    return (-torch.log2(lik.double())).sum()


def bits_per_sample_fp32(lik):
    """Q4, Flex form: per-sample ``sum(dim=(1,2,3))`` (Flex-Rate.../b_model/b_model.py:80-90)."""
    return torch.log(lik).sum(dim=(1, 2, 3)) / (-math.log(2))


def reflect_pad64(im):
    """``Model.pad`` LHBDC/model/m.py:101-108: bottom/right reflection to a multiple of 64."""
    H, W = im.shape[2], im.shape[3]
    p1 = (64 - (H % 64)) % 64
    p2 = (64 - (W % 64)) % 64
    return F.pad(im, (0, p2, 0, p1), mode="reflect")
