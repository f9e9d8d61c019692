"""TEST INFRASTRUCTURE (oracle).  Torch restatement of ICIP2024/src/opt_helpers.py:23-51 (prediction_flowonly,
get_best_down_ratio_prediction), ICIP2024/src/model/m.py:71-82 (convert_scales) and ICIP2024/src/utils.py:273-283
(MSE, PSNR).  The warp is ``oracle.warp.warp_ac1``, pinned bit for bit against the reference's own
``FlowGuidedB.warp`` by oracle/make_golden.py."""
import torch
import torch.nn.functional as F

from .warp import warp_ac1


def convert_scales(scale1, scale2, x):
    if not torch.is_tensor(scale1):
        scale1 = torch.tensor([scale1])
        scale2 = torch.tensor([scale2])
    scale1 = scale1.view(-1, 1, 1, 1).to(x.device).float()
    scale2 = scale2.view(-1, 1, 1, 1).to(x.device).float()
    scale1 = torch.round(scale1 * 10 ** 2) / (10 ** 2)
    scale2 = torch.round(scale2 * 10 ** 2) / (10 ** 2)
    return scale1, scale2


def prediction_flowonly(model, xcur, xref1, xref2, scale1, scale2, down_ratio):
    scale1, scale2 = convert_scales(scale1, scale2, xref1)
    f21, f12 = model.estimate_flow(xref1, xref2, down_ratio).chunk(2, 1)
    f21 = F.interpolate(f21, scale_factor=2, mode="bilinear", align_corners=False) * 2
    f12 = F.interpolate(f12, scale_factor=2, mode="bilinear", align_corners=False) * 2
    f21 = f21 * scale1
    f12 = f12 * scale2
    wref1 = warp_ac1(xref1, f21)
    wref2 = warp_ac1(xref2, f12)
    mask = 0.5
    return mask * wref1 + (1 - mask) * wref2


def get_best_down_ratio_prediction(model, xref1, xref2, scale1, scale2, xcur, level=None, beta=None):
    best_pred_psnr = 0
    for down_ratio in [1, 2, 4, 8, 16]:
        x_hat = prediction_flowonly(model, xcur, xref1, xref2, scale1, scale2, down_ratio)
        mse = torch.mean((torch.clamp(x_hat, 0, 1) - xcur) ** 2)
        psnr = 10 * torch.log10((1 ** 2) / mse)
        if psnr > best_pred_psnr:
            best_pred_psnr = psnr
            best_down_ratio = down_ratio
    return best_down_ratio, best_pred_psnr


# ------------------------------------------------------------------------------------------------------------------
# Checkerboard / channel-group context loop: ICIP2024/src/model/compression_bottlenecks.py:36-47 (ste_round) and
# :229-269 (Offset_ELIC.forward; Res_ELIC :471-511 is the same loop), restated with the sub-modules passed in.
def ste_round(x):
    return (torch.round(x) - x).detach() + x


def elic_context_likelihoods(y, hyper_params, context_prediction_models, channel_context_models, entropy_parameters,
                             gaussian_conditional, inv_gain=None):
    """Plain-torch statement of the channel-group / checkerboard entropy loop (reference lines cited above).
    Group g covers channels [lo_g, hi_g) with sizes 6, 6, 12, 24, rest.  For each group:
      1. quantise the group, blank its anchor sites ((row + col) even) and run the group's context conv;
      2. blank the non-anchor sites ((row + col) odd) of the conv output;
      3. the parameter net sees [that | channel context of all earlier quantised groups (absent for g = 0) | hyper];
      4. its output splits into (scales, means) for the group's Gaussian likelihood."""
    total = y.shape[1]
    edges = [0, 6, 12, 24, 48, total]
    rows = torch.arange(y.shape[2], device=y.device).view(-1, 1)
    cols = torch.arange(y.shape[3], device=y.device).view(1, -1)
    anchor = ((rows + cols) % 2 == 0)                    # [0::2, 0::2] and [1::2, 1::2]
    result = {}
    for g in range(5):
        lo, hi = edges[g], edges[g + 1]
        group = y[:, lo:hi]
        visible = ste_round(group).masked_fill(anchor, 0.0)
        ctx = context_prediction_models[g](visible).masked_fill(~anchor, 0.0)
        pieces = [ctx]
        if g > 0:
            pieces.append(channel_context_models[g - 1](ste_round(y[:, :lo])))
        pieces.append(hyper_params)
        scales, means = entropy_parameters[g](torch.cat(pieces, dim=1)).chunk(2, 1)
        result[f"y_{g}"] = gaussian_conditional(group, scales, means=means)[1]
    y_hat = ste_round(y)
    if inv_gain is not None:
        y_hat = y_hat * inv_gain.view(1, -1, 1, 1)
    return result, y_hat


# ------------------------------------------------------------------------------------------------------------------
# ICIP2024/src/model/helpers.py:35-59 (OffsetDiversity.prep / forward) restated on oracle.deform; pinned against the
# reference's own class run on torchvision (oracle/make_golden_icip.py -> tests/golden/icip_reference.npz).
def offset_diversity_prep(out, flow, magnitude):
    o1, o2, mask = torch.chunk(out, 3, dim=1)
    mask = torch.sigmoid(mask)
    offset = torch.tanh(torch.cat((o1, o2), dim=1)) * magnitude
    offset = offset + flow.flip(1).repeat(1, offset.size(1) // 2, 1, 1)
    return offset, mask


def offset_diversity_forward(weight, bias, magnitude, x1, offset1, flow1, x2, offset2, flow2):
    from .deform import deform_conv2d
    offset1, mask1 = offset_diversity_prep(offset1, flow1, magnitude)
    offset2, mask2 = offset_diversity_prep(offset2, flow2, magnitude)
    return deform_conv2d(torch.cat((x1, x2), dim=1), torch.cat((offset1, offset2), dim=1), weight, bias,
                         padding=(1, 1), mask=torch.cat((mask1, mask2), dim=1))
