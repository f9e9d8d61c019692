"""GPU parity: K-BLEND (three reference forms) and the uint8-domain SSE kernel."""
import numpy as np
import pytest
import torch

from oracle import warp as o_warp

pytestmark = pytest.mark.gpu


def _inputs(N, H, W, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g).cuda()
    return r(N, 3, H, W), r(N, 3, H, W), r(N, 3, H, W)


@pytest.mark.parametrize("shape", [(1, 1088, 1920), (2, 33, 47), (1, 8, 4)])
def test_blend_mask_is_bit_exact(shape):
    from b200vc import ops
    N, H, W = shape
    fw, bw, x = _inputs(N, H, W)
    mask = torch.rand(N, 1, H, W, device="cuda")
    pred_o, res_o = o_warp.blend_residual_lhbdc(mask, fw, bw, x)
    both = torch.cat([fw, bw], 1)  # the kernel reads the two halves of the concat buffer in place
    pred, res, _ = ops.blend_residual("mask", mask, both[:, :3], both[:, 3:], x)
    assert torch.equal(pred, pred_o) and torch.equal(res, res_o)
    _, res_only, _ = ops.blend_residual("mask", mask, fw, bw, x, want_pred=False)
    assert torch.equal(res_only, res_o)


def test_blend_flex_and_half_forms():
    from b200vc import ops
    xb, xa, x = _inputs(2, 40, 56, seed=1)
    logits = 3 * torch.randn(2, 2, 40, 56, device="cuda")
    pred_o, res_o = o_warp.blend_residual_flex(logits, xb, xa, x)
    pred, res, _ = ops.blend_residual("normw", logits, xb, xa, x)
    assert (pred - pred_o).abs().max().item() < 1e-6 and (res - res_o).abs().max().item() < 1e-6
    pred, _, sse = ops.blend_residual("half", None, xb * 1.5 - 0.2, xa, x, want_res=False, want_sse=True)
    assert torch.equal(pred, 0.5 * (xb * 1.5 - 0.2) + (1 - 0.5) * xa)
    for n in range(2):
        want = ((torch.clamp(pred[n], 0, 1) - x[n]).double() ** 2).sum().item()
        assert abs(sse[n].item() - want) / want < 1e-6
    mse_o = o_warp.blend_half_mse(xb * 1.5 - 0.2, xa, x).item()
    assert abs(sse.sum().item() / x.numel() - mse_o) / mse_o < 1e-5


def test_sse_u8_matches_the_numpy_psnr_path():
    """LHBDC/test/testing.py:176-182: float_to_uint8 on the unpadded crop, MSE in float64."""
    from b200vc import gop, ops
    a = (torch.rand(1, 3, 64, 96, device="cuda") * 1.2 - 0.1)
    b = (a + 0.02 * torch.randn_like(a))
    a[0, 0, 0, :4] = torch.tensor([0.5 / 255, 1.5 / 255, 2.5 / 255, 254.5 / 255])  # ties: round half to even
    h, w = 60, 90
    to_u8 = lambda t: np.round(np.clip(t[0, :, :h, :w].cpu().numpy(), 0, 1) * 255.).astype(np.uint8)
    ua, ub = to_u8(a).astype(np.float64), to_u8(b).astype(np.float64)
    want = ((ua - ub) ** 2).sum()
    got = ops.sse_u8(a, b, h, w)
    assert got.item() == want
    mse = want / ua.size
    assert abs(gop.psnr_from_sse(got, ua.size).item() - 10 * np.log10(255.0 ** 2 / mse)) < 1e-9
    # per-sample sums for a batch: frame 0 as above, frame 1 = its operands swapped and shifted
    a2, b2 = torch.cat([a, b.roll(3, 3)], 0), torch.cat([b, a], 0)
    got2 = ops.sse_u8(a2, b2, h, w)
    to_u8n = lambda t, n: np.round(np.clip(t[n, :, :h, :w].cpu().numpy(), 0, 1) * 255.).astype(np.float64)
    want2 = [((to_u8n(a2, n) - to_u8n(b2, n)) ** 2).sum() for n in range(2)]
    assert got2.shape == (2,) and got2[0].item() == want2[0] == want and got2[1].item() == want2[1]
