"""GPU parity: the ICIP2024 down-ratio search (opt_helpers.py:23-51) with the fused warp->blend->clamp->SSE kernel
vs its torch restatement.  The flow estimator is a conv net (out of scope), so a deterministic stand-in with the
reference's ``estimate_flow(xref1, xref2, down_ratio)`` signature and output geometry (half resolution, 4 channels)
is used on both sides."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


class _FlowStub:
    """Half-resolution, down_ratio-dependent flow built from the frames themselves (m.py:84-102 geometry)."""

    def estimate_flow(self, xref1, xref2, down_ratio):
        d = F.avg_pool2d(xref1 - xref2, down_ratio * 2)
        flow = torch.cat([d[:, :2], -d[:, 1:3]], 1) * (6.0 + down_ratio)
        return F.interpolate(flow, scale_factor=down_ratio, mode="bilinear", align_corners=False) * down_ratio


@pytest.mark.parametrize("shape", [(1, 256, 384), (2, 128, 192)])
def test_down_ratio_search_matches_reference(shape):
    from b200vc import icip, synthetic
    from oracle import icip as o_icip
    N, H, W = shape
    seq = synthetic.make_sequence(3 * N, H, W, seed=21, device="cuda")
    xref1, xcur, xref2 = seq[0:N], seq[N:2 * N], seq[2 * N:3 * N]
    model = _FlowStub()
    for scales in ((0.5, 0.5), (0.33333, 0.66667)):
        for r in (1, 4, 16):
            want = o_icip.prediction_flowonly(model, xcur, xref1, xref2, *scales, r)
            got = icip.prediction_flowonly(model, xcur, xref1, xref2, *scales, r)
            assert torch.equal(got, want)
        best_o, psnr_o = o_icip.get_best_down_ratio_prediction(model, xref1, xref2, *scales, xcur, 0, 0)
        best_p, psnr_p = icip.get_best_down_ratio_prediction(model, xref1, xref2, *scales, xcur, 0, 0)
        print(f"search {shape} scales={scales}: best ratio {best_p} (oracle {best_o}), PSNR {psnr_p.item():.4f} dB")
        assert best_p == best_o
        assert abs(psnr_p.item() - psnr_o.item()) < 1e-3
    s1, s2 = icip.convert_scales(0.33333, 0.66667, xcur)
    assert s1.shape == (1, 1, 1, 1) and abs(s1.item() - 0.33) < 1e-6 and abs(s2.item() - 0.67) < 1e-6
