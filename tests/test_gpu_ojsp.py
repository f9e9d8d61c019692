"""GPU parity: the OJSP2025 down-sampling-ratio search (video_model.py:621-666) with the fused warp -> squared-error
kernel vs its torch restatement (oracle/ojsp.py).  The flow estimator is a conv net of the absent DCVC-FM code, so a
deterministic stand-in with ``optic_flow(cur, ref) -> [N,2,h,w]`` is used on both sides.  Bars: the warped frame is
bit-exact (same kernel arithmetic as K-WARP AC1); PSNR within 1e-3 dB (fp64 block sums vs torch's fp32 mean); the
selected ratio and motion field identical."""
import pytest
import torch
import torch.nn.functional as F

from oracle import ojsp as o_ojsp
from oracle import warp as o_warp
from gpu_util import warp_case

pytestmark = pytest.mark.gpu


class _FlowStub:
    """Motion 'estimate' from the frame difference: depends on the resolution, so every ratio gives another field."""

    def optic_flow(self, cur, ref):
        d = F.avg_pool2d(cur - ref, 3, stride=1, padding=1)
        return torch.cat([d[:, 0:1] - d[:, 2:3], d[:, 1:2]], 1) * 9.0


@pytest.mark.parametrize("shape", [(1, 3, 64, 96), (2, 3, 135, 241), (1, 3, 1088, 1920)])
def test_warp_sse_matches_oracle(shape):
    from b200vc import ops
    img, flow = warp_case(77 + shape[2], *shape)
    x = torch.rand(shape, generator=torch.Generator().manual_seed(5)).cuda()
    want_hat = o_warp.warp_ac1(img, flow)
    want = ((x - want_hat).double() ** 2).sum(dim=(1, 2, 3))
    sse, hat = ops.warp_sse(img, flow, x, "ac1", want_pred=True)
    assert torch.equal(hat, ops.backwarp(img, flow, "ac1"))
    assert (hat - want_hat).abs().max().item() < 1e-5
    rel = ((sse - want).abs() / want).max().item()
    print(f"warp_sse {shape}: rel err of SSE {rel:.2e}")
    assert rel < 1e-6
    sse2, none = ops.warp_sse(img, flow, x, "ac1")
    assert none is None and torch.equal(sse2, sse)


def test_warp_sse_4k_frame():
    """BASELINE config 5 geometry: 2160 x 3840."""
    from b200vc import ops
    img, flow = warp_case(4, 1, 3, 2160, 3840, amp=5.0)
    x = (img + 0.02 * torch.randn_like(img)).clamp(0, 1)
    sse, _ = ops.warp_sse(img, flow, x, "ac1")
    want = o_ojsp.PSNR(x, o_warp.warp_ac1(img, flow)).item()
    got = (10 * torch.log10(1.0 / (sse.sum() / x.numel()))).item()
    print(f"4K PSNR {got:.5f} dB (oracle {want:.5f})")
    assert abs(got - want) < 1e-3
    # zero flow with align_corners=True is the identity up to the rounding of the normalised coordinates
    # (~1e-4 px at W=3840, times the unit gradients of a white-noise image); the reference has the same residue
    sse0, hat0 = ops.warp_sse(img, torch.zeros_like(flow), img, "ac1", want_pred=True)
    assert (hat0 - o_warp.warp_ac1(img, torch.zeros_like(flow))).abs().max().item() < 1e-5
    assert (hat0 - img).abs().max().item() < 2e-3 and sse0.item() / img.numel() < 1e-7


@pytest.mark.parametrize("shape,prev", [((1, 270, 480), 1), ((1, 270, 480), 2.5), ((2, 136, 200), 8.75)])
def test_ratio_search_matches_reference(shape, prev):
    from b200vc import ojsp, synthetic
    N, H, W = shape
    seq = synthetic.make_sequence(2 * N, H, W, seed=31, device="cuda")
    x, ref = seq[0:N], seq[N:2 * N]
    dpb = {"ref_frame": ref, "ref_down_ratio": prev}
    model = _FlowStub()
    mv_o, ratio_o, psnr_o = o_ojsp.optimize_down_sampling_ratio(model, x, dpb)
    mv_p, ratio_p = ojsp.optimize_down_sampling_ratio(model, x, dpb)
    per = torch.stack([ojsp.warp_psnr(ref, ojsp.candidate_flow(model, x, ref, r), x) for r in ojsp.DOWNSAMPLING_RATIOS])
    err = (per.float().cpu() - psnr_o.float().cpu()).abs().max().item()
    print(f"search {shape} prev={prev}: ratio {ratio_p} (oracle {ratio_o}); max PSNR diff over 32 candidates {err:.2e} dB")
    assert err < 1e-3
    assert ratio_p == ratio_o
    assert torch.equal(mv_p, mv_o)


def test_ratio_search_keeps_the_previous_ratio_inside_the_bias():
    """video_model.py:655-660: a gain below 0.1 dB does not change the ratio."""
    from b200vc import ojsp, synthetic

    class _Weak(_FlowStub):
        def optic_flow(self, cur, ref):
            return super().optic_flow(cur, ref) * 0.02

    seq = synthetic.make_sequence(2, 136, 200, seed=8, device="cuda")
    dpb = {"ref_frame": seq[1:2], "ref_down_ratio": 4.25}
    mv_o, ratio_o, _ = o_ojsp.optimize_down_sampling_ratio(_Weak(), seq[0:1], dpb)
    mv_p, ratio_p = ojsp.optimize_down_sampling_ratio(_Weak(), seq[0:1], dpb)
    assert ratio_o == 4.25 and ratio_p == 4.25 and torch.equal(mv_p, mv_o)


def test_ratio_search_rejects_an_unknown_previous_ratio():
    from b200vc import ojsp, synthetic
    seq = synthetic.make_sequence(2, 64, 96, seed=3, device="cuda")
    with pytest.raises(ValueError):
        ojsp.optimize_down_sampling_ratio(_FlowStub(), seq[0:1], {"ref_frame": seq[1:2], "ref_down_ratio": 3.3})
