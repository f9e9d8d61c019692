"""GPU parity: K-WARP / K-WARP2 through the C-ABI vs the oracle (torch grid_sample on the same device) and
vs the golden vectors produced by the reference's own functions.  Tolerance: the north star asks for warped
frames within 1e-5 relative in fp32; we hold max |diff| < 1e-5 on [0,1] images (and report bit-exactness)."""
import os

import numpy as np
import pytest
import torch

from oracle import warp as o_warp
from gpu_util import warp_case

pytestmark = pytest.mark.gpu

ORACLE = {"lhbdc": o_warp.backwarp_lhbdc, "flex": o_warp.backwarp_flex, "ac1": o_warp.warp_ac1}
TOL = 1e-5


@pytest.mark.parametrize("variant", ["lhbdc", "flex", "ac1"])
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_warp_matches_reference_golden(golden_dir, variant, tag):
    from b200vc import ops
    gold = np.load(os.path.join(golden_dir, "warp_reference.npz"))
    img = torch.from_numpy(gold[f"{tag}_img"]).cuda()
    flow = torch.from_numpy(gold[f"{tag}_flow"]).cuda()
    want = torch.from_numpy(gold[f"{tag}_{variant}"]).cuda()
    got = ops.backwarp(img, flow, variant)
    # golden came from ATen's CPU path, whose coordinate arithmetic differs in the last ulp from the CUDA
    # path the kernel mirrors: 1 ulp of a normalised coordinate ~ 1e-6 px at these sizes
    assert (got - want).abs().max().item() < 2e-5


@pytest.mark.parametrize("variant", ["lhbdc", "flex", "ac1"])
@pytest.mark.parametrize("shape", [(1, 3, 64, 96), (2, 3, 37, 53), (1, 1, 8, 8), (3, 2, 16, 12), (1, 64, 68, 120),
                                   (2, 13, 40, 52), (1, 96, 34, 60), (1, 3, 1, 40), (1, 3, 40, 1)])
def test_warp_matches_oracle_on_device(variant, shape):
    from b200vc import ops
    if variant != "flex" and 1 in shape[2:]:
        pytest.skip("(W-1)/2 == 0: the reference divides by zero")
    img, flow = warp_case(sum(shape) + 1000 * list(ORACLE).index(variant), *shape)
    want = ORACLE[variant](img, flow)
    got = ops.backwarp(img, flow, variant)
    assert got.shape == want.shape and got.dtype == torch.float32
    err = (got - want).abs().max().item()
    exact = (got == want).float().mean().item()
    print(f"warp[{variant}] {shape}: max|diff|={err:.3e} bit-exact={exact:.4f}")
    assert err < TOL, err


@pytest.mark.parametrize("variant", ["lhbdc", "flex", "ac1"])
def test_warp_full_hd(variant):
    """BASELINE config-2 geometry: 1920x1080 padded to 1088x1920."""
    from b200vc import ops
    img, flow = warp_case(7, 1, 3, 1088, 1920, amp=4.0)
    want = ORACLE[variant](img, flow)
    got = ops.backwarp(img, flow, variant)
    err = (got - want).abs().max().item()
    exact = (got == want).float().mean().item()
    print(f"warp[{variant}] 1088x1920: max|diff|={err:.3e} bit-exact={exact:.5f}")
    assert err < TOL, err


def test_warp_size_independent_properties():
    from b200vc import ops
    img, _ = warp_case(3, 2, 3, 136, 240)
    zero = torch.zeros(2, 2, 136, 240, device="cuda")
    # ICIP variant: zero flow is the identity up to the reference's own coordinate rounding (the unnormalised
    # coordinate carries ~1e-5 px of fp32 error at W=240; on a white-noise image that is ~2e-5 in value)
    assert (ops.backwarp(img, zero, "ac1") - img).abs().max().item() < 1e-4
    shift = zero.clone()
    shift[:, 0] = 3.0
    out = ops.backwarp(img, shift, "ac1")
    assert (out[..., :-3] - img[..., 3:]).abs().max().item() < 1e-4
    assert (out[..., -1] - img[..., -1]).abs().max().item() < 1e-4
    # LHBDC variant: zero flow is the identity up to coordinate rounding (2.4e-7)
    assert (ops.backwarp(img, zero, "lhbdc") - img).abs().max().item() < 1e-4
    # Flex variant: zero flow is the 2x2 box mean, fading to zero at the top/left border
    out = ops.backwarp(img, zero, "flex")
    box = (img[..., :-1, :-1] + img[..., 1:, :-1] + img[..., :-1, 1:] + img[..., 1:, 1:]) / 4
    assert (out[..., 1:, 1:] - box).abs().max().item() < 1e-4
    # linearity in the image
    _, flow = warp_case(4, 2, 3, 136, 240)
    a, b = ops.backwarp(img, flow, "lhbdc"), ops.backwarp(1 - img, flow, "lhbdc")
    assert (a + b - 1).abs().max().item() < 1e-5
    # everything pushed far outside: border variants return edge pixels, zeros variant returns 0
    far = torch.full_like(zero, 1e6)
    assert (ops.backwarp(img, far, "lhbdc") - img[..., -1:, -1:]).abs().max().item() < 1e-6
    assert ops.backwarp(img, far, "flex").abs().max().item() == 0.0


def test_warp_writes_into_concat_slice_and_rejects_bad_input():
    from b200vc import ops
    img, flow = warp_case(5, 2, 3, 32, 48)
    buf = torch.zeros(2, 8, 32, 48, device="cuda")
    ops.backwarp(img, flow, "lhbdc", out=buf[:, 3:6])
    assert torch.equal(buf[:, 3:6], ops.backwarp(img, flow, "lhbdc"))
    assert buf[:, :3].abs().max().item() == 0 and buf[:, 6:].abs().max().item() == 0
    with pytest.raises(RuntimeError, match="flow shape"):
        ops.backwarp(img, flow[:, :, :16], "lhbdc")
    with pytest.raises(RuntimeError, match="float32"):
        ops.backwarp(img.half(), flow, "lhbdc")


@pytest.mark.parametrize("shape", [(1, 192, 256), (2, 64, 128), (1, 1088, 1920), (3, 4, 4), (2, 36, 132), (1, 136, 260),
                                   (4, 1088, 1920)])
def test_warp2_matches_unfused_reference_chain(shape):
    """Fused m.py:55-63 vs the oracle's glue + two grid_samples + cat."""
    from b200vc import ops
    N, H, W = shape
    g = torch.Generator().manual_seed(11)
    hh, ww = H // 4, W // 4
    h4, w4 = hh + (64 - hh % 64) % 64, ww + (64 - ww % 64) % 64
    xb = torch.rand(N, 3, H, W, generator=g).cuda()
    xa = torch.rand(N, 3, H, W, generator=g).cuda()
    flow_hat = (2.0 * torch.randn(N, 4, h4, w4, generator=g)).cuda()
    fab = (1.5 * torch.randn(N, 2, h4, w4, generator=g)).cuda()
    fba = (1.5 * torch.randn(N, 2, h4, w4, generator=g)).cuda()
    # ~2 % of the vectors point far outside the frame: exercises the border clip (ix == W-1 / iy == H-1 exactly, where
    # ATen's clamped "+1" taps carry zero weight) on every side
    far = (torch.rand(N, 1, h4, w4, generator=g) < 0.02).cuda()
    flow_hat = torch.where(far, flow_hat * 300.0, flow_hat)
    cb, ca = o_warp.lhbdc_flow_glue(flow_hat, fab, fba, hh, ww)
    want = torch.cat([o_warp.backwarp_lhbdc(xb, cb), o_warp.backwarp_lhbdc(xa, ca)], 1)
    got, flows = ops.warp2_lhbdc(xb, xa, flow_hat, fab, fba, return_flows=True)
    ferr = (flows - torch.cat([cb, ca], 1)).abs().max().item()
    fexact = (flows == torch.cat([cb, ca], 1)).float().mean().item()
    err = (got - want).abs().max().item()
    print(f"warp2 {shape}: flow max|diff|={ferr:.3e} (bit-exact {fexact:.4f}); image max|diff|={err:.3e}")
    assert ferr == 0.0 and fexact == 1.0
    assert torch.equal(got, want), "fused glue + warps must equal the torch chain bit for bit"
    assert torch.equal(ops.warp2_lhbdc(xb, xa, flow_hat, fab, fba), got)   # the no-flows instantiation
    # given identical flows the fused warp equals the stand-alone kernel bit for bit
    again = torch.cat([ops.backwarp(xb, flows[:, :2].contiguous(), "lhbdc"),
                       ops.backwarp(xa, flows[:, 2:].contiguous(), "lhbdc")], 1)
    assert torch.equal(got, again)


@pytest.mark.parametrize("shape", [(1, 1088, 1920), (2, 45, 70)])
def test_search_form_warp_blend_sse(shape):
    """ICIP2024/src/opt_helpers.py:23-51 (prediction_flowonly + clamp + MSE) as one kernel."""
    from b200vc import ops
    N, H, W = shape
    x1, f1 = warp_case(31, N, 3, H, W, amp=3.0)
    x2, f2 = warp_case(32, N, 3, H, W, amp=3.0)
    xc = torch.rand(N, 3, H, W, device="cuda")
    want_pred = 0.5 * o_warp.warp_ac1(x1 * 1.3 - 0.1, f1) + (1 - 0.5) * o_warp.warp_ac1(x2, f2)
    sse, pred = ops.warp2_half_sse(x1 * 1.3 - 0.1, x2, f1, f2, xc, "ac1", want_pred=True)
    assert torch.equal(pred, want_pred)
    for n in range(N):
        want = ((torch.clamp(want_pred[n], 0, 1) - xc[n]).double() ** 2).sum().item()
        assert abs(sse[n].item() - want) / want < 1e-6
    mse = o_warp.blend_half_mse(o_warp.warp_ac1(x1 * 1.3 - 0.1, f1), o_warp.warp_ac1(x2, f2), xc).item()
    assert abs(sse.sum().item() / xc.numel() - mse) / mse < 1e-5
    sse2, none = ops.warp2_half_sse(x1 * 1.3 - 0.1, x2, f1, f2, xc, "ac1")
    assert none is None and torch.equal(sse, sse2)


@pytest.mark.parametrize("env", [{"B200VC_WARP_TMA": "0", "B200VC_WARP2_TMA": "0"},
                                 {"B200VC_WARP_TMA": "1", "B200VC_WARP2_TMA": "1"}])
@pytest.mark.parametrize("size", [(256, 256), (272, 480)])
def test_gather_and_tma_staged_kernels_agree_bit_for_bit(env, size):
    """Both implementations of every warp (direct gather / TMA-staged taps) are bit-exact against the oracle on
    smooth flows; the choice is read once per process, so each configuration runs in its own interpreter."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ, H=str(size[0]), W=str(size[1]), **env)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "try_warp_tma.py")], env=e, capture_output=True,
                       text=True, timeout=240)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
    assert r.stdout.count("bit-exact 1.0") >= 3


def test_warp_4k_frame():
    """BASELINE config 5 geometry (OJSP2025: 3840x2160, align_corners=True warp): the largest single-plane size."""
    from b200vc import ops
    g = torch.Generator().manual_seed(2160)
    img = torch.rand(1, 3, 2160, 3840, generator=g).cuda()
    f = 6.0 * torch.randn(1, 2, 2160 // 8, 3840 // 8, generator=g)
    flow = torch.nn.functional.interpolate(f, size=(2160, 3840), mode="bilinear", align_corners=False).cuda()
    want = o_warp.warp_ac1(img, flow)
    got = ops.backwarp(img, flow, "ac1")
    assert torch.equal(got, want)
    xc = torch.rand(1, 3, 2160, 3840, generator=g).cuda()
    sse, _ = ops.warp2_half_sse(img, img.flip(3), flow, -flow, xc, "ac1")
    pred = 0.5 * want + 0.5 * o_warp.warp_ac1(img.flip(3), -flow)
    ref = ((pred.clamp(0, 1) - xc).double() ** 2).sum().item()
    assert abs(sse.item() - ref) / ref < 1e-6


def test_empty_batches_return_empty_results():
    """Edge case the torch ops of the reference handle too: a zero-sized batch."""
    from b200vc import modules, ops
    e3 = torch.empty(0, 3, 16, 24, device="cuda")
    e2 = torch.empty(0, 2, 16, 24, device="cuda")
    assert ops.backwarp(e3, e2, "lhbdc").shape == (0, 3, 16, 24)
    pred, res, sse = ops.blend_residual("half", None, e3, e3, e3, want_sse=True)
    assert pred.shape == (0, 3, 16, 24) and sse.numel() == 0
    e128 = torch.empty(0, 128, 4, 4, device="cuda")
    r = ops.gauss_cond(e128, e128, e128)
    assert r["y_hat"].shape == (0, 128, 4, 4) and r["bits"].numel() == 0
    gdn = modules.GDN(128).cuda().eval()
    assert gdn(e128).shape == (0, 128, 4, 4)
