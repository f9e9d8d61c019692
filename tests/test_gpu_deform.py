"""GPU parity: K-DCN (modulated deformable convolution) through the C-ABI vs the real ``torchvision.ops.deform_conv2d``
CUDA operator -- the dependency the reference calls at ICIP2023/src/model/m.py:29-34 and
ICIP2024/src/model/helpers.py:40,57 -- and vs the oracle restatement.  Tolerance: 2e-5 of the output magnitude (the
per-pixel sums run in a different order than torchvision's cuBLAS GEMM; sampling itself is the same arithmetic)."""
import pytest
import torch

from oracle import deform as o_deform

pytestmark = pytest.mark.gpu
tv = pytest.importorskip("torchvision.ops")


def _case(seed, N, Cin, Cout, H, W, k, s, p, d, groups, og, use_mask=True, use_bias=True, amp=2.5):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin // groups, k, k, generator=g) * (1.0 / (Cin // groups * k * k) ** 0.5)
    Ho = (H + 2 * p - (d * (k - 1) + 1)) // s + 1
    Wo = (W + 2 * p - (d * (k - 1) + 1)) // s + 1
    off = amp * torch.randn(N, 2 * og * k * k, Ho, Wo, generator=g)
    off[:, :, 0, 0] *= 30.0
    m = torch.sigmoid(torch.randn(N, og * k * k, Ho, Wo, generator=g)) if use_mask else None
    b = torch.randn(Cout, generator=g) if use_bias else None
    cu = lambda t: t.cuda() if t is not None else None
    return cu(x), cu(off), cu(w), cu(b), cu(m), dict(stride=(s, s), padding=(p, p), dilation=(d, d))


CASES = [
    # the reference's layers: ICIP2024 fusion (2C -> C, groups 16) and ICIP2023 alignment (C -> C, groups 8)
    (1, 128, 64, 68, 120, 3, 1, 1, 1, 16, 16), (1, 192, 96, 34, 60, 3, 1, 1, 1, 16, 16), (1, 256, 128, 17, 30, 3, 1, 1, 1, 16, 16),
    (2, 96, 96, 40, 52, 3, 1, 1, 1, 8, 8), (1, 64, 64, 33, 47, 3, 1, 1, 1, 8, 8), (1, 32, 32, 64, 64, 3, 1, 1, 1, 8, 8),
    # generic geometry: stride, dilation, groups != offset groups, odd channel counts per group, 1x1 and 5x5 kernels
    (1, 16, 8, 21, 23, 3, 2, 1, 1, 8, 2), (2, 12, 20, 15, 17, 3, 1, 2, 2, 4, 3), (1, 6, 6, 9, 9, 1, 1, 0, 1, 2, 1),
    (1, 10, 15, 12, 14, 5, 1, 2, 1, 5, 2), (1, 8, 8, 16, 16, 3, 1, 1, 1, 1, 8),
]


@pytest.mark.parametrize("case", CASES)
def test_matches_torchvision_cuda(case):
    from b200vc import ops
    x, off, w, b, m, kw = _case(sum(case), *case)
    want = tv.deform_conv2d(x, off, w, b, mask=m, **kw)
    scale = max(1.0, want.abs().max().item())
    for fast in (True, False):   # group-channels-last path (where the shape qualifies) and the NCHW gather kernel
        got = ops.deform_conv2d(x, off, w, b, mask=m, use_workspace=fast, **kw)
        assert got.shape == want.shape and got.dtype == torch.float32
        err = (got - want).abs().max().item()
        print(f"dcn {case} workspace={fast}: max|diff|={err:.3e} (|out| max {scale:.2f})")
        assert err < 2e-5 * scale, err


@pytest.mark.parametrize("use_mask,use_bias", [(False, True), (True, False), (False, False)])
def test_optional_mask_and_bias(use_mask, use_bias):
    from b200vc import ops
    x, off, w, b, m, kw = _case(5, 1, 64, 64, 30, 44, 3, 1, 1, 1, 8, 8, use_mask, use_bias)
    want = tv.deform_conv2d(x, off, w, b, mask=m, **kw)
    got = ops.deform_conv2d(x, off, w, b, mask=m, **kw)
    assert (got - want).abs().max().item() < 2e-5 * max(1.0, want.abs().max().item())
    ora = o_deform.deform_conv2d(x, off, w, b, mask=m, **kw)
    assert (got - ora).abs().max().item() < 2e-5 * max(1.0, want.abs().max().item())


def test_zero_offsets_are_a_plain_grouped_convolution(strict_fp32):
    from b200vc import ops
    x, off, w, b, m, kw = _case(9, 2, 32, 32, 24, 28, 3, 1, 1, 1, 8, 8)
    got = ops.deform_conv2d(x, torch.zeros_like(off), w, b, mask=torch.ones_like(m), **kw)
    want = torch.nn.functional.conv2d(x, w, b, stride=1, padding=1, groups=8)
    assert (got - want).abs().max().item() < 2e-5 * max(1.0, want.abs().max().item())


def test_full_size_icip_level(strict_fp32):
    """ICIP2024 fusion at the first feature level of a 1080p frame: [1, 128, 544, 960] -> [1, 64, 544, 960]."""
    from b200vc import ops
    x, off, w, b, m, kw = _case(3, 1, 128, 64, 544, 960, 3, 1, 1, 1, 16, 16, amp=1.5)
    want = tv.deform_conv2d(x, off, w, b, mask=m, **kw)
    got = ops.deform_conv2d(x, off, w, b, mask=m, **kw)
    assert (got - want).abs().max().item() < 2e-5 * max(1.0, want.abs().max().item())


def test_module_mirror_and_patch_share_torchvision_state():
    import b200vc
    from b200vc import icip
    torch.manual_seed(1)
    ref = tv.DeformConv2d(64, 64, kernel_size=3, padding=1, groups=8).cuda()
    mine = icip.DeformConv2d(64, 64, kernel_size=3, padding=1, groups=8).cuda()
    mine.load_state_dict(ref.state_dict())
    x, off, _, _, m, _ = _case(2, 1, 64, 64, 20, 24, 3, 1, 1, 1, 8, 8)
    want = ref(x, off, m)
    assert (mine(x, off, m) - want).abs().max().item() < 2e-5 * max(1.0, want.abs().max().item())
    before = b200vc.ops.launch_count()
    patched = b200vc.patch(torch.nn.Sequential(ref))[0]
    got = patched(x, off, m)
    assert b200vc.ops.launch_count() == before + 1
    assert (got - want).abs().max().item() < 2e-5 * max(1.0, want.abs().max().item())


def test_offset_diversity_matches_reference_module():
    """ICIP2024/src/model/helpers.py:35-59 restated with torchvision's operator vs the mirror on K-DCN."""
    from b200vc import icip
    torch.manual_seed(4)
    C = 64
    mine = icip.OffsetDiversity(C, magnitude=10.0).cuda()
    ref_fusion = tv.DeformConv2d(2 * C, C, kernel_size=3, padding=1, groups=16).cuda()
    ref_fusion.load_state_dict(mine.fusion.state_dict())
    g = torch.Generator().manual_seed(6)
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    x1, x2 = r(1, C, 34, 60), r(1, C, 34, 60)
    o1, o2 = r(1, 3 * 8 * 9, 34, 60), r(1, 3 * 8 * 9, 34, 60)
    f1, f2 = 3 * r(1, 2, 34, 60), 3 * r(1, 2, 34, 60)
    off1, m1 = mine.prep(o1, f1)
    off2, m2 = mine.prep(o2, f2)
    want = ref_fusion(torch.cat((x1, x2), 1), torch.cat((off1, off2), 1), torch.cat((m1, m2), 1))
    got = mine(x1, o1, f1, x2, o2, f2)
    assert (got - want).abs().max().item() < 2e-5 * max(1.0, want.abs().max().item())


def test_bad_shapes_raise():
    from b200vc import ops
    x, off, w, b, m, kw = _case(1, 1, 16, 16, 10, 10, 3, 1, 1, 1, 4, 4)
    with pytest.raises(RuntimeError):
        ops.deform_conv2d(x, off[:, :-2], w, b, mask=m, **kw)
    with pytest.raises(RuntimeError):
        ops.deform_conv2d(x, off, w, b, mask=m[:, :-1], **kw)
    with pytest.raises(RuntimeError):
        ops.deform_conv2d(x.cpu(), off, w, b, mask=m, **kw)


def test_offset_diversity_matches_the_reference_golden(golden_dir):
    """vs the reference's own ``OffsetDiversity`` class run on torchvision (tests/golden/icip_reference.npz)."""
    import os

    import numpy as np
    from b200vc import icip, ops
    g = np.load(os.path.join(golden_dir, "icip_reference.npz"))
    t = lambda k: torch.from_numpy(g[k]).cuda()
    mod = icip.OffsetDiversity(16, 10.0).cuda().eval()
    mod.fusion.load_state_dict({"weight": t("weight"), "bias": t("bias")})
    with torch.no_grad():
        out = mod(t("x1"), t("o1"), t("f1"), t("x2"), t("o2"), t("f2"))
    assert (out - t("out")).abs().max().item() < 2e-5 * max(1.0, t("out").abs().max().item())
    assert (mod.warp(t("x1"), t("f1")) - t("warped")).abs().max().item() < 2e-5
    y_hat, _ = ops.round_checker(t("ste_in").view(1, 1, 8, 8))
    assert torch.equal(y_hat.view(-1), t("ste_out"))
