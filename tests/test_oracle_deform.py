"""CPU: the deformable-convolution oracle (oracle/deform.py) against the real ``torchvision.ops.deform_conv2d`` --
the operator the reference calls (ICIP2023/src/model/m.py:29-34, ICIP2024/src/model/helpers.py:40,57) -- including
strides, dilation, groups != offset groups, missing mask / bias, and samples far outside the image."""
import pytest
import torch

from oracle import deform as o_deform

tv = pytest.importorskip("torchvision.ops")

CASES = [
    # N, Cin, Cout, H, W, k, stride, pad, dil, groups, offset groups, mask, bias
    (1, 8, 4, 10, 12, 3, 1, 1, 1, 2, 2, True, True),
    (2, 12, 12, 9, 7, 3, 1, 1, 1, 4, 2, True, False),
    (1, 16, 8, 11, 13, 3, 2, 1, 1, 8, 8, True, True),
    (1, 6, 6, 8, 8, 3, 1, 2, 2, 1, 3, False, True),
    (2, 4, 8, 6, 9, 1, 1, 0, 1, 2, 1, True, True),
    (1, 8, 8, 7, 7, 5, 1, 2, 1, 8, 4, False, False),
]


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_torchvision_cpu(case):
    N, Cin, Cout, H, W, k, s, p, d, groups, og, use_mask, use_bias = case
    g = torch.Generator().manual_seed(sum(case[:6]))
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin // groups, k, k, generator=g) * 0.3
    Ho = (H + 2 * p - (d * (k - 1) + 1)) // s + 1
    Wo = (W + 2 * p - (d * (k - 1) + 1)) // s + 1
    off = 2.5 * torch.randn(N, 2 * og * k * k, Ho, Wo, generator=g)
    off[:, :, 0, 0] *= 20.0                                   # far outside the image
    m = torch.sigmoid(torch.randn(N, og * k * k, Ho, Wo, generator=g)) if use_mask else None
    b = torch.randn(Cout, generator=g) if use_bias else None
    want = tv.deform_conv2d(x, off, w, b, stride=(s, s), padding=(p, p), dilation=(d, d), mask=m)
    got = o_deform.deform_conv2d(x, off, w, b, stride=(s, s), padding=(p, p), dilation=(d, d), mask=m)
    assert got.shape == want.shape
    err = (got - want).abs().max().item()
    assert err < 2e-5 * max(1.0, want.abs().max().item()), err


def test_offset_diversity_and_ste_round_match_the_reference_classes():
    """Golden vectors produced by the reference's own ``OffsetDiversity`` (ICIP2024/src/model/helpers.py:35-69, on
    torchvision) and ``ste_round`` (compression_bottlenecks.py:36-47): oracle/make_golden_icip.py."""
    import os

    import numpy as np

    from oracle import icip as o_icip
    from oracle import warp as o_warp
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "icip_reference.npz"))
    t = lambda k: torch.from_numpy(g[k])
    off1, m1 = o_icip.offset_diversity_prep(t("o1"), t("f1"), 10.0)
    assert torch.equal(off1, t("off1")) and torch.equal(m1, t("m1"))
    out = o_icip.offset_diversity_forward(t("weight"), t("bias"), 10.0, t("x1"), t("o1"), t("f1"), t("x2"), t("o2"), t("f2"))
    assert (out - t("out")).abs().max().item() < 2e-5 * max(1.0, t("out").abs().max().item())
    assert torch.equal(o_icip.ste_round(t("ste_in")), t("ste_out"))
    assert torch.equal(o_warp.warp_ac1(t("x1"), t("f1")), t("warped"))
