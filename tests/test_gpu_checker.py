"""GPU parity: K-CHK (checkerboard context glue) and the channel-group entropy loop of the ICIP
ELIC-style compressors (ICIP2024/src/model/compression_bottlenecks.py:229-269) vs its torch restatement
(oracle/icip.py) with identical randomly initialised sub-modules.  Bars: the glue kernels are bit-exact; likelihoods
within 1e-5 relative (K-GC's bar)."""
import pytest
import torch
import torch.nn as nn

from oracle import cai
from oracle import icip as o_icip

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 128, 68, 120), (2, 7, 5, 9), (1, 1, 1, 1), (3, 64, 17, 30)])
def test_round_checker_is_bit_exact(shape):
    from b200vc import ops
    g = torch.Generator().manual_seed(sum(shape))
    y = (4.0 * torch.randn(shape, generator=g)).cuda()
    y.view(-1)[:6] = torch.tensor([0.5, 1.5, -0.5, -2.5, -0.3, 0.0], device="cuda")[:y.numel()][:min(6, y.numel())]
    want = o_icip.ste_round(y)
    half = want.clone()
    half[:, :, 0::2, 0::2] = 0
    half[:, :, 1::2, 1::2] = 0
    got, got_half = ops.round_checker(y)
    assert torch.equal(got, want) and torch.equal(got_half, half)
    # signed zeros too: the reference's (round(x) - x) + x turns -0.0 into +0.0
    assert torch.equal(torch.signbit(got), torch.signbit(want))
    only, none = ops.round_checker(y, want_half=False)
    assert none is None and torch.equal(only, want)


@pytest.mark.parametrize("parity", [0, 1])
def test_checker_mask_in_place_and_into_a_slice(parity):
    from b200vc import ops
    x = torch.randn(2, 5, 9, 14, device="cuda")
    want = x.clone()
    if parity == 1:
        want[:, :, 0::2, 1::2] = 0
        want[:, :, 1::2, 0::2] = 0
    else:
        want[:, :, 0::2, 0::2] = 0
        want[:, :, 1::2, 1::2] = 0
    assert torch.equal(ops.checker_mask(x, zero_parity=parity), want)
    buf = torch.full((2, 12, 9, 14), 7.0, device="cuda")
    ops.checker_mask(x, out=buf[:, 3:8], zero_parity=parity)
    assert torch.equal(buf[:, 3:8], want) and (buf[:, :3] == 7).all() and (buf[:, 8:] == 7).all()
    y = x.clone()
    ops.checker_mask(y, out=y, zero_parity=parity)
    assert torch.equal(y, want)


def _modules(M, N, seed):
    torch.manual_seed(seed)
    groups = [6, 6, 12, 24, M - 48]
    conv5 = lambda i, o: nn.Conv2d(i, o, 5, padding=2)
    ctx = nn.ModuleList(conv5(c, 2 * M) for c in groups)
    chan = nn.ModuleList(nn.Sequential(conv5(c, N), nn.ReLU(inplace=True), conv5(N, 2 * M)) for c in [6, 12, 24, 48])
    ent = nn.ModuleList(nn.Sequential(nn.Conv2d(i, M, 1), nn.LeakyReLU(inplace=True), nn.Conv2d(M, 2 * o, 1))
                        for i, o in zip([4 * M] + [6 * M] * 4, groups))
    return ctx.cuda().eval(), chan.cuda().eval(), ent.cuda().eval()


@pytest.mark.parametrize("N", [1, 2])
def test_context_loop_matches_reference_restatement(N, strict_fp32):
    from b200vc import icip, modules
    M, H, W = 64, 34, 60
    ctx, chan, ent = _modules(M, 32, 5)
    g = torch.Generator().manual_seed(9)
    y = (3.0 * torch.randn(N, M, H, W, generator=g)).cuda()
    hyper = torch.randn(N, 2 * M, H, W, generator=g).cuda()
    inv_gain = (1.0 + 0.1 * torch.randn(M, generator=g)).abs().cuda()
    gc_o = cai.GaussianConditional(None).cuda().eval()
    gc_p = modules.GaussianConditional(None).cuda().eval()
    with torch.no_grad():
        want, y_hat_o = o_icip.elic_context_likelihoods(y, hyper, ctx, chan, ent, gc_o, inv_gain)
        before = icip.ops.launch_count()
        got, y_hat_p = icip.elic_context_likelihoods(y, hyper, ctx, chan, ent, gc_p, inv_gain=inv_gain)
        launches = icip.ops.launch_count() - before
    assert torch.equal(y_hat_p, y_hat_o)
    assert set(got) == set(want) == {f"y_{i}" for i in range(5)}
    for k in want:
        rel = ((got[k] - want[k]).abs() / want[k]).max().item()
        print(f"context loop N={N} {k} {tuple(want[k].shape)}: max rel err of the likelihoods {rel:.2e}")
        assert got[k].shape == want[k].shape and rel < 1e-5
    assert launches == 1 + 5 + 5   # one quantisation pass, five mask passes, five likelihood passes
