"""GPU parity: K-GDN / K-IGDN through the C-ABI vs the oracle GDN (x**2 -> conv1x1 -> rsqrt/sqrt -> mul) run
on the same device with TF32 disabled.  Tolerance 1e-5 relative (north star)."""
import pytest
import torch

from oracle import cai

pytestmark = pytest.mark.gpu


def _pair(C, inverse, trained_like, seed=0):
    from b200vc import modules
    g = torch.Generator().manual_seed(seed)
    o = cai.GDN(C, inverse=inverse)
    if trained_like:  # gamma = 0.1*I + 0.01*|N|, beta = 1 + 0.1*|N|
        with torch.no_grad():
            ped = o.beta_reparam.pedestal
            beta = 1.0 + 0.1 * torch.randn(C, generator=g).abs()
            gamma = 0.1 * torch.eye(C) + 0.01 * torch.randn(C, C, generator=g).abs()
            o.beta.copy_(torch.sqrt(torch.max(beta + ped, ped)))
            o.gamma.copy_(torch.sqrt(torch.max(gamma + ped, ped)))
    p = modules.GDN(C, inverse=inverse)
    p.load_state_dict(o.state_dict())
    return o.cuda().eval(), p.cuda().eval()


def _x(N, C, H, W, seed=1):
    g = torch.Generator().manual_seed(seed)
    scale = torch.exp(torch.empty(1, C, 1, 1).uniform_(-2.3, 2.3, generator=g))  # logU[0.1, 10] per channel
    return (torch.randn(N, C, H, W, generator=g) * scale).cuda()


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("C,shape", [(128, (1, 68, 120)), (128, (2, 33, 47)), (128, (1, 5, 8)), (64, (1, 40, 64)),
                                     (192, (1, 34, 60)), (128, (1, 136, 240)), (192, (1, 5, 8)), (192, (2, 40, 52)),
                                     (192, (3, 17, 20)), (192, (1, 136, 240))])
def test_gdn_matches_oracle(strict_fp32, impl, inverse, C, shape):
    from b200vc import modules, ops
    o, p = _pair(C, inverse, trained_like=True)
    x = _x(shape[0], C, shape[1], shape[2])
    with torch.no_grad():
        want = o(x)
        got = ops.gdn(x, modules.gdn_params(p), inverse=inverse, impl=impl)
    err = ((got - want).abs() / want.abs().clamp(min=1e-6)).max().item()
    print(f"gdn C={C} inv={inverse} impl={impl} {shape}: max rel err {err:.3e}")
    assert err < 1e-5, err


def test_gdn_at_init_closed_form_and_addend(strict_fp32):
    from b200vc import modules
    o, p = _pair(128, False, trained_like=False)
    x = _x(1, 128, 20, 36)
    skip = torch.randn_like(x)
    with torch.no_grad():
        got = modules.gdn_forward(p, x)
        torch.testing.assert_close(got, x / torch.sqrt(1 + 0.1 * x ** 2), rtol=1e-5, atol=1e-7)
        want = got + skip
        fused = modules.gdn_forward(p, x, addend=skip.clone())  # the addend tensor is consumed (in-place add)
        assert torch.equal(fused, want)


@pytest.mark.parametrize("C", [128, 192])
@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("inverse", [False, True])
def test_gdn_residual_add_both_forms(strict_fp32, impl, inverse, C):
    """out = gdn(x) + addend: separate output buffer vs in-place accumulation (TMA reduce-add on tcgen05)."""
    from b200vc import _lib, modules, ops
    o, p = _pair(C, inverse, trained_like=True)
    x = _x(2, C, 37, 52)
    skip = torch.randn_like(x)
    params = modules.gdn_params(p)
    with torch.no_grad():
        base = o(x)
        want = base + skip
    scale = (base.abs() + skip.abs()).clamp(min=1e-3)  # the sum may cancel: judge against the summands
    inplace = ops.gdn(x, params, inverse=inverse, addend=skip.clone(), impl=impl)
    out = torch.empty_like(x)
    rc = _lib.load().b200vc_gdn_f32(x.data_ptr(), params.data_ptr(), skip.data_ptr(), out.data_ptr(), 2, C,
                                    37 * 52, int(inverse), impl, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    for got in (inplace, out):
        err = ((got - want).abs() / scale).max().item()
        assert err < 1e-5, err
    assert torch.equal(inplace, out)


def test_gdn_params_track_weight_updates(strict_fp32):
    from b200vc import modules
    o, p = _pair(128, True, trained_like=True)
    x = _x(1, 128, 8, 12)
    with torch.no_grad():
        a = p(x)
        p.beta.mul_(1.5)
        o.beta.mul_(1.5)
        b = p(x)
        assert not torch.equal(a, b)
        torch.testing.assert_close(b, o(x), rtol=1e-5, atol=1e-7)


def test_gdn_full_hd_layer_and_determinism(strict_fp32):
    """Largest GDN of the 1080p residual path: [1,128,544,960]."""
    from b200vc import modules
    o, p = _pair(128, False, trained_like=True)
    x = _x(1, 128, 544, 960)
    with torch.no_grad():
        want = o(x)
        got = p(x)
        again = p(x)
    err = ((got - want).abs() / want.abs().clamp(min=1e-6)).max().item()
    print(f"gdn 544x960: max rel err {err:.3e}")
    assert err < 1e-5
    assert torch.equal(got, again)


def test_gdn_c192_tensor_core_layer(strict_fp32):
    """C = 192 (mbt2018_mean at quality >= 5; the joint-autoregressive base class) on the tcgen05 kernel: the second GDN
    of a 1080p I-frame analysis transform, [1,192,272,480]; deterministic; batch- and position-permutation invariant."""
    from b200vc import modules, ops
    o, p = _pair(192, False, trained_like=True)
    params = modules.gdn_params(p)
    x = _x(2, 192, 272, 480)
    with torch.no_grad():
        want = o(x)
    got = ops.gdn(x, params, impl=2)
    err = ((got - want).abs() / want.abs().clamp(min=1e-6)).max().item()
    print(f"gdn C=192 272x480 (tcgen05): max rel err {err:.3e}")
    assert err < 1e-5
    assert torch.equal(ops.gdn(x, params, impl=2), got)
    assert torch.equal(ops.gdn(x[1:2].contiguous(), params, impl=2), got[1:2])
    assert torch.equal(ops.gdn(x.flip(-1).contiguous(), params, impl=2), got.flip(-1))
    assert torch.equal(ops.gdn(x, params), got)                       # auto == tcgen05 for C = 192
    exact = ops.gdn(x, params, impl=1)
    assert ((got - exact).abs() <= 1e-5 * exact.abs() + 1e-12).all()


def test_gdn_unsupported_channels_raise():
    from b200vc import ops
    x = torch.randn(1, 96, 4, 4, device="cuda")
    params = torch.zeros(96 + 4 * 96 * 96, device="cuda")
    with pytest.raises(RuntimeError, match="unsupported channel count"):
        ops.gdn(x, params)


@pytest.mark.parametrize("impl", [1, 2])
def test_gdn_full_size_properties(strict_fp32, impl):
    """Size-independent properties at the bench's largest shape ([2,128,544,960]), per implementation:
    GDN acts per spatial position, so (a) it commutes bit for bit with any permutation of the positions (here a flip
    and a roll, which move every position into another tile / lane / pipeline stage), (b) it is batch-invariant bit
    for bit, and (c) IGDN(GDN(x)) == x * sqrt(norm(GDN(x))) / sqrt(norm(x)) -- checked against the norm kernel."""
    from b200vc import modules, ops
    _, p = _pair(128, False, trained_like=True)
    params = modules.gdn_params(p)
    x = _x(2, 128, 544, 960)
    y = ops.gdn(x, params, impl=impl)
    assert torch.equal(ops.gdn(x.flip(-1).contiguous(), params, impl=impl), y.flip(-1))
    assert torch.equal(ops.gdn(x.roll(shifts=(37, 411), dims=(2, 3)), params, impl=impl), y.roll(shifts=(37, 411), dims=(2, 3)))
    assert torch.equal(ops.gdn(x[1:2].contiguous(), params, impl=impl), y[1:2])
    # out-of-place residual form == plain + addend, one rounding apart at most
    skip = _x(2, 128, 544, 960)
    fused = ops.gdn(x, params, addend=skip.clone(), impl=impl)
    assert ((fused - (y + skip)).abs() <= 2e-6 * (y.abs() + skip.abs())).all()
    # inverse direction is the exact reciprocal factor of the forward one on the same input
    z = ops.gdn(x, params, inverse=True, impl=impl)
    prod = (y.double() * z.double())
    assert ((prod - x.double() ** 2).abs() <= 1e-5 * x.double() ** 2 + 1e-30).all()


def test_residual_block_stride1_keeps_its_input(strict_fp32):
    """with stride 1 and in_ch == out_ch the identity IS the caller's input; the fused residual add must not
    write into it (CompressAI's `out += identity` writes `out`)."""
    from b200vc import modules
    torch.manual_seed(3)
    blk_o = cai.ResidualBlockWithStride(128, 128, stride=1).cuda().eval()
    blk_p = modules.ResidualBlockWithStride(128, 128, stride=1).cuda().eval()
    assert blk_p.skip is None
    blk_p.load_state_dict(blk_o.state_dict())
    x = _x(1, 128, 24, 36)
    keep = x.clone()
    with torch.no_grad():
        got = blk_p(x)
        want = blk_o(keep.clone())
    assert torch.equal(x, keep), "the block overwrote its input"
    assert ((got - want).abs() <= 1e-5 * want.abs() + 1e-6).all()


def test_data_edits_need_and_get_cache_invalidation(strict_fp32):
    """`.data` edits do not bump the tensor version: `invalidate_caches` (also run by update() and after
    load_state_dict) drops the derived GDN operands."""
    from b200vc import modules
    o, p = _pair(128, False, trained_like=True)
    x = _x(1, 128, 8, 12)
    with torch.no_grad():
        a = p(x)
        p.beta.data.mul_(1.5)
        o.beta.data.mul_(1.5)
        modules.invalidate_caches(p)
        b = p(x)
        assert not torch.equal(a, b)
        torch.testing.assert_close(b, o(x), rtol=1e-5, atol=1e-7)
        sd = {k: v.clone() for k, v in p.state_dict().items()}
        sd["beta"] = sd["beta"] * 0.5
        p.load_state_dict(sd)
        o.load_state_dict(sd)
        torch.testing.assert_close(p(x), o(x), rtol=1e-5, atol=1e-7)
