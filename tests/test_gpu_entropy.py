"""GPU parity: K-GC / K-EB / bit sums through the C-ABI vs the oracle restatement of CompressAI's
GaussianConditional / EntropyBottleneck run on the same device.
Bars (north star): symbols / indexes / dequantised values bit-exact; likelihoods within 1e-5 relative;
bit totals within 1e-4 relative (we hold 1e-6 against the order-independent fp64 total)."""
import math

import pytest
import torch

from oracle import cai
from oracle import warp as o_warp
from gpu_util import gc_case, rel_err

pytestmark = pytest.mark.gpu

# The factorised-prior likelihood is |sigmoid(u) - sigmoid(l)|: a difference of nearly equal numbers, so a
# 1-ulp difference in a logit is amplified.  The oracle's 3x3 products run through torch.matmul (cuBLAS), whose
# kernel -- and accumulation order -- changes with the problem shape: at the 1080p shape [1,128,17,30] the
# kernel matches it bit for bit, at tiny shapes it differs by up to ~1e-5 relative.  Bar: 5e-5 relative on
# the likelihood (SURVEY Appendix C.7 accepts formula-level agreement here); bits totals are held to 1e-6.
EB_TOL = 5e-5


def _gc():
    from b200vc import modules
    o = cai.GaussianConditional(None).eval()
    o.update_scale_table(cai.get_scale_table())
    p = modules.GaussianConditional(None).eval()
    p.update_scale_table(modules.get_scale_table())
    return o.cuda(), p.cuda()


@pytest.mark.parametrize("shape", [(1, 128, 68, 120), (1, 128, 20, 32), (2, 6, 9, 7), (3, 1, 1, 1), (1, 80, 68, 120)])
def test_gauss_cond_forward(shape):
    from b200vc import ops
    o, p = _gc()
    y, sigma, mu = gc_case(sum(shape), *shape)
    with torch.no_grad():
        y_hat_o, lik_o = o(y, sigma, means=mu)
    y_hat, lik = p(y, sigma, means=mu)
    assert torch.equal(y_hat, y_hat_o)
    err = rel_err(lik, lik_o)
    print(f"gauss_cond {shape}: lik max rel err {err:.3e}, bit-exact {(lik == lik_o).float().mean().item():.5f}")
    assert err < 1e-5
    assert lik.min().item() >= float(torch.tensor(1e-9, dtype=torch.float32))
    r = ops.gauss_cond(y, sigma, mu, want_y_hat=False, want_lik=False, want_bits=True)
    for n in range(shape[0]):
        want64 = o_warp.bits_fp64(lik_o[n]).item()
        assert abs(r["bits"][n].item() - want64) / want64 < 1e-6
    ref32 = o_warp.bits_fp32(lik_o).item()  # the reference's own fp32 tree sum
    assert abs(r["bits"].sum().item() - ref32) / ref32 < 1e-4


def test_gauss_cond_on_channel_chunks_and_no_means():
    """scales/means arrive as the two channel chunks of h_s's output (LHBDC/model/layers.py:100-101)."""
    o, p = _gc()
    g = torch.Generator().manual_seed(5)
    params = torch.randn(2, 32, 10, 14, generator=g).cuda() * 3
    scales, means = params.chunk(2, 1)
    y = torch.randn(2, 16, 10, 14, generator=g).cuda() * 4
    with torch.no_grad():
        yo, lo = o(y, scales, means=means)
        y2, l2 = o(y, scales.abs())
    yp, lp = p(y, scales, means=means)
    assert torch.equal(yp, yo) and rel_err(lp, lo) < 1e-5
    yq, lq = p(y, scales.abs())
    assert torch.equal(yq, y2) and rel_err(lq, l2) < 1e-5


@pytest.mark.parametrize("shape", [(1, 128, 68, 120), (2, 5, 7, 3)])
def test_symbols_and_indexes_are_bit_exact(shape):
    o, p = _gc()
    y, sigma, mu = gc_case(77 + sum(shape), *shape)
    sigma.view(-1)[:64] = o.scale_table  # exact table hits (the compare is inclusive)
    y.view(-1)[64:70] = mu.view(-1)[64:70] + torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, -2.5], device="cuda")
    with torch.no_grad():
        sym_o = o.quantize(y, "symbols", mu)
        idx_o = o.build_indexes(sigma)
    assert torch.equal(p.quantize(y, "symbols", mu), sym_o)
    assert torch.equal(p.build_indexes(sigma), idx_o)
    assert torch.equal(p.quantize(y, "dequantize", mu), o.quantize(y, "dequantize", mu))
    assert idx_o.min().item() == 0 and idx_o.max().item() == 63


def test_gauss_cond_inverse_gain_epilogue():
    """Flex-Rate: y_hat leaves through inv_gain_unit (b_model/layers.py:146) -- fused as an epilogue."""
    from b200vc import ops
    o, _ = _gc()
    y, sigma, mu = gc_case(9, 2, 16, 12, 20)
    gain = (1 + 0.1 * torch.randn(16, device="cuda")).abs()
    with torch.no_grad():
        y_hat_o, lik_o = o(y, sigma, means=mu)
    r = ops.gauss_cond(y, sigma, mu, inv_gain=gain)
    assert torch.equal(r["y_hat"], gain.view(1, -1, 1, 1) * y_hat_o)
    assert rel_err(r["lik"], lik_o) < 1e-5


def _eb(C, perturbed):
    from b200vc import modules
    g = torch.Generator().manual_seed(3)
    o = cai.EntropyBottleneck(C).eval()
    if perturbed:
        with torch.no_grad():
            for i in range(5):
                m = getattr(o, f"_matrix{i}")
                m.add_(0.3 * torch.randn(m.shape, generator=g))
                if i < 4:
                    getattr(o, f"_factor{i}").copy_(0.5 * torch.randn(getattr(o, f"_factor{i}").shape, generator=g))
            o.quantiles[:, 0, 1] = 0.7 * torch.randn(C, generator=g)
    p = modules.EntropyBottleneck(C).eval()
    p.load_state_dict(o.state_dict())
    return o.cuda(), p.cuda()


@pytest.mark.parametrize("perturbed", [False, True])
@pytest.mark.parametrize("shape", [(1, 128, 17, 30), (1, 128, 5, 8), (2, 7, 3, 5)])
def test_entropy_bottleneck_forward(shape, perturbed):
    from b200vc import modules, ops
    o, p = _eb(shape[1], perturbed)
    z = (3.0 * torch.randn(*shape, generator=torch.Generator().manual_seed(1))).cuda()
    with torch.no_grad():
        z_hat_o, lik_o = o(z)
    z_hat, lik = p(z)
    assert torch.equal(z_hat, z_hat_o)
    err = rel_err(lik, lik_o)
    print(f"entropy_bottleneck {shape} perturbed={perturbed}: lik max rel err {err:.3e}")
    assert err < EB_TOL
    r = ops.entropy_bottleneck(z, modules.eb_packed(p), want_lik=False, want_symbols=True)
    med = o.quantiles[:, 0, 1].view(1, -1, 1, 1)
    assert torch.equal(r["symbols"], torch.round(z - med).int())
    for n in range(shape[0]):
        want64 = o_warp.bits_fp64(lik_o[n]).item()
        assert abs(r["bits"][n].item() - want64) / want64 < 1e-6


def test_entropy_bottleneck_gain_prologue_and_epilogue():
    """Flex-Rate hyper_gain_unit / hyper_inv_gain_unit (b_model/layers.py:140-143) fused around the EB."""
    from b200vc import modules, ops
    o, p = _eb(16, True)
    g = torch.Generator().manual_seed(12)
    z = (2.0 * torch.randn(2, 16, 6, 10, generator=g)).cuda()
    gain = (1 + 0.1 * torch.randn(16, generator=g)).abs().cuda()
    inv = (1 + 0.1 * torch.randn(16, generator=g)).abs().cuda()
    with torch.no_grad():
        z_hat_o, lik_o = o(gain.view(1, -1, 1, 1) * z)
    r = ops.entropy_bottleneck(z, modules.eb_packed(p), gain=gain, inv_gain=inv)
    assert torch.equal(r["z_hat"], inv.view(1, -1, 1, 1) * z_hat_o)
    assert rel_err(r["lik"], lik_o) < EB_TOL


def test_bit_sums_are_deterministic_and_shape_only():
    from b200vc import ops
    y, sigma, mu = gc_case(21, 1, 128, 68, 120)
    a = ops.gauss_cond(y, sigma, mu, want_y_hat=False, want_lik=False)["bits"]
    b = ops.gauss_cond(y, sigma, mu, want_y_hat=False, want_lik=False)["bits"]
    assert torch.equal(a, b)
    # batching does not change a sample's total (per-sample partial layout depends on C*HW only)
    y2, s2, m2 = (torch.cat([t, t.flip(0)], 0) for t in (y, sigma, mu))
    c = ops.gauss_cond(y2, s2, m2, want_y_hat=False, want_lik=False)["bits"]
    assert c[0].item() == a[0].item() and c[1].item() == a[0].item()


# ---- round 2: persistent K-GC (partial slots walked by resident CTAs, per-thread cp.async ring, packed fp32) ----------
@pytest.mark.parametrize("shape", [(1, 1, 1, 4), (3, 1, 1, 3), (2, 7, 5, 9), (2, 8, 16, 8), (5, 3, 33, 31), (1, 16, 17, 60),
                                   (300, 2, 2, 2), (2, 128, 20, 32), (1, 5, 205, 4), (3, 1, 64, 64), (1, 1, 64, 65),
                                   (2, 12, 34, 60)])
def test_gauss_cond_slot_geometry_edge_shapes(shape):
    """Vector (HW % 4 == 0) and scalar path, one slot per sample, hundreds of samples per CTA, partial last chunks, sizes
    around the 1024-element chunk and 4096-element slot boundaries; one element at the scale bound, one in erfc's far
    tail (the guarded form).  Every output is bit-equal to the oracle chain; both kernel forms give the same totals."""
    from b200vc import ops
    o, _ = _gc()
    N, C, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    r = lambda: torch.randn(N, C, H, W, generator=g).cuda()
    y, s, m = 4 * r(), r().abs() * 3, r()
    s[0, 0, 0, 0] = 0.0
    y.view(-1)[-1] = 60.0
    table = o.scale_table
    out = ops.gauss_cond(y, s, m, want_lik=True, want_bits=True, want_symbols=True, scale_table=table)
    lean = ops.gauss_cond(y, s, m, want_lik=False, want_bits=True)
    with torch.no_grad():
        yh, lik = o(y, s, means=m)
    assert torch.equal(out["y_hat"], yh) and torch.equal(lean["y_hat"], yh)
    assert torch.equal(out["lik"], lik)
    assert torch.equal(out["symbols"], torch.round(yh - m).int())
    assert torch.equal(out["indexes"], o.build_indexes(s))
    ref = -torch.log2(lik.double()).flatten(1).sum(1)
    for res in (out, lean):
        assert ((res["bits"] - ref).abs() <= 1e-6 * ref.abs() + 1e-6).all()


def test_gauss_cond_operands_near_the_exponent_limits_take_the_scalar_routines():
    """sigma or |y - mu| beyond 2^59 leave the shared-reciprocal division's proven range: the kernel must fall back to
    __fdiv_rn / erfcf element by element and still equal the oracle bit for bit."""
    from b200vc import ops
    o, _ = _gc()
    g = torch.Generator().manual_seed(11)
    y = (3 * torch.randn(1, 4, 8, 16, generator=g)).cuda()
    s = (torch.rand(1, 4, 8, 16, generator=g) * 2).cuda()
    m = torch.randn(1, 4, 8, 16, generator=g).cuda()
    s[0, 1, 2, 3] = 3.0e19          # > 2^59
    y[0, 2, 5, 7] = -2.5e30
    s[0, 3, 0, 0] = float("inf")
    with torch.no_grad():
        yh, lik = o(y, s, means=m)
    out = ops.gauss_cond(y, s, m, want_lik=True, want_bits=True)
    assert torch.equal(out["y_hat"], yh)
    assert torch.equal(out["lik"], lik)


def test_packed_fp32_likelihood_arithmetic_is_bit_identical_to_libdevice():
    """tools/gc_math_check: erfc2 (csrc/gc_math.cuh) against erfcf over ALL 2^32 arguments in every template form, the
    shared-reciprocal division against __fdiv_rn over 2e10 in-range quotients.  0 mismatches, exit code 0."""
    import json
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tools", "_bin", "gc_math_check")
    if not os.path.exists(exe):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                        os.path.join(root, "tools", "gc_math_check.cu")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    print(rep)
    assert r.returncode == 0
    assert rep["erfc2_mismatch"] == 0 and rep["div2_mismatch_random"] == 0 and rep["div2_mismatch_gc_shaped"] == 0
