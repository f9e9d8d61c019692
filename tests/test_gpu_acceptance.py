"""T1 -- the north star's own acceptance run (BASELINE.json): one synthetic 1080p GOP-8
(``[1, 9, 3, 1088, 1920]`` after the reference's reflection pad) through ``GopCoder(b200vc.Model)`` against the oracle
restatement of ``LHBDC/model/m.py:32-98`` driven in the *same level batches* on the same GPU, strict fp32
(reference loop: ``LHBDC/test/testing.py:167-186``).

Bars (north star): quantised symbols / CDF indexes bit-exact apart from documented round-half ties, per-frame
estimated bits and GOP bpp within 1e-4 relative.

Two views:

* **free-running**: both codecs run the whole hierarchy from the same two anchors; nothing is shared afterwards.
  This is the acceptance statement itself (bits / bpp / PSNR of the GOP).
* **stage-wise**: every quantiser of every frame is compared on identical inputs (the oracle stage is fed what the
  product stage was fed), so a single flipped symbol cannot cascade and mask every later comparison; each flip is
  then shown to be a round-half near-tie of the un-quantised value.

Both views run for the exact-fp32 GDN kernel (``impl=1``: every symbol must be equal) and for the default tcgen05
3xTF32 GDN kernel (flips counted, bounded and proven to be near-ties).
"""
import math

import pytest
import torch

from gpu_util import build_models

pytestmark = pytest.mark.gpu

H, W = 1080, 1920
# A flip is "a documented round-half tie" when the un-quantised value sits this close to k + 0.5 (absolute, in
# quantisation steps, relative to the magnitude of the value for large latents): three orders of magnitude
# below the quantisation step, three above the kernels' 1e-6 relative error.
TIE_TOL = 1e-3


@pytest.fixture(scope="module")
def setup(strict_fp32):
    from b200vc import synthetic
    from b200vc.lhbdc import reflect_pad64
    orc, prod = build_models("cuda")
    frames = reflect_pad64(synthetic.make_sequence(9, H, W, seed=1234, device="cuda"))
    assert tuple(frames.shape) == (9, 3, 1088, 1920)
    return orc, prod, frames


def _per_sample_bits(result):
    return sum((-torch.log2(l.double())).sum(dim=(1, 2, 3)) for l in result["likelihoods"].values())


@torch.no_grad()
def oracle_gop(orc, frames, sched, crop):
    """The oracle through the same level batches as ``GopCoder.code`` (b200vc/gop.py), free-running."""
    from b200vc import ops
    T = sched.gop + 1
    decoded = {0: frames[0:1], sched.gop: frames[sched.gop:sched.gop + 1]}
    bits = torch.zeros(T, dtype=torch.float64, device=frames.device)
    sse = torch.zeros(T, dtype=torch.float64, device=frames.device)
    for level_frames in sched.by_level():
        xb = torch.cat([decoded[sched.refs[f][0]] for f in level_frames], 0)
        xa = torch.cat([decoded[sched.refs[f][1]] for f in level_frames], 0)
        xc = torch.cat([frames[f:f + 1] for f in level_frames], 0)
        x_hat, _, _, parts = orc(xb, xc, xa, train=False, return_parts=True)
        b = _per_sample_bits(parts["flow_result"]) + _per_sample_bits(parts["residual_result"])
        h, w = crop
        u8 = lambda t: torch.round(t[:, :, :h, :w].clamp(0, 1) * 255.0)
        s = ((u8(x_hat) - u8(xc)).double() ** 2).sum(dim=(1, 2, 3))
        for k, f in enumerate(level_frames):
            decoded[f], bits[f], sse[f] = x_hat[k:k + 1], b[k], s[k]
        del parts
    return bits, sse, decoded


@pytest.mark.parametrize("impl", [1, 0], ids=["gdn_exact_fp32", "gdn_tcgen05_default"])
def test_gop8_1080p_free_running_bits_and_bpp(setup, impl, monkeypatch):
    from b200vc import gop, ops
    orc, prod, frames = setup
    monkeypatch.setattr(ops, "_GDN_IMPL", impl)
    sched = gop.LHBDC_GOP8
    bits, sse, dec = gop.GopCoder(prod, sched).code(frames[None], (H, W), want_decoded=True)
    bits_o, sse_o, dec_o = oracle_gop(orc, frames, sched, (H, W))
    worst = 0.0
    exact = impl == 1
    for f in sched.order:
        rel = abs(bits[0, f].item() - bits_o[f].item()) / bits_o[f].item()
        dx = (dec[0, f] - dec_o[f][0]).abs()
        psnr = gop.psnr_from_sse(sse[0, f].cpu(), 3 * H * W).item()
        psnr_o = gop.psnr_from_sse(sse_o[f].cpu(), 3 * H * W).item()
        print(f"  frame {f} level {sched.levels[f]}: bits oracle {bits_o[f].item():.2f} kernels {bits[0, f].item():.2f} "
              f"rel {rel:.2e}; PSNR {psnr_o:.4f} / {psnr:.4f} dB; x_hat max|d| {dx.max().item():.2e} "
              f"frac>1e-3 {(dx > 1e-3).float().mean().item():.2e}")
        worst = max(worst, rel)
        assert abs(psnr - psnr_o) < 1e-2
        if exact:
            assert dx.max().item() == 0.0, "exact-fp32 GDN: the decoded frames must be identical"
    bpp = bits[0].sum().item() / (7 * H * W)
    bpp_o = bits_o.sum().item() / (7 * H * W)
    rel_bpp = abs(bpp - bpp_o) / bpp_o
    print(f"GOP-8 1080p (GDN impl {impl}): bpp oracle {bpp_o:.6f} kernels {bpp:.6f} rel {rel_bpp:.2e}; "
          f"worst per-frame bits rel {worst:.2e}")
    # North-star bar: total estimated bpp within 1e-4 relative.  With the exact-fp32 GDN kernel every frame is also
    # within 1e-4 (measured 1.4e-8: the codecs are identical).  With the default tcgen05 GDN (~1e-6 relative) a handful
    # of round-half near-ties per 10^6 symbols flip (proven to be near-ties by the stage-wise test below); free-running,
    # a flip in a level-0 frame perturbs the references of the deeper levels, and with RANDOM weights (a 9 dB "codec")
    # that perturbation is amplified, not damped -- per-frame bits of level-2 frames then move by a few 1e-4 while the
    # GOP total stays inside the bar (measured 8e-6).
    assert rel_bpp < 1e-4
    assert worst < (1e-4 if exact else 1e-3)


def _tie_distance(v):
    """Distance of ``v`` from the nearest k + 0.5 (0 = exactly on a rounding boundary)."""
    return (v - torch.floor(v) - 0.5).abs()


def _compare_symbols(name, got, want, raw, stats, strict):
    """Equal, or differing only where the un-quantised oracle value ``raw`` is a round-half near-tie."""
    bad = got != want
    n_bad = int(bad.sum().item())
    stats["symbols"] += got.numel()
    stats["flips"] += n_bad
    if n_bad:
        d = _tie_distance(raw[bad])
        scale = raw[bad].abs().clamp(min=1.0)
        worst = (d / scale).max().item()
        step = (got[bad] - want[bad]).abs().max().item()
        stats["worst_tie"] = max(stats["worst_tie"], worst)
        print(f"    {name}: {n_bad} of {got.numel()} differ; all |delta| = {step}; "
              f"max distance from a .5 boundary {worst:.2e} (relative to max(1,|v|))")
        assert step == 1, f"{name}: a symbol moved by more than one step"
        assert worst < TIE_TOL, f"{name}: a differing symbol is not a round-half near-tie ({worst:.3e})"
    if strict:
        assert n_bad == 0, f"{name}: {n_bad} symbols differ with the exact-fp32 GDN kernel"


@torch.no_grad()
def _oracle_compressor_stages(comp, x, z_hat_forced):
    """Oracle view of ``compress``'s tensor half (LHBDC/model/layers.py:93-104) on input ``x``; the Gaussian
    parameters come from ``z_hat_forced`` (the product's z_hat) so the y comparison is not hostage to a z tie."""
    y = comp.g_a(x)
    z = comp.h_a(y)
    med = comp.entropy_bottleneck._get_medians().reshape(1, -1, 1, 1)
    z_sym = torch.round(z - med).int()
    scales, means = comp.h_s(z_hat_forced).chunk(2, 1)
    idx = comp.gaussian_conditional.build_indexes(scales)
    y_sym = comp.gaussian_conditional.quantize(y, "symbols", means)
    res = comp(x)
    return {"y": y, "z": z, "med": med, "z_symbols": z_sym, "y_symbols": y_sym, "y_indexes": idx, "means": means,
            "scales": scales, "bits": _per_sample_bits(res)}


@pytest.mark.parametrize("impl", [1, 0], ids=["gdn_exact_fp32", "gdn_tcgen05_default"])
def test_gop8_1080p_symbols_stage_by_stage(setup, impl, monkeypatch):
    from b200vc import gop, ops
    from oracle import warp as o_warp
    orc, prod, frames = setup
    monkeypatch.setattr(ops, "_GDN_IMPL", impl)
    sched = gop.LHBDC_GOP8
    strict = impl == 1
    decoded = {0: frames[0:1], 8: frames[8:9]}
    stats = {"symbols": 0, "flips": 0, "worst_tie": 0.0}
    worst_bits = 0.0
    with torch.no_grad():
        for level_frames in sched.by_level():
            xb = torch.cat([decoded[sched.refs[f][0]] for f in level_frames], 0)
            xa = torch.cat([decoded[sched.refs[f][1]] for f in level_frames], 0)
            xc = torch.cat([frames[f:f + 1] for f in level_frames], 0)
            x_hat, bits, p = prod.forward_device(xb, xc, xa, return_parts=True)
            print(f"  level frames {level_frames}:")
            # -- motion front end (SPyNet x4 + pooling + pad + difference): same torch convs, kernel glue bit-exact
            diff_o, fab_o, fba_o, hh, ww = orc.motion(xb, xc, xa)
            e = (p["diff_flow"] - diff_o).abs().max().item() / diff_o.abs().max().item()
            print(f"    motion: max|d diff_flow| / max|diff_flow| = {e:.2e}")
            assert e < 1e-5
            # -- motion-vector compressor on the product's own input
            so = _oracle_compressor_stages(orc.mv_compressor, p["diff_flow"], p["mv"]["z_hat"])
            _compare_symbols("mv z_symbols", p["mv"]["z_symbols"], so["z_symbols"], so["z"] - so["med"], stats, strict)
            _compare_symbols("mv y_symbols", p["mv"]["y_symbols"], so["y_symbols"], so["y"] - so["means"], stats, strict)
            assert torch.equal(p["mv"]["y_indexes"], so["y_indexes"])
            rel = ((p["bits_flow"] - so["bits"]).abs() / so["bits"]).max().item()
            # -- flow glue + both warps + concat, mask, blend on the product's decoded flow
            cb, ca = o_warp.lhbdc_flow_glue(p["flow_hat"], p["flow_ab"], p["flow_ba"], hh, ww)
            warped_o = torch.cat([orc.backwarp(xb, cb), orc.backwarp(xa, ca)], 1)
            assert torch.equal(p["warped"], warped_o), "fused glue + warps are not bit-exact at 1088x1920"
            mask_o = orc.masknet(p["warped"])
            assert torch.equal(p["mask"], mask_o)
            pred_o, res_o = o_warp.blend_residual_lhbdc(mask_o, warped_o[:, 0:3], warped_o[:, 3:6], xc)
            assert torch.equal(p["pred"], pred_o) and torch.equal(p["residual"], res_o)
            # -- residual compressor on the product's own residual
            sr = _oracle_compressor_stages(orc.residual_compressor, p["residual"], p["res"]["z_hat"])
            _compare_symbols("res z_symbols", p["res"]["z_symbols"], sr["z_symbols"], sr["z"] - sr["med"], stats, strict)
            _compare_symbols("res y_symbols", p["res"]["y_symbols"], sr["y_symbols"], sr["y"] - sr["means"], stats, strict)
            assert torch.equal(p["res"]["y_indexes"], sr["y_indexes"])
            rel = max(rel, ((p["bits_residual"] - sr["bits"]).abs() / sr["bits"]).max().item())
            print(f"    per-frame bits of both compressors vs oracle on the same inputs: worst rel {rel:.2e}")
            worst_bits = max(worst_bits, rel)
            for k, f in enumerate(level_frames):
                decoded[f] = x_hat[k:k + 1]
            del p, so, sr
    frac = 1.0 - stats["flips"] / stats["symbols"]
    print(f"GOP-8 1080p stage-wise (GDN impl {impl}): {stats['flips']} of {stats['symbols']} symbols differ "
          f"(equal fraction {frac:.9f}); worst tie distance {stats['worst_tie']:.2e}; worst bits rel {worst_bits:.2e}")
    assert worst_bits < 1e-4
    # default tcgen05 path: ~1e-6 relative GDN error on latents of magnitude 1..100 => a few round-half near-ties per
    # 10^6 symbols may fall on the other side (measured: 4 in the 1 044 480 residual symbols of frame 4)
    assert stats["flips"] <= (0 if strict else 8 * max(1, stats["symbols"] // 1_000_000))
