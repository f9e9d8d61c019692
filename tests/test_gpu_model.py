"""GPU parity at module level: the LHBDC model mirror (kernels swapped in) vs the oracle model (torch ops), same
calibrated weights, TF32 disabled; plus the golden outputs of the reference's own ``Model.forward``.
Quantisers sit between the compared tensors, so parity is stated as (a) symbols identical up to rare
round-half ties caused by last-ulp upstream differences, (b) total bits within 1e-4 relative (north star)."""
import copy
import os

import numpy as np
import pytest
import torch

from gpu_util import build_models

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def models(strict_fp32):
    return build_models("cuda")


@pytest.fixture(scope="module")
def triple(golden_dir):
    gold = np.load(os.path.join(golden_dir, "lhbdc_model_reference.npz"))
    tri = torch.from_numpy(gold["synthetic_u8"]).cuda().float() / 255.0
    return gold, (tri[0:1], tri[1:2], tri[2:3])


def test_spynet_with_kernel_warp(models, triple):
    orc, prod = models
    _, (xb, xc, xa) = triple
    with torch.no_grad():
        want = orc.FlowNet(xc, xb)
        got = prod.FlowNet(xc, xb)
    err = (got - want).abs().max().item()
    print(f"SPyNet flow max|diff| = {err:.3e} px (|flow| max {want.abs().max().item():.2f})")
    assert err < 1e-3


@pytest.mark.parametrize("which,ch", [("residual_compressor", 3), ("mv_compressor", 4)])
def test_hyperprior_teacher_forced(models, which, ch):
    orc, prod = models
    o, p = getattr(orc, which), getattr(prod, which)
    g = torch.Generator().manual_seed(2)
    x = (0.3 * torch.randn(2, ch, 128, 192, generator=g)).cuda()
    with torch.no_grad():
        ro = o(x)
        rp = p(x)
        so = o.symbols(x)
        sp = p.symbols(x)
        x_hat_b, by, bz = p.forward_bits(x)
    for k in ("y_symbols", "y_indexes", "z_symbols"):
        same = (so[k] == sp[k]).float().mean().item()
        print(f"{which} {k}: identical fraction {same:.6f}")
        assert same == 1.0  # north star: bit-matching symbols (measured 1.000000 on the default tcgen05 GDN path too)
    assert tuple(sp["shape"]) == tuple(so["shape"])
    bits_o = sum((-torch.log2(l.double())).sum().item() for l in ro["likelihoods"].values())
    bits_p = sum((-torch.log2(l.double())).sum().item() for l in rp["likelihoods"].values())
    print(f"{which}: bits oracle {bits_o:.3f} kernels {bits_p:.3f} rel {abs(bits_p - bits_o) / bits_o:.2e}")
    assert abs(bits_p - bits_o) / bits_o < 1e-4
    assert abs((by + bz).sum().item() - bits_p) / bits_p < 1e-6  # bits-only pass == materialised likelihoods
    assert torch.equal(x_hat_b, rp["x_hat"])
    close = ((rp["x_hat"] - ro["x_hat"]).abs() < 1e-4).float().mean().item()
    assert close > 0.9999, close


def test_model_forward_matches_oracle(models, triple):
    orc, prod = models
    _, (xb, xc, xa) = triple
    with torch.no_grad():
        x_o, rate_o, size_o, parts = orc(xb, xc, xa, train=False, return_parts=True)
        x_p, rate_p, size_p = prod(xb, xc, xa, train=False)
        _, bits, pp = prod.forward_device(xb, xc, xa)
    rel = abs(size_p - parts["size64"]) / parts["size64"]
    print(f"Model.forward: size oracle {size_o:.2f} (fp64 {parts['size64']:.2f}) kernels {size_p:.2f} rel {rel:.2e}")
    assert rel < 1e-4
    assert abs(rate_p.item() - rate_o.item()) / rate_o.item() < 1e-4
    assert isinstance(size_p, float) and rate_p.dtype == torch.float32
    close = ((x_p - x_o).abs() < 1e-4).float().mean().item()
    print(f"x_hat within 1e-4: {close:.5f}; max|diff| {(x_p - x_o).abs().max().item():.3e}")
    assert close > 0.9999
    assert abs(bits.sum().item() - size_p) < 1e-6 * size_p


def test_model_forward_matches_reference_golden(models, triple):
    """Golden = the reference's own LHBDC/model/m.py run (CPU, through the compressai stand-in)."""
    _, prod = models
    gold, (xb, xc, xa) = triple
    with torch.no_grad():
        x_p, rate_p, size_p = prod(xb, xc, xa, train=False)
    want = torch.from_numpy(gold["synthetic_x_hat"]).cuda()
    rel = abs(size_p - float(gold["synthetic_size64"])) / float(gold["synthetic_size64"])
    close = ((x_p - want).abs() < 2e-3).float().mean().item()
    print(f"vs reference golden: size rel {rel:.2e}, x_hat within 2e-3: {close:.5f}")
    # golden = CPU run of the reference's m.py (oneDNN convolutions), this = cuDNN: bits to the north-star bar,
    # x_hat to the conv-backend noise amplified by the synthesis transform
    assert rel < 1e-4 and close > 0.999


def test_patch_swaps_kernels_into_a_foreign_model(models, triple):
    """patch() on a model built from other classes (here: the oracle's) == the product mirror, bit for bit."""
    import b200vc
    orc, prod = models
    _, (xb, xc, xa) = triple
    foreign = b200vc.patch(copy.deepcopy(orc))
    with torch.no_grad():
        x_f, rate_f, size_f = foreign(xb, xc, xa, train=False)
        x_p, rate_p, size_p = prod(xb, xc, xa, train=False)
        assert torch.equal(x_f, x_p) and size_f == size_p
        assert torch.equal(foreign.backwarp(xb, torch.ones(1, 2, 192, 192, device="cuda")),
                           prod.backwarp(xb, torch.ones(1, 2, 192, 192, device="cuda")))
        # operator-level swap only (the foreign forward keeps running): still on the kernels, same answer class
        unfused = b200vc.patch(copy.deepcopy(orc), fuse=False)
        x_u, _, size_u = unfused(xb, xc, xa, train=False)
    assert abs(size_u - size_p) / size_p < 1e-4


def test_encode_b_symbols(models, triple):
    import b200vc
    from oracle import lhbdc as o_lhbdc
    orc, prod = models
    _, (xb, xc, xa) = triple
    with torch.no_grad():
        mo, ro = o_lhbdc.encode_B_symbols(orc, xa, xc, xb)
        mp, rp = b200vc.encode_B_symbols(prod, xa, xc, xb)
    for a, b, nm in ((mo, mp, "mv"), (ro, rp, "res")):
        for k in ("y_symbols", "y_indexes", "z_symbols"):
            same = (a[k] == b[k]).float().mean().item()
            print(f"encode_B {nm} {k}: identical fraction {same:.6f}")
            assert same == 1.0
        assert a[k].dtype == torch.int32 and b[k].dtype == torch.int32


def test_gop_coder_batches_levels_and_is_deterministic(models):
    from b200vc import gop, synthetic
    _, prod = models
    frames = synthetic.make_sequence(17, 192, 192, seed=5, device="cuda")
    gops = torch.stack([frames[0:9], frames[8:17]], 0)
    coder = gop.GopCoder(prod, gop.LHBDC_GOP8)
    bits, sse, dec = coder.code(gops, (180, 190), want_decoded=True)
    bits2, sse2 = coder.code(gops, (180, 190))
    assert torch.equal(bits, bits2) and torch.equal(sse, sse2)
    assert bits.shape == (2, 9) and (bits[:, 1:8] > 0).all() and (bits[:, [0, 8]] == 0).all()
    assert torch.isfinite(dec).all() and (sse[:, 1:8] > 0).all()
    # a frame coded alone gives the same bits as inside the batch (up to cuDNN algorithm choice by batch size)
    with torch.no_grad():
        _, b4, _ = prod.forward_device(gops[1:2, 0], gops[1:2, 4], gops[1:2, 8])
    assert abs(b4.item() - bits[1, 4].item()) / b4.item() < 1e-4


def test_forward_device_is_cuda_graph_capturable(models, triple):
    """the library never allocates, synchronises or keeps mutable state, so the whole B-frame step can
    be captured once and replayed on new frames (tensor maps and workspaces are baked in at capture time)."""
    _, prod = models
    _, (xb, xc, xa) = triple
    sb, sc, sa = xb.clone(), xc.clone(), xa.clone()
    with torch.no_grad():
        for _ in range(2):                                   # warm-up: lazy tables, cuDNN plans
            prod.forward_device(sb, sc, sa)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            x_hat, bits, _ = prod.forward_device(sb, sc, sa)
        # replay on different frames: swap the two references
        sb.copy_(xa)
        sa.copy_(xb)
        graph.replay()
        torch.cuda.synchronize()
        got_x, got_bits = x_hat.clone(), bits.clone()
        want_x, want_bits, _ = prod.forward_device(xa, xc, xb)
    assert torch.equal(got_bits, want_bits)
    assert torch.equal(got_x, want_x)


# ---- I-frame anchors: mbt2018_mean-shaped codec (LHBDC/test/testing.py:78-86,209) ----------------------------------
@pytest.mark.parametrize("quality", [3, 7])
def test_mbt2018_mean_anchor_codec_matches_oracle(strict_fp32, quality):
    """The product mirror (GDN C = 128 / 192 on tcgen05, K-EB, K-GC) vs the oracle's CompressAI-shaped
    MeanScaleHyperprior with the same random-init weights: image_compress (decoded image, bits)."""
    from b200vc import modules as M
    from b200vc import synthetic
    from oracle import cai
    torch.manual_seed(quality)
    N, Mm = M.MBT2018_MEAN_CFG[quality]
    orc = cai.MeanScaleHyperprior(N, Mm).eval()
    orc.keep_latents = True
    prod = M.mbt2018_mean(quality).eval()
    prod.load_state_dict(orc.state_dict())
    orc.update(force=True); prod.update(force=True)
    orc.cuda(); prod.cuda()
    x = synthetic.make_sequence(2, 128, 192, seed=3, device="cuda")
    with torch.no_grad():
        out = orc(x)
        want_bits = sum((torch.log(l.double()).flatten(1).sum(1) / -np.log(2.0)) for l in out["likelihoods"].values())
        x_hat, bits = M.image_compress(x, prod)
        api = prod(x)                                   # CompressAI-style result dict through the same kernels
    rel = ((bits - want_bits).abs() / want_bits).max().item()
    dx = (x_hat - out["x_hat"]).abs().max().item()
    sym_o = torch.round(out["latents"]["y_hat"] - out["latents"]["means_hat"])
    _, _, _, parts = prod.forward_bits(x, want_symbols=True)
    same = (parts["y_symbols"].float() == sym_o).float().mean().item()
    print(f"mbt2018_mean q={quality} (N={N}, M={Mm}): bits rel {rel:.2e}, x_hat max|d| {dx:.2e}, y symbols equal {same:.6f}")
    assert rel < 1e-4 and same > 0.9999
    assert dx < 1e-3 * max(1.0, out["x_hat"].abs().max().item())
    # two passes through cuDNN's transposed convolutions (not run-to-run deterministic in the last ulp)
    assert torch.allclose(api["x_hat"], x_hat, rtol=1e-4, atol=1e-5) and set(api["likelihoods"]) == {"y", "z"}


def test_gop_coder_codes_anchors_with_an_i_frame_codec(models):
    """GopCoder(anchor_codec=...): anchors go through image_compress first (testing.py:127-152), B-frames are predicted
    from the DECODED anchors, and the anchors' bits / SSE are booked."""
    from b200vc import gop, synthetic
    from b200vc import modules as M
    _, prod = models
    torch.manual_seed(1)
    icodec = M.mbt2018_mean(3).eval().cuda()
    icodec.update(force=True)
    frames = synthetic.make_sequence(17, 192, 192, seed=5, device="cuda")
    gops = torch.stack([frames[0:9], frames[8:17]], 0)
    coder = gop.GopCoder(prod, gop.LHBDC_GOP8, anchor_codec=icodec)
    bits, sse, dec = coder.code(gops, (180, 190), want_decoded=True)
    with torch.no_grad():
        d, s = M.image_compress(torch.cat([gops[:, 0], gops[:, 8]], 0), icodec)   # the coder's own batch composition
        d0, d8, s0, s8 = d[:2], d[2:], s[:2], s[2:]
        _, b4, _ = prod.forward_device(d0, gops[:, 4], d8)
    # (cuDNN's transposed convolutions are not run-to-run deterministic: last-ulp differences in the decoded anchors)
    assert torch.allclose(dec[:, 0], d0, rtol=1e-4, atol=1e-5) and torch.allclose(dec[:, 8], d8, rtol=1e-4, atol=1e-5)
    assert ((bits[:, 0] - s0).abs() / s0 < 1e-6).all() and ((bits[:, 8] - s8).abs() / s8 < 1e-6).all()
    assert (sse[:, [0, 8]] > 0).all() and (bits > 0).all()
    assert ((bits[:, 4] - b4).abs() / b4 < 1e-4).all()
    # the same GOPs without the codec: anchors uncoded, frame 4 predicted from the source frames
    plain, _ = gop.GopCoder(prod, gop.LHBDC_GOP8).code(gops, (180, 190))
    assert (plain[:, [0, 8]] == 0).all() and not torch.equal(plain[:, 4], bits[:, 4])
