"""GPU parity: the Flex-Rate model mirror (FLEX warp, NORMW blend, gain-fused entropy kernels, GDN) vs the oracle
restatement with identical weights, and vs the golden outputs of the reference's own BidirFlowRef."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def models(strict_fp32):
    from b200vc import flexrate, synthetic
    from oracle import flexrate as o_flex
    torch.manual_seed(0)
    orc = o_flex.BidirFlowRef(n=4, N=128).eval()
    synthetic.calibrate_flex_(orc, 0)
    prod = flexrate.BidirFlowRef(n=4, N=128).eval()
    prod.load_state_dict(orc.state_dict())
    return orc.cuda(), prod.cuda()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "flexrate_model_reference.npz"))


@pytest.mark.parametrize("tag,n,l", [("n1_l1", [1], 1.0), ("n0_l066", [0], 0.66)])
def test_flexrate_forward(models, gold, tag, n, l):
    orc, prod = models
    tri = torch.from_numpy(gold["frames_u8"]).cuda().float() / 255.0
    xb, xc, xa = tri[0:1], tri[1:2], tri[2:3]
    with torch.no_grad():
        want = orc(xb, xc, xa, n=n, l=l, train=False)
        got = prod(xb, xc, xa, n=n, l=l, train=False)
    rel = abs(got["size"].item() - want["size"].item()) / want["size"].item()
    close = ((got["x_hat"] - want["x_hat"]).abs() < 1e-3).float().mean().item()
    print(f"flexrate {tag}: size oracle {want['size'].item():.2f} kernels {got['size'].item():.2f} rel {rel:.2e}; "
          f"x_hat within 1e-3: {close:.5f}")
    assert rel < 1e-4 and close > 0.99
    assert got["size"].shape == (1,) and got["rate"].shape == (1,)
    rel_g = abs(got["size"].item() - float(gold[f"{tag}_size"][0])) / float(gold[f"{tag}_size"][0])
    close_g = ((got["x_hat"].cpu() - torch.from_numpy(gold[f"{tag}_x_hat"])).abs() < 2e-3).float().mean().item()
    print(f"   vs reference golden: size rel {rel_g:.2e}, x_hat within 2e-3: {close_g:.5f}")
    assert rel_g < 1e-3 and close_g > 0.98


def test_flexrate_compressor_api_and_batch(models):
    """flow_compressor(x, n, l) keeps the reference's result dict; per-sample sizes for N > 1."""
    orc, prod = models
    g = torch.Generator().manual_seed(4)
    x = (0.3 * torch.randn(2, 19, 64, 128, generator=g)).cuda()
    with torch.no_grad():
        ro = orc.flow_compressor(x, [2], 0.33, False)
        rp = prod.flow_compressor(x, [2], 0.33, False)
        _, by, bz = prod.flow_compressor.forward_bits(x, [2], 0.33)
    for k in ("y", "z"):
        bo = (-torch.log2(ro["likelihoods"][k].double())).sum(dim=(1, 2, 3))
        bp = (-torch.log2(rp["likelihoods"][k].double())).sum(dim=(1, 2, 3))
        assert ((bo - bp).abs() / bo).max().item() < 1e-4
    assert by.shape == (2,) and bz.shape == (2,)
    tot = sum((-torch.log2(v.double())).sum(dim=(1, 2, 3)) for v in rp["likelihoods"].values())
    assert ((by + bz - tot).abs() / tot).max().item() < 1e-6


def test_patch_rebinds_flex_backwarp(models):
    import copy

    import b200vc
    orc, prod = models
    foreign = b200vc.patch(copy.deepcopy(orc), fuse=False)
    img = torch.rand(1, 3, 32, 48, device="cuda")
    flow = 2 * torch.randn(1, 2, 32, 48, device="cuda")
    assert torch.equal(foreign.backwarp(img, flow), prod.backwarp(img, flow))
    assert torch.equal(foreign.backwarp(img, flow), orc.backwarp(img, flow))


def test_flex_gop16_matches_the_reference_loop(models):
    """BASELINE config 3: GOP-16, per-level (n, l) from the reference's ``qualities`` table
    (Flex-Rate.../test/testing.py:71-89, 186-200).  GopCoder batches the frames of a hierarchy level; the oracle
    model is driven frame by frame in the reference's coding order."""
    from b200vc import gop, synthetic
    orc, prod = models
    sch = gop.FLEX_GOP16
    frames = synthetic.make_sequence(17, 128, 192, seed=12, device="cuda")
    _, quality = gop.FLEX_QUALITIES[3]
    bits, sse, dec = gop.GopCoder(prod, sch, level_quality=quality).code(frames[None], (120, 190), want_decoded=True)
    assert bits.shape == (1, 17) and (bits[0, 1:16] > 0).all() and bits[0, 0] == 0 and bits[0, 16] == 0
    # The oracle is driven frame by frame in the reference's coding order, each frame from the references GopCoder
    # itself decoded (so that one rounding tie cannot propagate down the hierarchy and hide every later comparison).
    coding_order = [8, 4, 2, 1, 3, 6, 5, 7, 12, 10, 9, 11, 14, 13, 15]
    worst_bits, worst_mean, off_frames = 0.0, 0.0, []
    with torch.no_grad():
        for order in coding_order:
            n, l = quality[sch.levels[order]]
            ra, rb = sch.refs[order]
            out = orc(dec[:, ra], frames[order:order + 1], dec[:, rb], n=[n], l=l, train=False)
            worst_bits = max(worst_bits, abs(out["size"].item() - bits[0, order].item()) / out["size"].item())
            d = (out["x_hat"] - dec[0, order]).abs()
            print(f"  frame {order:2d} level {sch.levels[order]}: bits oracle {out['size'].item():.1f} kernels {bits[0, order].item():.1f}"
                  f" max|dx| {d.max().item():.3e} mean|dx| {d.mean().item():.3e} |x_hat| max {out['x_hat'].abs().max().item():.2f}")
            worst_mean = max(worst_mean, d.mean().item())
            if (d > 1e-3).float().mean().item() > 0.02:
                off_frames.append(order)
    print(f"flex GOP-16: worst per-frame bits rel err {worst_bits:.2e}; frames whose x_hat differs: {off_frames}")
    # GopCoder codes the frames of a level as one batch, the oracle one frame at a time: cuDNN picks other algorithms,
    # the latents move by ~1e-6 relative, and with random weights a hyper-latent symbol that flips at a rounding
    # boundary moves its whole frame by ~1e-2 (bits by ~1e-4).  What this test pins is the schedule: a wrong reference
    # pair or a wrong (n, l) changes the bits by percents and the frame by tenths.
    assert worst_bits < 1e-3
    assert worst_mean < 5e-2
    assert len(off_frames) <= 4
