"""GPU parity: the ICIP2024 ``FlowGuidedB`` mirror (K-WARP feature warps, K-DCN fusion, gain-folded K-EB, the
checkerboard context loop on K-CHK + K-GC, fused search) vs the oracle restatement (pinned bit for bit to the
reference's own ICIP2024/src/model/*.py by oracle/make_golden_flowguided.py), same weights, same device, TF32 off."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def models(strict_fp32):
    from b200vc import flowguided, synthetic
    from oracle import flowguided as o_fg
    torch.manual_seed(0)
    orc = o_fg.FlowGuidedB().eval()
    synthetic.calibrate_flowguided_(orc, 0)
    prod = flowguided.FlowGuidedB().eval()
    prod.load_state_dict(orc.state_dict())           # strict: the mirror has exactly the reference's key layout
    return orc.cuda(), prod.cuda()


def _bits64(res):
    return sum((-torch.log2(v.double())).sum() for v in res["likelihoods"].values()).item()


@pytest.mark.parametrize("which", ["offset_compressor", "residual_compressor"])
@pytest.mark.parametrize("s", [0, 2, 3.25])
def test_elic_bottlenecks_match_oracle(models, which, s):
    """Offset_ELIC / Res_ELIC.forward (compression_bottlenecks.py:213-289, :454-530) on identical inputs."""
    orc, prod = models
    g = torch.Generator().manual_seed(5)
    r = lambda *sh: torch.randn(*sh, generator=g).cuda()
    H8, W8 = 16, 24
    mult, dmult = (5, 4) if which == "offset_compressor" else (1, 1)
    pyr = lambda m: (r(1, 64 * m, 4 * H8, 4 * W8), r(1, 96 * m, 2 * H8, 2 * W8), r(1, 128 * m, H8, W8))
    f, fd, temp = pyr(mult), pyr(dmult), r(1, 128, H8 // 2, W8 // 2)
    o, p = getattr(orc, which), getattr(prod, which)
    with torch.no_grad():
        want = o(*f, *fd, temp, s)
        got = p(*f, *fd, temp, s)
        got_b = p.forward_bits(*f, *fd, temp, s)
    assert set(want) == set(got)
    for k in want:
        if k == "likelihoods":
            assert list(want[k]) == list(got[k])
            for name in want[k]:
                # K-GC / K-EB are bit-exact on identical (y, scales, means) (tests/test_gpu_entropy.py, test_gpu_checker.py);
                # here the Gaussian parameters come out of cuDNN convolutions that read differently laid-out buffers on
                # the two sides (channel slices of one quantised tensor vs fresh tensors), so they may differ in the
                # last bit, which the CDF difference amplifies at large scales (measured 1.6e-5 on one element)
                rel = ((got[k][name] - want[k][name]).abs() / want[k][name]).max().item()
                assert rel < 1e-4, (name, rel)
        else:
            # decoded tensors: cuDNN convolutions over differently laid-out inputs on the two sides; seen 1.2e-8 .. 1.15e-5
            # relative depending on the box (a flipped symbol would show as >= 1e-2), bar 5e-5
            d = (got[k] - want[k]).abs().max().item()
            assert d <= 5e-5 * max(1.0, want[k].abs().max().item()), (k, d)
    b_o = _bits64(want)
    rel = abs(got_b["bits"].sum().item() - b_o) / b_o
    print(f"{which} s={s}: bits oracle {b_o:.2f}, bits-only pass rel {rel:.2e}")
    assert rel < 1e-6
    assert "likelihoods" not in got_b


@pytest.mark.parametrize("args", [(0.5, 0.5, 2, 2), (0.25, 0.75, 0, 1), (0.5, 0.5, 3.5, 16)])
@pytest.mark.parametrize("shape", [(128, 192), (256, 448)])
def test_forward_matches_oracle(models, args, shape):
    from b200vc import synthetic
    orc, prod = models
    s1, s2, s, ratio = args
    fr = synthetic.make_sequence(3, *shape, seed=21, device="cuda")
    with torch.no_grad():
        want = orc(fr[0:1], fr[2:3], s1, s2, fr[1:2], s, ratio)
        got = prod(fr[0:1], fr[2:3], s1, s2, fr[1:2], s, ratio)
    rel = abs(got["size"].item() - want["size"].item()) / want["size"].item()
    d = (got["x_hat"] - want["x_hat"]).abs()
    print(f"FlowGuidedB.forward {shape} {args}: size oracle {want['size'].item():.2f} kernels {got['size'].item():.2f} "
          f"rel {rel:.2e}; x_hat max|d| {d.max().item():.2e}")
    assert rel < 1e-4
    assert (d < 1e-3).float().mean().item() > 0.999
    assert abs(got["rate"].item() - want["rate"].item()) / want["rate"].item() < 1e-4


def test_forward_matches_reference_golden(models, golden_dir):
    """Golden = the reference's own FlowGuidedB.forward (CPU run through the compressai stand-in)."""
    _, prod = models
    gold = np.load(os.path.join(golden_dir, "flowguided_reference.npz"))
    fr = torch.from_numpy(gold["frames_u8"]).cuda().float() / 255.0
    for tag in ("a", "b", "c"):
        s1, s2, s, ratio = gold[f"fwd_{tag}_args"]
        s = int(s) if float(s).is_integer() else float(s)
        with torch.no_grad():
            got = prod(fr[0:1], fr[2:3], float(s1), float(s2), fr[1:2], s, int(ratio))
        rel = abs(got["size"].item() - float(gold[f"fwd_{tag}_size"])) / float(gold[f"fwd_{tag}_size"])
        close = ((got["x_hat"].cpu() - torch.from_numpy(gold[f"fwd_{tag}_x_hat"])).abs() < 2e-3).float().mean().item()
        print(f"vs reference golden ({tag}): size rel {rel:.2e}; x_hat within 2e-3: {close:.5f}")
        assert rel < 1e-4 and close > 0.999


def test_sequence_coder_matches_the_reference_loop(models):
    """test.py:37-93 at one rate level on a 17-frame GOP-16: coding order, nearest-two references from the decoded
    buffer, temporal scales, per-frame down-ratio search, bits and uint8 PSNR."""
    from b200vc import flowguided, synthetic
    from oracle import flowguided as o_fg
    orc, prod = models
    frames = synthetic.make_sequence(17, 128, 192, seed=3, device="cuda")
    bits, sse, ratios, dec = flowguided.SequenceCoder(prod, gop=16).code(frames, (120, 190), 2, want_decoded=True)
    bits_o, sse_o, ratios_o = o_fg.code_sequence(orc, frames, 2, (120, 190))
    assert ratios == ratios_o, (ratios, ratios_o)
    b, so = bits.cpu().tolist(), sse.cpu().tolist()
    worst = max(abs(b[t] - bits_o[t]) / bits_o[t] for t in range(1, 16))
    tot = abs(sum(b) - sum(bits_o)) / sum(bits_o)
    print(f"GOP-16 sequence: worst per-frame bits rel {worst:.2e}, total rel {tot:.2e}, ratios {ratios[1:16]}")
    assert b[0] == 0 and b[16] == 0 and so[0] == 0 and so[16] == 0
    assert tot < 1e-4 and worst < 1e-3
    assert dec.shape == frames.shape
