"""CPU: the oracle restatement against the golden vectors produced by the reference's own code
(oracle/make_golden.py).  Bit-exact where only elementwise/gather arithmetic is involved; tolerant where
oneDNN convolution kernels (ISA dependent) sit in between."""
import os

import numpy as np
import pytest
import torch

from oracle import lhbdc as o_lhbdc
from oracle import warp as o_warp


@pytest.fixture(scope="module")
def warp_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "warp_reference.npz"))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("variant,fn", [("lhbdc", o_warp.backwarp_lhbdc), ("flex", o_warp.backwarp_flex),
                                        ("ac1", o_warp.warp_ac1)])
def test_warp_restatement_matches_reference(warp_golden, tag, variant, fn):
    img = torch.from_numpy(warp_golden[f"{tag}_img"])
    flow = torch.from_numpy(warp_golden[f"{tag}_flow"])
    want = torch.from_numpy(warp_golden[f"{tag}_{variant}"])
    got = fn(img, flow)
    # same torch build => identical; a different CPU ISA may pick another grid_sample vector path
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)


def test_warp_known_geometry():
    """SURVEY Appendix C.1 probes: LHBDC 1-px flow moves W/(W-1) px; Flex zero flow = 2x2 mean with zero fade;
    ICIP 1-px flow is an exact shift."""
    W = 12
    ramp = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W).repeat(1, 1, 4, 1)
    one = torch.zeros(1, 2, 4, W)
    one[:, 0] = 1.0
    out = o_warp.backwarp_lhbdc(ramp, one)
    assert abs(out[0, 0, 1, 5].item() - (5 + W / (W - 1.0))) < 1e-5
    out = o_warp.warp_ac1(ramp, one)
    assert torch.allclose(out[0, 0, :, :-1], ramp[0, 0, :, 1:], atol=1e-5)
    img = torch.rand(1, 1, 6, 6)
    out = o_warp.backwarp_flex(img, torch.zeros(1, 2, 6, 6))
    box = (img[0, 0, :-1, :-1] + img[0, 0, 1:, :-1] + img[0, 0, :-1, 1:] + img[0, 0, 1:, 1:]) / 4
    assert torch.allclose(out[0, 0, 1:, 1:], box, atol=1e-6)
    assert torch.allclose(out[0, 0, 0, 0], img[0, 0, 0, 0] / 4, atol=1e-6)


@pytest.fixture(scope="module")
def model_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "lhbdc_model_reference.npz"))


@pytest.fixture(scope="module")
def oracle_model():
    from b200vc import synthetic
    torch.manual_seed(0)
    m = o_lhbdc.Model().eval()
    synthetic.calibrate_(m, 0)
    m.mv_compressor.update(force=True)
    m.residual_compressor.update(force=True)
    return m


@pytest.mark.parametrize("tag,key", [("synthetic", "synthetic_u8"), ("frames", "frames_crop_u8")])
def test_model_restatement_matches_reference(model_golden, oracle_model, tag, key):
    tri = torch.from_numpy(model_golden[key])
    xb, xc, xa = (tri[i:i + 1].float() / 255.0 for i in range(3))
    with torch.no_grad():
        x_hat, rate, size, parts = oracle_model(xb, xc, xa, train=False, return_parts=True)
    want = torch.from_numpy(model_golden[f"{tag}_x_hat"])
    # warped frame: no quantiser upstream of it except the mv latents -> tight
    fw = torch.from_numpy(model_golden[f"{tag}_fw"])
    assert (parts["fw"] - fw).abs().max().item() < 5e-3
    frac_close = ((x_hat - want).abs() < 1e-3).float().mean().item()
    assert frac_close > 0.999, frac_close
    assert abs(size - float(model_golden[f"{tag}_size"])) / float(model_golden[f"{tag}_size"]) < 1e-3
    assert abs(rate.item() - float(model_golden[f"{tag}_rate"])) / float(model_golden[f"{tag}_rate"]) < 1e-3
    # fp32 tree sum vs order-independent fp64 total
    assert abs(size - parts["size64"]) / parts["size64"] < 1e-5


def test_synthetic_sequence_is_deterministic(model_golden):
    from b200vc import synthetic
    seq = (synthetic.make_sequence(9, 192, 192, seed=1234) * 255).round().to(torch.uint8)
    want = torch.from_numpy(model_golden["synthetic_u8"])
    diff = (seq[[0, 4, 8]].int() - want.int()).abs()
    assert diff.max().item() <= 1 and (diff > 0).float().mean().item() < 1e-3


@pytest.mark.parametrize("tag,n,l", [("n1_l1", [1], 1.0), ("n0_l066", [0], 0.66)])
def test_flexrate_restatement_matches_reference(golden_dir, tag, n, l):
    """Golden = the reference's own b_model/b_model.py BidirFlowRef (imported through the compressai stand-in)."""
    from b200vc import synthetic
    from oracle import flexrate as o_flex
    gold = np.load(os.path.join(golden_dir, "flexrate_model_reference.npz"))
    torch.manual_seed(0)
    m = o_flex.BidirFlowRef(n=4, N=128).eval()
    synthetic.calibrate_flex_(m, 0)
    tri = torch.from_numpy(gold["frames_u8"])
    xb, xc, xa = (tri[i:i + 1].float() / 255.0 for i in range(3))
    with torch.no_grad():
        out = m(xb, xc, xa, n=n, l=l, train=False)
    want = torch.from_numpy(gold[f"{tag}_x_hat"])
    assert ((out["x_hat"] - want).abs() < 1e-3).float().mean().item() > 0.999
    assert abs(out["size"].item() - float(gold[f"{tag}_size"][0])) / float(gold[f"{tag}_size"][0]) < 1e-3
    assert out["rate"].shape == (1,) and abs(out["rate"].item() - out["size"].item() / (128 * 192)) < 1e-3
