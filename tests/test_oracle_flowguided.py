"""CPU: ``oracle/flowguided.py`` (ICIP2024 ``FlowGuidedB`` + its evaluation loop, incl. the ELIC checkerboard /
channel-group context loop of ``oracle/icip.py``) against ``tests/golden/flowguided_reference.npz`` -- outputs of the
REFERENCE'S OWN ``ICIP2024/src/model/*.py`` / ``opt_helpers.py`` / ``utils.py`` run through the compressai stand-in by
``oracle/make_golden_flowguided.py`` (which asserts bit equality on the generating machine; here oneDNN may pick other
convolution kernels, hence the small tolerances around quantisers)."""
import os

import numpy as np
import pytest
import torch

from oracle import flowguided as o_fg


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "flowguided_reference.npz"))


@pytest.fixture(scope="module")
def oracle_model():
    from b200vc import synthetic
    torch.manual_seed(0)
    m = o_fg.FlowGuidedB().eval()
    synthetic.calibrate_flowguided_(m, 0)
    return m


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_forward_matches_reference(gold, oracle_model, tag):
    fr = torch.from_numpy(gold["frames_u8"]).float() / 255.0
    s1, s2, s, ratio = gold[f"fwd_{tag}_args"]
    s = int(s) if float(s).is_integer() else float(s)
    with torch.no_grad():
        out = oracle_model(fr[0:1], fr[2:3], float(s1), float(s2), fr[1:2], s, int(ratio))
    want = torch.from_numpy(gold[f"fwd_{tag}_x_hat"])
    close = ((out["x_hat"] - want).abs() < 1e-3).float().mean().item()
    rel = abs(out["size"].item() - float(gold[f"fwd_{tag}_size"])) / float(gold[f"fwd_{tag}_size"])
    print(f"FlowGuidedB.forward({tag}): size rel {rel:.2e}, x_hat within 1e-3: {close:.5f}")
    assert close > 0.999 and rel < 1e-3
    assert abs(out["rate"].item() - float(gold[f"fwd_{tag}_rate"])) / float(gold[f"fwd_{tag}_rate"]) < 1e-3


def test_down_ratio_search_matches_reference(gold, oracle_model):
    fr = torch.from_numpy(gold["frames_u8"]).float() / 255.0
    with torch.no_grad():
        ratio, psnr = o_fg.get_best_down_ratio_prediction(oracle_model, fr[0:1], fr[2:3], 0.5, 0.5, fr[1:2])
    assert ratio == int(gold["search_ratio"]) and abs(float(psnr) - float(gold["search_psnr"])) < 1e-3


@pytest.mark.parametrize("n", [2, 17, 33, 40, 300, 600])
def test_schedule_and_reference_selection(gold, n):
    """Coding order incl. the reference's hard-coded 300 / 600-frame tails, I/B types, nearest-two reference picks
    and temporal scales for every B-frame (ICIP2024/src/utils.py:154-243)."""
    order, typ = o_fg.get_order_typ_list(16, n)
    assert order == gold[f"order_{n}"].tolist()
    assert [t == "I" for t in typ] == gold[f"types_{n}"].tolist()
    assert sorted(order) == list(range(n))
    buf, picks = [], []
    for o in order:
        if typ[o] != "I":
            i1, i2 = o_fg.select_references(o, buf)
            picks.append((o, buf[i1], buf[i2]))
            s1, s2 = o_fg.get_scales(o, buf[i1], buf[i2])
            if buf[i1] != buf[i2]:
                assert abs(s1 + s2 - 1.0) < 1e-12 and buf[i1] < buf[i2]
        buf = (buf + [o])[-32:] if len(buf) >= 32 else buf + [o]
    assert picks == [tuple(r) for r in gold[f"refs_{n}"].tolist()]


def test_state_dict_has_the_reference_layout(oracle_model):
    """Key layout of the reference checkpoint (incl. the members the joint-autoregressive base class creates and the
    ICIP subclasses never call)."""
    keys = set(oracle_model.state_dict())
    for k in ("feature_extractor.layer1.0.weight", "flow_estimator.up3.3.0.weight",
              "offset_compressor.g_a.1.beta", "offset_compressor.context_prediction.mask",
              "offset_compressor.context_prediction_models.4.mask", "offset_compressor.Gain",
              "offset_compressor.entropy_bottleneck._matrix0", "offset_diversity_l2.fusion.weight",
              "residual_compressor.InverseHyperGain", "residual_compressor.g_o1.4.bias",
              "reconstructor.layer1.4.0.weight", "residue_temporal_conditioner.g_a3.3.BottleneckBlock.4.bias"):
        assert k in keys, k
