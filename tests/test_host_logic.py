"""CPU: host-side logic of the product package -- interface mirror, checkpoint layout, GOP schedule and
sharding, error behaviour (no CPU fallback)."""
import pytest
import torch

import b200vc
from b200vc import gop, modules, ops, synthetic
from oracle import cai
from oracle import lhbdc as o_lhbdc


def test_checkpoint_layout_matches_the_reference_tree():
    prod, orc = b200vc.Model(), o_lhbdc.Model()
    a, b = prod.state_dict(), orc.state_dict()
    assert list(a) and set(a) == set(b)
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    prod.load_state_dict(b, strict=True)
    # the CompressAI key names the reference's checkpoints carry
    for k in ("mv_compressor.g_a.0.gdn.beta", "mv_compressor.g_a.0.gdn.gamma_reparam.lower_bound.bound",
              "residual_compressor.g_s.1.igdn.gamma", "residual_compressor.entropy_bottleneck._matrix0",
              "residual_compressor.entropy_bottleneck.quantiles", "mv_compressor.gaussian_conditional.scale_table",
              "mv_compressor.gaussian_conditional.lower_bound_scale.bound", "FlowNet.netBasic.5.netBasic.8.weight",
              "masknet.conv4.bias"):
        assert k in a, k


def test_module_mirror_has_the_compressai_surface():
    for name in ("GDN", "EntropyBottleneck", "GaussianConditional", "ResidualBlock", "ResidualBlockUpsample",
                 "ResidualBlockWithStride", "MeanScaleHyperprior", "conv3x3", "subpel_conv3x3"):
        assert hasattr(modules, name) and hasattr(cai, name)
    g, o = modules.GDN(16, inverse=True), cai.GDN(16, inverse=True)
    assert torch.equal(g.beta, o.beta) and torch.equal(g.gamma, o.gamma) and g.inverse
    e, oe = modules.EntropyBottleneck(8), cai.EntropyBottleneck(8)
    assert torch.equal(e._matrix2, oe._matrix2) and torch.equal(e.quantiles, oe.quantiles)
    assert torch.equal(e.target, oe.target)
    m = b200vc.Model()
    m.mv_compressor.update(force=True)
    assert torch.equal(m.mv_compressor.gaussian_conditional.scale_table, cai.get_scale_table())


def test_no_cpu_fallback():
    x = torch.rand(1, 3, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.backwarp(x, torch.zeros(1, 2, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        modules.GDN(32).eval()(torch.rand(1, 32, 4, 4))
    with pytest.raises(NotImplementedError):
        modules.EntropyBottleneck(4).train()(torch.rand(1, 4, 2, 2))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        b200vc.Model().mv_compressor.compress(torch.rand(1, 4, 64, 64))
    with pytest.raises(ValueError):
        ops.backwarp(x, x, variant="nope")


def test_gop_schedules_match_the_reference_tables():
    s = gop.LHBDC_GOP8
    assert s.by_level() == [[4], [2, 6], [1, 3, 5, 7]]
    # every frame's references are decoded before it (anchors 0 and gop are given)
    for sch in (gop.LHBDC_GOP8, gop.FLEX_GOP16):
        done = {0, sch.gop}
        for f in sch.order:
            a, b = sch.refs[f]
            assert a in done and b in done and a < f < b and (f - a) == (b - f)
            done.add(f)
        assert done == set(range(sch.gop + 1))
    assert gop.FLEX_GOP16.by_level()[0] == [8] and len(gop.FLEX_GOP16.by_level()) == 4
    assert gop.num_gops(97, 8) == 12 and gop.num_gops(600, 8) == 74 and gop.num_gops(8, 8) == 0


@pytest.mark.parametrize("units,world", [(74, 8), (12, 1), (38, 8), (3, 8), (0, 4)])
def test_shard_units_is_a_contiguous_partition(units, world):
    seen = []
    for r in range(world):
        rng = gop.shard_units(units, world, r)
        seen += list(rng)
        assert len(rng) in (units // world, units // world + 1)
    assert seen == list(range(units))
    with pytest.raises(ValueError):
        gop.shard_units(4, 2, 2)


def test_calibration_is_deterministic_and_layout_agnostic():
    torch.manual_seed(0)
    a = synthetic.calibrate_(o_lhbdc.Model(), 0).state_dict()
    torch.manual_seed(0)
    b = synthetic.calibrate_(o_lhbdc.Model(), 0).state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)
    p = b200vc.Model()
    p.load_state_dict(o_lhbdc.Model().state_dict())
    synthetic.calibrate_(p, 3)  # same recipe applies to the product tree


def test_psnr_from_sse():
    sse = torch.tensor(3.0 * 10 * 10 * 4.0, dtype=torch.float64)  # every value off by 2
    import math
    assert abs(gop.psnr_from_sse(sse, 300).item() - 10 * math.log10(255.0 ** 2 / 4)) < 1e-9


def test_flexrate_mirror_checkpoint_layout_and_gains():
    from b200vc import flexrate
    from oracle import flexrate as o_flex
    torch.manual_seed(0)
    orc = o_flex.BidirFlowRef(n=4, N=128)
    prod = flexrate.BidirFlowRef(n=4, N=128)
    a, b = prod.state_dict(), orc.state_dict()
    assert set(a) == set(b) and all(a[k].shape == b[k].shape for k in a)
    prod.load_state_dict(b, strict=True)
    for k in ("flow_compressor.gain_unit.gain_matrix", "residual_compressor.hyper_inv_gain_unit.gain_matrix",
              "flow_predictor.down_path.4.block.2.weight", "Mask.up_path.2.up.1.bias", "flow_compressor.g_s.7.0.weight"):
        assert k in a, k
    assert a["flow_compressor.g_s.7.0.weight"].shape[0] == 4 * 4 and a["flow_compressor.g_a.0.conv1.weight"].shape[1] == 19
    synthetic.calibrate_flex_(orc, 0)
    prod.load_state_dict(orc.state_dict())
    gp, go = prod.flow_compressor.gain_unit, orc.flow_compressor.gain_unit
    for n, l in (([1], 1.0), ([0], 0.33), ([2], 0.66)):
        assert torch.equal(gp.gain(n, l), go.gain(n, l)) and gp.gain(n, l).shape == (1, 128)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        prod.flow_compressor.compress(torch.rand(1, 19, 64, 64), [0], 1.0)


def _reference_literals(path, names):
    """Top-level literal assignments of a reference script, evaluated with ast.literal_eval (no code is executed)."""
    import ast
    import os
    if not os.path.exists(path):
        pytest.skip("/root/reference is not mounted here")
    out = {}
    for node in ast.parse(open(path).read()).body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name) \
                and node.targets[0].id in names:
            out[node.targets[0].id] = ast.literal_eval(node.value)
    return out


def test_schedules_and_qualities_equal_the_reference_scripts():
    """gop.py's tables against the literals in the reference's own drivers (LHBDC/test/testing.py:70-74,
    Flex-Rate.../test/testing.py:71-89)."""
    ref = _reference_literals("/root/reference/Flex-Rate-Hier-Bidir-Video-Compression/test/testing.py",
                              {"coding_order", "decoding_info", "hier_levels", "qualities"})
    sch = gop.FLEX_GOP16
    assert {k: tuple(v) for k, v in ref["decoding_info"].items()} == sch.refs
    assert ref["hier_levels"] == sch.levels
    assert sorted(ref["coding_order"][2:]) == sorted(sch.refs)
    assert tuple((q, {k: tuple(v) for k, v in d.items()}) for q, d in ref["qualities"]) == gop.FLEX_QUALITIES
    ref8 = _reference_literals("/root/reference/LHBDC/test/testing.py", {"coding_order", "decoding_info", "hier_levels"})
    if {"decoding_info", "hier_levels"} <= set(ref8):
        assert {k: tuple(v) for k, v in ref8["decoding_info"].items()} == gop.LHBDC_GOP8.refs
        assert ref8["hier_levels"] == gop.LHBDC_GOP8.levels


def test_ojsp_ratio_table_equals_the_reference():
    import ast
    import os
    import re
    path = "/root/reference/OJSP2025/video_model.py"
    if not os.path.exists(path):
        pytest.skip("/root/reference is not mounted here")
    from b200vc import ojsp
    m = re.search(r"downsampling_ratios\s*=\s*(\[[^\]]*\])", open(path).read())
    assert m and tuple(float(v) for v in ast.literal_eval(m.group(1))) == ojsp.DOWNSAMPLING_RATIOS


def test_icip_constants_equal_the_reference():
    import os
    import re
    base = "/root/reference"
    if not os.path.exists(base):
        pytest.skip("/root/reference is not mounted here")
    from b200vc import icip, modules
    src = open(f"{base}/ICIP2024/src/opt_helpers.py").read()
    m = re.search(r"for down_ratio in \[([^\]]*)\]", src)
    assert m and tuple(int(v) for v in m.group(1).split(",")) == icip.DOWN_RATIOS
    elic = open(f"{base}/ICIP2023/src/model/elic.py").read()
    consts = {k: float(re.search(rf"^{k}\s*=\s*([0-9.]+)", elic, re.M).group(1))
              for k in ("SCALES_MIN", "SCALES_MAX", "SCALES_LEVELS")}
    assert (consts["SCALES_MIN"], consts["SCALES_MAX"], int(consts["SCALES_LEVELS"])) == \
        (modules.SCALES_MIN, modules.SCALES_MAX, modules.SCALES_LEVELS)
    cb = open(f"{base}/ICIP2024/src/model/compression_bottlenecks.py").read()
    groups = re.search(r"uneven_groups = \[(.*?)\]\n", cb, re.S).group(1)
    bounds = [int(v) for v in re.findall(r"y\[:, (?::)?(\d+)", groups)]
    # y[:, :6], y[:, 6:12], y[:, 12:24], y[:, 24:48], y[:, 48:]  ->  group sizes 6, 6, 12, 24, rest
    starts = sorted(set(bounds))
    sizes = tuple(b - a for a, b in zip([0] + starts[:-1], starts))
    assert sizes + (None,) == icip.ELIC_GROUPS


def test_elic_group_layout_is_validated_before_any_kernel_runs():
    from b200vc import icip
    for M in (40, 48):          # 6 + 6 + 12 + 24 = 48 leaves nothing (or less) for the last group
        with pytest.raises(ValueError):
            icip.elic_context_likelihoods(torch.zeros(1, M, 2, 2), None, None, None, None, None)
    with pytest.raises(ValueError):
        icip.elic_context_likelihoods(torch.zeros(1, 64, 2, 2), None, None, None, None, None, group_sizes=(6, 6))


def test_flowguided_mirror_has_the_reference_checkpoint_layout():
    """b200vc.flowguided.FlowGuidedB: same state-dict keys and shapes as the oracle restatement, which
    oracle/make_golden_flowguided.py pins key for key to the reference's ICIP2024/src/model/m.py."""
    import torch
    from b200vc import flowguided
    from oracle import flowguided as o_fg
    torch.manual_seed(0)
    ref, mir = o_fg.FlowGuidedB().state_dict(), flowguided.FlowGuidedB().state_dict()
    assert set(ref) == set(mir)
    assert all(ref[k].shape == mir[k].shape for k in ref)
    for n in (2, 17, 40, 300, 600):
        assert flowguided.get_order_typ_list(16, n) == o_fg.get_order_typ_list(16, n)
    for order, buf in ((8, [0, 16]), (4, [0, 16, 8]), (12, [0, 16, 8, 4]), (1, [0]), (598, [599, 595, 593, 597, 594, 596])):
        assert flowguided.select_references(order, buf) == o_fg.select_references(order, buf)
    assert flowguided.get_scales(4, 0, 8) == o_fg.get_scales(4, 0, 8) == (0.5, 0.5)
    assert flowguided.get_scales(3, 3, 3) == (0, 0)


def test_mbt2018_mean_mirror_has_compressai_topology_and_keys():
    """b200vc.modules.mbt2018_mean(q): CompressAI's cfgs["mbt2018-mean"] widths, the oracle's (CompressAI-shaped)
    state-dict keys, the zoo function's argument checks (LHBDC/test/testing.py:209)."""
    import pytest
    from b200vc import modules as M
    from oracle import cai
    assert M.MBT2018_MEAN_CFG[4] == (128, 192) and M.MBT2018_MEAN_CFG[5] == (192, 320) and len(M.MBT2018_MEAN_CFG) == 8
    for q in (1, 7):
        N, Mm = M.MBT2018_MEAN_CFG[q]
        prod, orc = M.mbt2018_mean(q), cai.MeanScaleHyperprior(N, Mm)
        assert sorted(prod.state_dict()) == sorted(orc.state_dict())
        assert all(prod.state_dict()[k].shape == v.shape for k, v in orc.state_dict().items())
        prod.load_state_dict(orc.state_dict(), strict=True)
    with pytest.raises(ValueError):
        M.mbt2018_mean(9)
    with pytest.raises(ValueError):
        M.mbt2018_mean(3, metric="ms-ssim")
    with pytest.raises(RuntimeError):
        M.mbt2018_mean(3, pretrained=True)
