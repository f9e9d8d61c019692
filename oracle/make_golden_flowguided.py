"""TEST INFRASTRUCTURE (oracle).  Pins ``oracle/flowguided.py`` (and through it ``oracle/icip.py``'s ELIC context loop,
VERDICT r1 item 7) to the REFERENCE'S OWN CODE, read from ``/root/reference`` and never copied:

* ``ICIP2024/src/model/{m,helpers,compression_bottlenecks,elic,layers}.py`` are imported verbatim as the package
  ``src.model`` through the ``compressai`` stand-in (``oracle/shim.py``); ``FlowGuidedB()`` built under
  ``torch.manual_seed(0)`` must have exactly the oracle's state dict (same keys, same values: the restatement creates
  its parameters in the reference's order);
* ``Offset_ELIC.forward`` (:213-289), ``Res_ELIC.forward`` (:454-530), ``FlowGuidedB.forward`` (m.py:181-260) are run
  on the reference and on the oracle with identical weights and must agree bit for bit (CPU, one thread);
* ``opt_helpers.py`` ``prediction_flowonly`` / ``get_best_down_ratio_prediction`` and ``utils.py``
  ``get_order_typ_list`` / ``select_references`` / ``update_buffer`` / ``get_scales`` are extracted with ``ast`` and
  compared likewise.

Writes ``tests/golden/flowguided_reference.npz`` (inputs as uint8 frames; outputs; schedules).

    python -m oracle.make_golden_flowguided      # needs /root/reference; the GPU box never runs this
"""
import ast
import importlib
import os
import sys
import textwrap

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

REF = "/root/reference/ICIP2024"
OUT = os.path.join(ROOT, "tests", "golden", "flowguided_reference.npz")


def extract(path, names, ns):
    src = open(path).read()
    for node in ast.parse(src).body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in names:
            exec(compile(textwrap.dedent(ast.get_source_segment(src, node)), f"{path}:{node.name}", "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def same(a, b, what):
    if isinstance(a, dict):
        assert set(a) == set(b), (what, set(a) ^ set(b))
        for k in a:
            same(a[k], b[k], f"{what}[{k}]")
    else:
        assert torch.equal(torch.as_tensor(a), torch.as_tensor(b)), f"{what}: oracle restatement != reference"


def main():
    torch.set_num_threads(1)
    from b200vc import synthetic
    from oracle import flowguided as o_fg
    from oracle import shim
    shim.install()
    sys.path.insert(0, REF)
    ref_m = importlib.import_module("src.model.m")
    ns = extract(os.path.join(REF, "src/utils.py"),
                 ["MSE", "PSNR", "get_order_typ_list", "select_references", "update_buffer", "get_scales"],
                 {"torch": torch, "np": np, "F": F, "nn": nn})
    extract(os.path.join(REF, "src/opt_helpers.py"), ["prediction_flowonly", "get_best_down_ratio_prediction"], ns)

    torch.manual_seed(0)
    ref = ref_m.FlowGuidedB().eval()
    torch.manual_seed(0)
    orc = o_fg.FlowGuidedB().eval()
    sd_r, sd_o = ref.state_dict(), orc.state_dict()
    assert list(sd_r) == list(sd_o), "state-dict keys / order differ"
    for k in sd_r:
        assert torch.equal(sd_r[k], sd_o[k]), f"same-seed construction differs at {k}"
    synthetic.calibrate_flowguided_(ref, 0)
    synthetic.calibrate_flowguided_(orc, 0)
    same(dict(ref.state_dict()), dict(orc.state_dict()), "calibrated state dict")

    frames_u8 = (synthetic.make_sequence(3, 128, 192, seed=21) * 255).round().to(torch.uint8)
    fr = frames_u8.float() / 255.0
    x1, xc, x2 = fr[0:1], fr[1:2], fr[2:3]
    out = {"frames_u8": frames_u8.numpy()}
    with torch.no_grad():
        # ---- the two gain-modulated ELIC bottlenecks alone (the context loop of VERDICT item 7) --------------------
        g = torch.Generator().manual_seed(5)
        r = lambda *s: torch.randn(*s, generator=g)
        H8, W8 = 16, 24                                              # 1/8 resolution of a 128 x 192 frame
        pyr = lambda c1, c2, c3: (r(1, c1, 4 * H8, 4 * W8), r(1, c2, 2 * H8, 2 * W8), r(1, c3, H8, W8))
        f = pyr(64 * 5, 96 * 5, 128 * 5)
        fd = pyr(64 * 4, 96 * 4, 128 * 4)
        temp = r(1, 128, H8 // 2, W8 // 2)
        for s in (0, 2, 3.25):
            a = ref.offset_compressor(*f, *fd, temp, s)
            b = orc.offset_compressor(*f, *fd, temp, s)
            same(a, b, f"Offset_ELIC.forward(s={s})")
        out["offset_elic_bits_s2"] = np.float64(sum((-torch.log2(v.double())).sum() for v in
                                                    orc.offset_compressor(*f, *fd, temp, 2)["likelihoods"].values()))
        f = pyr(64, 96, 128)
        fd = pyr(64, 96, 128)
        for s in (1, 4, 0.5):
            same(ref.residual_compressor(*f, *fd, temp, s), orc.residual_compressor(*f, *fd, temp, s),
                 f"Res_ELIC.forward(s={s})")
        # ---- whole model ---------------------------------------------------------------------------------------------
        for tag, (s1, s2, s, ratio) in {"a": (0.5, 0.5, 2, 2), "b": (0.25, 0.75, 0, 1), "c": (0.5, 0.5, 3.5, 16)}.items():
            a = ref(x1, x2, s1, s2, xc, s, ratio)
            b = orc(x1, x2, s1, s2, xc, s, ratio)
            same(a, b, f"FlowGuidedB.forward({tag})")
            out[f"fwd_{tag}_x_hat"] = a["x_hat"].numpy()
            out[f"fwd_{tag}_size"] = np.float32(a["size"].item())
            out[f"fwd_{tag}_rate"] = np.float32(a["rate"].item())
            out[f"fwd_{tag}_args"] = np.array([s1, s2, s, ratio], dtype=np.float64)
        # ---- down-ratio search -------------------------------------------------------------------------------------
        ra, pa = ns["get_best_down_ratio_prediction"](ref, x1, x2, 0.5, 0.5, xc, 2, None)
        rb, pb = o_fg.get_best_down_ratio_prediction(orc, x1, x2, 0.5, 0.5, xc)
        assert ra == rb and torch.equal(pa, pb), (ra, rb, pa, pb)
        out["search_ratio"], out["search_psnr"] = np.int64(ra), np.float32(pa.item())
        same(ns["prediction_flowonly"](ref, xc, x1, x2, 0.5, 0.5, 4),
             __import__("oracle.icip", fromlist=["x"]).prediction_flowonly(orc, xc, x1, x2, 0.5, 0.5, 4), "prediction_flowonly")
    # ---- schedules --------------------------------------------------------------------------------------------------
    for n in (2, 17, 33, 40, 300, 600):
        o_r, t_r = ns["get_order_typ_list"](16, n)
        o_o, t_o = o_fg.get_order_typ_list(16, n)
        assert list(o_r) == list(o_o) and list(t_r) == list(t_o), n
        out[f"order_{n}"] = np.array(o_r, dtype=np.int64)
        out[f"types_{n}"] = np.array([t == "I" for t in t_r])
        # reference selection / scales while walking that order with the reference's own buffer functions
        buf, buf_o, picks = [], [], []
        for order in o_r:
            if t_r[order] != "I":
                _, _, o1, o2 = ns["select_references"](None, order, buf, buf_o)
                i1, i2 = o_fg.select_references(order, buf_o)
                assert (buf_o[i1], buf_o[i2]) == (o1, o2), (n, order)
                assert tuple(ns["get_scales"](order, o1, o2)) == tuple(o_fg.get_scales(order, o1, o2))
                picks.append((order, o1, o2))
            buf, buf_o = ns["update_buffer"](buf, buf_o, order, order)
        out[f"refs_{n}"] = np.array(picks, dtype=np.int64).reshape(-1, 3)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, f"{os.path.getsize(OUT) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
