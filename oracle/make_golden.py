"""TEST INFRASTRUCTURE (oracle).  Generates ``tests/golden/*.npz`` by running the REFERENCE'S OWN CODE
(read from ``/root/reference``, never copied) in the build container:

* the four warp functions are pulled out of the reference source files with ``ast`` and executed verbatim
  (``LHBDC/model/flow.py`` is imported whole; the others need absent dependencies at module import time);
* ``LHBDC/model/m.py`` ``Model`` is imported verbatim through the ``compressai`` stand-in of ``oracle/shim.py``
  and run on seeded inputs and on a crop of the bundled ``LHBDC/frames`` triple.

Run:  python -m oracle.make_golden        (needs /root/reference; the GPU box never runs this)
"""
import ast
import os
import sys
import textwrap

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "video-compression_b200"))

from oracle import lhbdc as o_lhbdc  # noqa: E402
from oracle import shim  # noqa: E402
from oracle import warp as o_warp  # noqa: E402


def extract_method(path, cls, name):
    """Compile one method of one class of a reference file, verbatim, into a plain function."""
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == name:
                    code = textwrap.dedent(ast.get_source_segment(src, item))
                    ns = {"torch": torch, "F": F, "np": np, "device": torch.device("cpu"), "nn": torch.nn}
                    exec(compile(code, f"{path}:{cls}.{name}", "exec"), ns)
                    return ns[name]
    raise KeyError(f"{cls}.{name} not found in {path}")


def warp_inputs(seed, N, C, H, W, amp):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(N, C, H, W, generator=g)
    flow = amp * torch.randn(N, 2, H, W, generator=g)
    # 2 % of the vectors leave the frame; a few exact-integer and half-integer displacements
    far = torch.rand(N, 1, H, W, generator=g) < 0.02
    flow = torch.where(far, flow * 40.0, flow)
    flow[:, :, 0, :] = 0.0
    flow[:, :, 1, :] = 1.0
    flow[:, :, 2, :] = -0.5
    return img, flow


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)

    # ---- W2: flow.backwarp, imported whole --------------------------------------------------------
    sys.path.insert(0, os.path.join(REF, "LHBDC", "model"))
    import flow as ref_flow  # LHBDC/model/flow.py
    sys.path.pop(0)
    ref_flow.device = torch.device("cpu")
    model_backwarp = extract_method(os.path.join(REF, "LHBDC/model/m.py"), "Model", "backwarp")
    flex_backwarp = extract_method(
        os.path.join(REF, "Flex-Rate-Hier-Bidir-Video-Compression/b_model/b_model.py"), "BidirFlowRef", "backwarp")
    icip_warp = extract_method(os.path.join(REF, "ICIP2024/src/model/m.py"), "FlowGuidedB", "warp")
    ojsp_warp = extract_method(os.path.join(REF, "OJSP2025/video_model.py"), "DMC", "warp")
    torch.Tensor.cuda = lambda self, *a, **k: self  # BidirFlowRef.backwarp calls .cuda() on its grid

    cases = {}
    for tag, (N, C, H, W, amp) in {"a": (2, 3, 20, 28, 1.5), "b": (1, 3, 34, 60, 4.0), "c": (1, 5, 17, 23, 2.0)}.items():
        img, flow = warp_inputs(100 + ord(tag), N, C, H, W, amp)
        ref_flow.backwarp_tenGrid.clear()
        w2 = ref_flow.backwarp(img, flow)
        w1 = model_backwarp(None, img, flow)
        w3 = flex_backwarp(None, img, flow)
        w4 = icip_warp(None, img, flow)
        w4b = ojsp_warp(None, img, flow)
        assert torch.equal(w1, w2) and torch.equal(w4, w4b)
        # the restatement must reproduce the reference's own functions bit for bit on the same machine
        assert torch.equal(o_warp.backwarp_lhbdc(img, flow), w1)
        assert torch.equal(o_warp.backwarp_flex(img, flow), w3)
        assert torch.equal(o_warp.warp_ac1(img, flow), w4)
        cases.update({f"{tag}_img": img.numpy(), f"{tag}_flow": flow.numpy(), f"{tag}_lhbdc": w1.numpy(),
                      f"{tag}_flex": w3.numpy(), f"{tag}_ac1": w4.numpy()})
    np.savez_compressed(os.path.join(OUT, "warp_reference.npz"), **cases)
    print("warp_reference.npz", len(cases), "arrays")

    # ---- Model.forward through the compressai stand-in ---------------------------------------------
    from b200vc import synthetic
    ref_m = shim.import_reference(os.path.join(REF, "LHBDC"), "model.m")
    ref_m.device = torch.device("cpu")
    sys.modules["model.flow"].device = torch.device("cpu")

    def build(cls):
        torch.manual_seed(0)
        m = cls().eval()
        synthetic.calibrate_(m, 0)
        m.mv_compressor.update(force=True)
        m.residual_compressor.update(force=True)
        return m

    rm, om = build(ref_m.Model), build(o_lhbdc.Model)
    sd_r, sd_o = rm.state_dict(), om.state_dict()
    assert set(sd_r) == set(sd_o)
    assert all(torch.equal(sd_r[k], sd_o[k]) for k in sd_r), "construction order drifted from the reference"

    seq = (synthetic.make_sequence(9, 192, 192, seed=1234) * 255).round().to(torch.uint8)
    from PIL import Image
    crop = []
    for nm in ("ref_1", "current", "ref_2"):
        im = np.asarray(Image.open(os.path.join(REF, "LHBDC/frames", nm + ".png")).convert("RGB"))
        crop.append(torch.from_numpy(im[400:592, 800:992].copy()).permute(2, 0, 1))
    crop = torch.stack(crop)  # [3,3,192,192] uint8: before, current, after

    out = {"synthetic_u8": seq[[0, 4, 8]].numpy(), "frames_crop_u8": crop.numpy()}
    with torch.no_grad():
        for tag, tri in (("synthetic", seq[[0, 4, 8]]), ("frames", crop)):
            xb, xc, xa = (tri[i:i + 1].float() / 255.0 for i in range(3))
            x_hat, rate, size = rm(xb, xc, xa, train=False)
            x2, r2, s2, parts = om(xb, xc, xa, train=False, return_parts=True)
            assert torch.equal(x_hat, x2) and rate.item() == r2.item() and size == s2
            out.update({f"{tag}_x_hat": x_hat.numpy(), f"{tag}_rate": np.float32(rate.item()),
                        f"{tag}_size": np.float64(size), f"{tag}_size64": np.float64(parts["size64"]),
                        f"{tag}_fw": parts["fw"].numpy(), f"{tag}_residual": parts["residual"].numpy()})
            print(tag, "rate", rate.item(), "size", size)
    np.savez_compressed(os.path.join(OUT, "lhbdc_model_reference.npz"), **out)
    print("lhbdc_model_reference.npz written")

    # ---- Flex-Rate BidirFlowRef through the compressai stand-in --------------------------------------
    from oracle import flexrate as o_flex
    ref_b = shim.import_reference(os.path.join(REF, "Flex-Rate-Hier-Bidir-Video-Compression"), "b_model.b_model")

    def build_flex(cls):
        torch.manual_seed(0)
        m = cls(n=4, N=128).eval()
        synthetic.calibrate_flex_(m, 0)
        return m

    rf, of = build_flex(ref_b.BidirFlowRef), build_flex(o_flex.BidirFlowRef)
    sd_r, sd_o = rf.state_dict(), of.state_dict()
    assert set(sd_r) == set(sd_o) and all(torch.equal(sd_r[k], sd_o[k]) for k in sd_r)
    tri = seq[[0, 4, 8]][:, :, :128, :192]
    xb, xc, xa = (tri[i:i + 1].float() / 255.0 for i in range(3))
    fout = {"frames_u8": tri.numpy()}
    with torch.no_grad():
        for tag, (n, l) in {"n1_l1": ([1], 1.0), "n0_l066": ([0], 0.66)}.items():
            r = rf(xb, xc, xa, n=n, l=l, train=False)
            o = of(xb, xc, xa, n=n, l=l, train=False)
            assert torch.equal(r["x_hat"], o["x_hat"]) and torch.equal(r["size"], o["size"])
            fout.update({f"{tag}_x_hat": r["x_hat"].numpy(), f"{tag}_size": r["size"].numpy(),
                         f"{tag}_rate": r["rate"].numpy()})
            print("flex", tag, "size", r["size"].tolist())
    np.savez_compressed(os.path.join(OUT, "flexrate_model_reference.npz"), **fout)
    print("flexrate_model_reference.npz written")


if __name__ == "__main__":
    main()
