"""TEST INFRASTRUCTURE (oracle).  Generates ``tests/golden/icip_reference.npz`` by running the REFERENCE'S OWN CODE
(read from ``/root/reference``, never copied) in the build container:

* ``ICIP2024/src/model/helpers.py`` class ``OffsetDiversity`` is pulled out of the source with ``ast`` and executed
  verbatim on top of the real ``torchvision.ops.DeformConv2d`` (the module itself imports fine here except for its
  siblings; only this class is needed);
* ``ICIP2024/src/model/compression_bottlenecks.py`` ``ste_round`` likewise.

Run:  python -m oracle.make_golden_icip      (needs /root/reference; the GPU box never runs this)
"""
import ast
import os
import textwrap

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.ops import DeformConv2d

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def extract(path, name, ns):
    """Compile one top-level class or function of a reference file, verbatim."""
    src = open(path).read()
    for node in ast.parse(src).body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name == name:
            code = textwrap.dedent(ast.get_source_segment(src, node))
            exec(compile(code, f"{path}:{name}", "exec"), ns)
            return ns[name]
    raise KeyError(f"{name} not found in {path}")


def main():
    torch.set_num_threads(1)
    ns = {"torch": torch, "nn": nn, "F": F, "DeformConv2d": DeformConv2d, "Tensor": torch.Tensor}
    OffsetDiversity = extract(os.path.join(REF, "ICIP2024/src/model/helpers.py"), "OffsetDiversity", ns)
    ste_round = extract(os.path.join(REF, "ICIP2024/src/model/compression_bottlenecks.py"), "ste_round", ns)

    torch.manual_seed(7)
    C, H, W = 16, 12, 16
    mod = OffsetDiversity(C, 10.0).eval()
    g = torch.Generator().manual_seed(8)
    r = lambda *s: torch.randn(*s, generator=g)
    x1, x2 = r(1, C, H, W), r(1, C, H, W)
    o1, o2 = r(1, 3 * 8 * 9, H, W), r(1, 3 * 8 * 9, H, W)       # chunk(3): two offset halves + mask, 8 groups x 9 taps
    f1, f2 = 3 * r(1, 2, H, W), 3 * r(1, 2, H, W)
    with torch.no_grad():
        out = mod(x1, o1, f1, x2, o2, f2)
        off1, m1 = mod.prep(o1, f1)
        warped = mod.warp(x1, f1)
    y = torch.tensor([-2.5, -1.5, -0.5, -0.3, 0.0, 0.3, 0.5, 1.5, 2.5, 3.49999, 1e6 + 0.5]) 
    y = torch.cat([y, 5 * r(53)])
    np.savez_compressed(
        os.path.join(OUT, "icip_reference.npz"),
        weight=mod.fusion.weight.detach().numpy(), bias=mod.fusion.bias.detach().numpy(),
        x1=x1.numpy(), x2=x2.numpy(), o1=o1.numpy(), o2=o2.numpy(), f1=f1.numpy(), f2=f2.numpy(),
        out=out.numpy(), off1=off1.numpy(), m1=m1.numpy(), warped=warped.numpy(),
        ste_in=y.numpy(), ste_out=ste_round(y).numpy())
    print("wrote", os.path.join(OUT, "icip_reference.npz"), "out", tuple(out.shape))


if __name__ == "__main__":
    main()
