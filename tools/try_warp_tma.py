#!/usr/bin/env python
"""Debug helper: one TMA-staged warp launch on a smooth flow, compared with the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from b200vc import ops  # noqa: E402
from oracle import warp as o_warp  # noqa: E402

H, W = int(os.environ.get("H", 256)), int(os.environ.get("W", 256))
g = torch.Generator().manual_seed(0)
img = torch.rand(1, 3, H, W, generator=g).cuda()
f = 2.0 * torch.randn(1, 2, H // 16, W // 16, generator=g)
flow = torch.nn.functional.interpolate(f, size=(H, W), mode="bilinear", align_corners=False).cuda()
for variant, fn in (("lhbdc", o_warp.backwarp_lhbdc), ("flex", o_warp.backwarp_flex), ("ac1", o_warp.warp_ac1)):
    got = ops.backwarp(img, flow, variant)
    torch.cuda.synchronize()
    want = fn(img, flow)
    print(variant, "max|diff|", (got - want).abs().max().item(), "bit-exact", (got == want).float().mean().item(), flush=True)
