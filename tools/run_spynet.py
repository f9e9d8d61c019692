#!/usr/bin/env python
"""Launch the SPyNet glue kernels at their 1080p shapes (target for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
import torch
from b200vc import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
g = torch.Generator().manual_seed(0)
x1 = torch.rand(N, 3, 1088, 1920, generator=g).cuda()
x2 = torch.rand(N, 3, 1088, 1920, generator=g).cuda()
flow = torch.nn.functional.interpolate(2.0 * torch.randn(N, 2, 34, 60, generator=g), size=(544, 960), mode="bilinear").cuda()
for _ in range(2):
    p1 = ops.spynet_pyramid(x1)
    p2 = ops.spynet_pyramid(x2)
    feat = ops.spynet_level(p1[-1], p2[-1], flow)
torch.cuda.synchronize()
print("done")
