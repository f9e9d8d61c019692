#!/usr/bin/env python
"""One K-DCN launch at the ICIP2024 fusion shape (ncu target)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
import torch
from b200vc import ops
g = torch.Generator().manual_seed(0)
Cin, Cout, H, W, groups = 128, 64, 544, 960, 16
x = torch.randn(1, Cin, H, W, generator=g).cuda()
w = (torch.randn(Cout, Cin // groups, 3, 3, generator=g) * 0.1).cuda()
b = torch.randn(Cout, generator=g).cuda()
off = torch.nn.functional.interpolate(2.0 * torch.randn(1, 2 * groups * 9, H // 8, W // 8, generator=g).cuda(), size=(H, W), mode="bilinear")
m = torch.sigmoid(torch.randn(1, groups * 9, H, W, generator=g)).cuda()
for _ in range(2):
    ops.deform_conv2d(x, off, w, b, padding=(1, 1), mask=m)
torch.cuda.synchronize()
print("done")
