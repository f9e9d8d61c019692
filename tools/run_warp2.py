#!/usr/bin/env python
"""A few launches of ops.warp2_lhbdc at [N,3,1088,1920] on smooth flows, for ncu (B200VC_WARP2_V2=0/1 picks the kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
import torch
from b200vc import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
g = torch.Generator().manual_seed(0)
H, W = 1088, 1920
xb = torch.rand(N, 3, H, W, generator=g).cuda()
xa = torch.rand(N, 3, H, W, generator=g).cuda()
sm = lambda c, a: torch.nn.functional.interpolate(a * torch.randn(N, c, 20, 32, generator=g), size=(272, 480), mode="bilinear").cuda()
fh, fab, fba = sm(4, 2.0), sm(2, 1.5), sm(2, 1.5)
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
for _ in range(3):
    flush.add_(1.0)
    ops.warp2_lhbdc(xb, xa, fh, fab, fba)
torch.cuda.synchronize()
