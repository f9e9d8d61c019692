#!/usr/bin/env python
"""One K-GC launch per batch size at the 1080p latent shape -- the target of an `ncu --set full -k regex:gauss_cond` capture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
import torch
from b200vc import ops
g = torch.Generator().manual_seed(0)
for N in [int(a) for a in sys.argv[1:]] or [1, 16]:
    y = (3.0 * torch.randn(N, 128, 68, 120, generator=g)).cuda()
    sc = (torch.rand(N, 128, 68, 120, generator=g) * 4.0).cuda()
    mu = torch.randn(N, 128, 68, 120, generator=g).cuda()
    for _ in range(2):
        ops.gauss_cond(y, sc, mu, want_lik=False, want_bits=True)
torch.cuda.synchronize()
