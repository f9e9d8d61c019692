#!/usr/bin/env python
"""Small odd-shaped launches of the round-2 K-GC kernel (persistent CTAs, per-thread cp.async ring, packed fp32): target for
`compute-sanitizer --tool memcheck` and `--tool racecheck`.  Also checks every output against the oracle chain."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
import torch
from b200vc import ops
from oracle import cai
g = torch.Generator().manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g).cuda()
table = torch.exp(torch.linspace(torch.log(torch.tensor(0.11)), torch.log(torch.tensor(256.0)), 64)).cuda()
gc = cai.GaussianConditional(None).cuda().eval()
# (N, C, H, W): vector path (HW % 4 == 0) and scalar path, one slot per sample, many samples per CTA, slots with a
# partial last chunk, sizes around the 1024- and 4096-element chunk / slot boundaries
shapes = ((1, 1, 1, 4), (3, 1, 1, 3), (2, 7, 5, 9), (2, 8, 16, 8), (5, 3, 33, 31), (1, 16, 17, 60), (300, 2, 2, 2),
          (2, 128, 20, 32), (1, 5, 205, 4), (3, 1, 64, 64), (1, 1, 64, 65), (2, 12, 34, 60))
for (N, C, H, W) in shapes:
    y, s, m = 4 * r(N, C, H, W), (r(N, C, H, W).abs() * 3), r(N, C, H, W)
    s[0, 0, 0, 0] = 0.0            # clamps to the scale bound
    y.view(-1)[-1] = 60.0          # far tail: the guarded erfc form
    out = ops.gauss_cond(y, s, m, want_lik=True, want_bits=True, want_symbols=True, scale_table=table)
    lean = ops.gauss_cond(y, s, m, want_lik=False, want_bits=True)
    yh, lik = gc(y, s, means=m)
    assert torch.equal(out["y_hat"], yh) and torch.equal(lean["y_hat"], yh), (N, C, H, W)
    assert torch.equal(out["lik"], lik), ((out["lik"] - lik).abs().max().item(), (N, C, H, W))
    ref = -torch.log2(lik.double()).flatten(1).sum(1)
    for o in (out, lean):
        assert ((o["bits"] - ref).abs() <= 1e-6 * ref.abs() + 1e-6).all(), (o["bits"], ref)
torch.cuda.synchronize()
print("done", len(shapes), "shapes")
