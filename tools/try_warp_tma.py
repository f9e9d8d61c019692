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

# fused LHBDC warp2 (TMA-staged when B200VC_WARP2_TMA=1) vs the unfused oracle chain
hh, ww = H // 4, W // 4
h4, w4 = hh + (64 - hh % 64) % 64, ww + (64 - ww % 64) % 64
xb, xa = torch.rand(2, 3, H, W, generator=g).cuda(), torch.rand(2, 3, H, W, generator=g).cuda()
smooth = lambda c, amp: torch.nn.functional.interpolate(amp * torch.randn(2, c, h4 // 8, w4 // 8, generator=g),
                                                        size=(h4, w4), mode="bilinear", align_corners=False).cuda()
fh, fab, fba = smooth(4, 2.0), smooth(2, 1.5), smooth(2, 1.5)
cb, ca = o_warp.lhbdc_flow_glue(fh, fab, fba, hh, ww)
want = torch.cat([o_warp.backwarp_lhbdc(xb, cb), o_warp.backwarp_lhbdc(xa, ca)], 1)
got, flows = ops.warp2_lhbdc(xb, xa, fh, fab, fba, return_flows=True)
torch.cuda.synchronize()
print("warp2 image bit-exact", (got == want).float().mean().item(), "flows bit-exact",
      (flows == torch.cat([cb, ca], 1)).float().mean().item(), flush=True)
assert torch.equal(got, want) and torch.equal(flows, torch.cat([cb, ca], 1))
print("OK")
