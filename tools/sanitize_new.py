#!/usr/bin/env python
"""Small odd-shaped launches of the kernels added late in round 1 (target for compute-sanitizer memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
import torch
from b200vc import modules, ops
g = torch.Generator().manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g).cuda()
for (N, H, W) in ((1, 70, 131), (2, 33, 20), (1, 192, 256), (1, 24, 30)):
    ops.spynet_pyramid(torch.rand(N, 3, H, W, generator=g).cuda())
for (N, H, W) in ((1, 35, 61), (2, 34, 60), (1, 17, 31)):
    ops.spynet_level(r(N, 3, H, W), r(N, 3, H, W), 3 * r(N, 2, H // 2, W // 2))
    ops.spynet_level(r(N, 3, H, W), r(N, 3, H, W), None)
# TMA-staged level kernel: >= 2 Mpx, W % 4 == 0; rough flow forces some CTAs through the in-kernel gather fallback
ops.spynet_level(r(2, 3, 1088, 1920), r(2, 3, 1088, 1920), 6 * r(2, 2, 544, 960))
ops.warp_sse(r(2, 3, 37, 53), 4 * r(2, 2, 37, 53), r(2, 3, 37, 53), "ac1", want_pred=True)
ops.warp_sse(r(1, 3, 40, 64), 40 * r(1, 2, 40, 64), r(1, 3, 40, 64), "flex")
for (Cin, Cout, H, W, k, s, p, d, groups, og) in ((16, 8, 21, 23, 3, 2, 1, 1, 8, 2), (64, 64, 19, 27, 3, 1, 1, 1, 8, 8),
                                                  (128, 64, 18, 30, 3, 1, 1, 1, 16, 16), (10, 15, 12, 14, 5, 1, 2, 1, 5, 2)):
    Ho = (H + 2 * p - (d * (k - 1) + 1)) // s + 1
    Wo = (W + 2 * p - (d * (k - 1) + 1)) // s + 1
    off = 3 * r(1, 2 * og * k * k, Ho, Wo)
    off[:, :, 0, 0] *= 50
    for fast in (True, False):
        ops.deform_conv2d(r(1, Cin, H, W), off, r(Cout, Cin // groups, k, k), r(Cout), stride=(s, s), padding=(p, p),
                          dilation=(d, d), mask=torch.sigmoid(r(1, og * k * k, Ho, Wo)), use_workspace=fast)
y = 3 * r(2, 7, 5, 9)
ops.round_checker(y)
buf = torch.zeros(2, 12, 5, 9, device="cuda")
ops.checker_mask(y[:, :5].contiguous(), out=buf[:, 3:8])
p = modules.gdn_params(modules.GDN(128).cuda().eval())
for (N, H, W) in ((1, 5, 8), (2, 33, 44)):
    x = r(N, 128, H, W)
    ops.gdn(x, p, impl=2)
    ops.gdn(x, p, inverse=True, addend=r(N, 128, H, W), impl=2)
torch.cuda.synchronize()
print("done")
