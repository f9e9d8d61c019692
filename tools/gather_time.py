#!/usr/bin/env python
"""Iso timing of the gather-form kernels (L2 flushed): warp_f32 (C=3 gather path and many-channel), warp_sse (OJSP
search form, 4K), warp2_half_sse (ICIP search form), spynet_level (gather form)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
os.environ.setdefault("B200VC_WARP_TMA", "0")
os.environ.setdefault("B200VC_SPYNET_TMA", "0")
import torch
import torch.nn.functional as F
from b200vc import ops
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
smooth = lambda N, H, W, a: F.interpolate(a * torch.randn(N, 2, H // 32 + 1, W // 32 + 1, generator=g), size=(H, W), mode="bilinear").cuda()

def timeit(name, fn, nbytes):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(8):
        flush.add_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    ms = tot / 8
    print(f"{name:46s} {ms*1e3:8.1f} us {nbytes/ms/1e6:6.0f} GB/s ({nbytes/ms/1e6/6539.2:.1%})", flush=True)

for N in (1, 4):
    H, W = 1088, 1920
    img = torch.rand(N, 3, H, W, generator=g).cuda(); fl = smooth(N, H, W, 4.0)
    timeit(f"warp_f32 gather lhbdc [{N},3,{H},{W}]", lambda: ops.backwarp(img, fl, "lhbdc"), 32 * N * H * W)
    timeit(f"warp_f32 gather flex  [{N},3,{H},{W}]", lambda: ops.backwarp(img, fl, "flex"), 32 * N * H * W)
    x2 = torch.rand(N, 3, H, W, generator=g).cuda(); fl2 = smooth(N, H, W, 4.0); xc = torch.rand(N, 3, H, W, generator=g).cuda()
    timeit(f"warp2_half_sse ac1    [{N},3,{H},{W}]", lambda: ops.warp2_half_sse(img, x2, fl, fl2, xc, "ac1"), 52 * N * H * W)
    first = torch.rand(N, 3, H, W, generator=g).cuda(); prev = smooth(N, H // 2, W // 2, 2.0)
    timeit(f"spynet_level gather   [{N},8,{H},{W}]", lambda: ops.spynet_level(first, img, prev), 4 * N * (14 * H * W + 2 * (H // 2) * (W // 2)))
for (C, H, W) in ((64, 544, 960), (96, 272, 480), (128, 136, 240)):
    img = torch.rand(1, C, H, W, generator=g).cuda(); fl = smooth(1, H, W, 3.0)
    timeit(f"warp_f32 many-channel ac1 [1,{C},{H},{W}]", lambda: ops.backwarp(img, fl, "ac1"), (2 * C + 2) * 4 * H * W)
H, W = 2160, 3840
img = torch.rand(1, 3, H, W, generator=g).cuda(); fl = smooth(1, H, W, 5.0); xc = torch.rand(1, 3, H, W, generator=g).cuda()
timeit(f"warp_sse ac1 [1,3,{H},{W}]", lambda: ops.warp_sse(img, fl, xc, "ac1"), 32 * H * W)
