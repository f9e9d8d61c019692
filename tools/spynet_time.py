#!/usr/bin/env python
"""Timing of the SPyNet level kernel (L2 flushed), gather vs TMA-staged (B200VC_SPYNET_TMA=0/1 in the environment)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
import torch
from b200vc import ops
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
g = torch.Generator().manual_seed(0)
print("B200VC_SPYNET_TMA =", os.environ.get("B200VC_SPYNET_TMA", "1"))
for (H, W) in ((1088, 1920), (544, 960), (272, 480)):
    for N in (1, 2, 4):
        a = torch.randn(N, 3, H, W, generator=g).cuda()
        b = torch.randn(N, 3, H, W, generator=g).cuda()
        flow = torch.nn.functional.interpolate(3.0 * torch.randn(N, 2, 17, 30, generator=g), size=(H // 2, W // 2), mode="bilinear").cuda()
        ops.spynet_level(a, b, flow)
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(10):
            flush.add_(1.0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); ops.spynet_level(a, b, flow); e.record()
            torch.cuda.synchronize()
            tot += s.elapsed_time(e)
        ms = tot / 10
        gb = 4 * N * (14 * H * W + 2 * (H // 2) * (W // 2)) / ms / 1e6
        print(f"N={N} {H}x{W}: {ms*1e3:7.1f} us  {gb:5.0f} GB/s ({gb/6539.2:.1%})", flush=True)
