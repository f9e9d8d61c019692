#!/usr/bin/env python
"""AC1 feature warps of the ICIP codecs ([1,64,544,960], [1,96,272,480], [1,128,136,240]), L2 flushed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
import torch
from b200vc import ops
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
g = torch.Generator().manual_seed(0)
for (C, H, W) in ((64, 544, 960), (96, 272, 480), (128, 136, 240), (3, 1088, 1920)):
    for N in (1, 2):
        img = torch.randn(N, C, H, W, generator=g).cuda()
        for kind in ("smooth", "noise"):
            if kind == "smooth":
                flow = torch.nn.functional.interpolate(4 * torch.randn(N, 2, H // 16, W // 16, generator=g), size=(H, W), mode="bilinear").cuda()
            else:
                flow = (3 * torch.randn(N, 2, H, W, generator=g)).cuda()
            ops.backwarp(img, flow, "ac1"); torch.cuda.synchronize()
            tot = 0.0
            for _ in range(8):
                flush.add_(1.0)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); ops.backwarp(img, flow, "ac1"); e.record(); torch.cuda.synchronize()
                tot += s.elapsed_time(e)
            ms = tot / 8
            gb = (2 * C + 2) * 4 * N * H * W / ms / 1e6
            print(f"C={C:3d} N={N} {H}x{W} {kind:6s}: {ms*1e3:7.1f} us {gb:5.0f} GB/s ({gb/6539.2:.1%})", flush=True)
