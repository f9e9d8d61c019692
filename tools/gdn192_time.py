#!/usr/bin/env python
"""Timing: C = 192 GDN on the tcgen05 kernel (impl 2) vs the exact CUDA-core kernel (impl 1), plain / inverse / in-place
residual forms, at the I-frame codec's 1080p layer shapes (L2 flushed between launches)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from b200vc import modules, ops  # noqa: E402

C = 192
params = modules.gdn_params(modules.GDN(C).cuda().eval())
flush_buf = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
for (N, H, W) in [(1, 544, 960), (1, 272, 480), (4, 272, 480), (1, 136, 240), (1, 68, 120)]:
    x = torch.randn(N, C, H, W, device="cuda")
    skip = torch.randn_like(x)
    for impl in (2, 1):
        for name, fn, words in (("plain", lambda: ops.gdn(x, params, impl=impl), 2),
                                ("inverse", lambda: ops.gdn(x, params, inverse=True, impl=impl), 2),
                                ("residual in place", lambda: ops.gdn(x, params, addend=skip, impl=impl), 3)):
            if impl == 1 and name != "plain":
                continue
            fn()
            torch.cuda.synchronize()
            tot = 0.0
            for _ in range(8):
                flush_buf.add_(1.0)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn()
                e.record()
                torch.cuda.synchronize()
                tot += s.elapsed_time(e)
            ms = tot / 8
            gb = words * 4 * C * N * H * W / ms / 1e6
            print(f"C=192 impl={impl} {name:18s} N={N} {H}x{W}: {ms*1e3:7.1f} us  {gb:5.0f} GB/s ({gb/6539.2:.1%})", flush=True)
