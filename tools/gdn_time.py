#!/usr/bin/env python
"""Timing only: tcgen05 GDN at [4,128,544,960], plain and in-place residual forms (L2 flushed between launches)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from b200vc import modules, ops  # noqa: E402

params = modules.gdn_params(modules.GDN(128).cuda().eval())
flush_buf = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
shapes = [(4, 544, 960)] if "--all" not in sys.argv else [(1, 544, 960), (4, 544, 960), (2, 272, 480), (4, 136, 240)]
for (N, H, W) in shapes:
    x = torch.randn(N, 128, H, W, device="cuda")
    skip = torch.randn_like(x)
    for name, fn, words in (("plain", lambda: ops.gdn(x, params, impl=2), 2),
                            ("inverse", lambda: ops.gdn(x, params, inverse=True, impl=2), 2),
                            ("residual in place", lambda: ops.gdn(x, params, addend=skip, impl=2), 3)):
        fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(10):
            flush_buf.add_(1.0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            tot += s.elapsed_time(e)
        ms = tot / 10
        gb = words * 512 * N * H * W / ms / 1e6
        print(f"impl=2 {name:18s} N={N} {H}x{W}: {ms*1e3:7.1f} us  {gb:5.0f} GB/s", flush=True)
