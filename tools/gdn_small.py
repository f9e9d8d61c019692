#!/usr/bin/env python
"""fp32 vs tcgen05 GDN on the small shapes of the bench step (crossover for the auto dispatch)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from b200vc import modules, ops  # noqa: E402

params = modules.gdn_params(modules.GDN(128).cuda().eval())
flush_buf = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
for (H, W) in [(20, 32), (40, 64), (80, 128), (136, 240), (160, 256)]:
    for N in (1, 2, 4):
        x = torch.randn(N, 128, H, W, device="cuda")
        skip = torch.randn_like(x)
        row = f"N={N} {H}x{W} ({N*H*W:7d} positions):"
        for impl in (1, 2):
            for name, fn in (("plain", lambda: ops.gdn(x, params, impl=impl)),
                             ("resid", lambda: ops.gdn(x, params, addend=skip, impl=impl))):
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
                row += f"  impl={impl} {name} {tot*100:6.1f} us"
        print(row, flush=True)
