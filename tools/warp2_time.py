#!/usr/bin/env python
"""warp2_lhbdc: gather kernel vs TMA-staged kernel (B200VC_WARP2_TMA=0/1) on smooth and white-noise quarter-res flows."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
import torch
from b200vc import ops
g = torch.Generator().manual_seed(0)
H, W = 1088, 1920
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
tag = "pipe=" + os.environ.get("B200VC_WARP2_PIPE", "1") + " tma=" + os.environ.get("B200VC_WARP2_TMA", "0")
for N in [int(a) for a in sys.argv[1:]] or (1, 4, 16):
    xb = torch.rand(N, 3, H, W, generator=g).cuda()
    xa = torch.rand(N, 3, H, W, generator=g).cuda()
    sm = lambda c, a: torch.nn.functional.interpolate(a * torch.randn(N, c, 20, 32, generator=g), size=(272, 480), mode="bilinear").cuda()
    nz = lambda c, a: (a * torch.randn(N, c, 272, 480, generator=g)).cuda()
    for kind, mk in (("smooth", sm), ("noise", nz)):
        fh, fab, fba = mk(4, 2.0), mk(2, 1.5), mk(2, 1.5)
        ops.warp2_lhbdc(xb, xa, fh, fab, fba)
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(8):
            flush.add_(1.0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); ops.warp2_lhbdc(xb, xa, fh, fab, fba); e.record(); torch.cuda.synchronize()
            tot += s.elapsed_time(e)
        ms = tot / 8
        gb = 50 * N * H * W / ms / 1e6
        print(f"{tag} N={N} {kind:6s}: {ms*1e3:7.1f} us {gb:5.0f} GB/s ({gb/6539.2:.1%})", flush=True)
