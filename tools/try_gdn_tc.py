#!/usr/bin/env python
"""First-light check of the tcgen05 GDN kernel (run under gpurun, wrapped in `timeout`): each shape in its own
launch, compared with the exact fp32 CUDA-core kernel and the oracle; then timing at the 1080p shapes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from b200vc import modules, ops  # noqa: E402
from oracle import cai  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator().manual_seed(0)
C = 128
o = cai.GDN(C)
with torch.no_grad():
    ped = o.beta_reparam.pedestal
    beta = 1.0 + 0.1 * torch.randn(C, generator=g).abs()
    gamma = 0.1 * torch.eye(C) + 0.01 * torch.randn(C, C, generator=g).abs()
    o.beta.copy_(torch.sqrt(torch.max(beta + ped, ped)))
    o.gamma.copy_(torch.sqrt(torch.max(gamma + ped, ped)))
p = modules.GDN(C)
p.load_state_dict(o.state_dict())
o, p = o.cuda().eval(), p.cuda().eval()
params = modules.gdn_params(p)
torch.cuda.synchronize()
print("params ready", flush=True)

for (N, H, W) in [(1, 8, 8), (1, 5, 8), (1, 16, 20), (2, 33, 44), (1, 68, 120), (3, 136, 240), (1, 544, 960)]:
    scale = torch.exp(torch.empty(1, C, 1, 1).uniform_(-2.3, 2.3, generator=g))
    x = (torch.randn(N, C, H, W, generator=g) * scale).cuda()
    skip = torch.randn(N, C, H, W, generator=g).cuda()
    for inverse in (False, True):
        ref = ops.gdn(x, params, inverse=inverse, impl=1)
        got = ops.gdn(x, params, inverse=inverse, impl=2)
        torch.cuda.synchronize()
        rel = ((got - ref).abs() / ref.abs().clamp(min=1e-6)).max().item()
        got2 = ops.gdn(x, params, inverse=inverse, addend=skip, impl=2)
        ref2 = ops.gdn(x, params, inverse=inverse, addend=skip, impl=1)
        torch.cuda.synchronize()
        abs2 = (got2 - ref2).abs().max().item()
        print(f"N={N} {H}x{W} inverse={inverse}: max rel err vs fp32 kernel {rel:.3e}; with addend max abs {abs2:.3e}; "
              f"finite={torch.isfinite(got).all().item()}", flush=True)

flush_buf = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
for (N, H, W) in [(1, 544, 960), (4, 544, 960), (1, 272, 480), (1, 136, 240)]:
    x = torch.randn(N, C, H, W, device="cuda")
    for impl in (1, 2):
        ops.gdn(x, params, impl=impl)
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(10):
            flush_buf.add_(1.0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            ops.gdn(x, params, impl=impl)
            e.record()
            torch.cuda.synchronize()
            tot += s.elapsed_time(e)
        ms = tot / 10
        gb = 1024 * N * H * W / ms / 1e6
        print(f"impl={impl} N={N} {H}x{W}: {ms*1e3:.1f} us  {gb:.0f} GB/s  ({gb/6539.2:.1%} of measured peak)", flush=True)
