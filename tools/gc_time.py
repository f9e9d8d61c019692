#!/usr/bin/env python
"""Iso timing of K-GC (gauss_cond_f32) at the 1080p latent shape [N,128,68,120] (L2 flushed), bits-only mode
(y_hat + fp64 bit totals: 16 B/element, what Model.forward runs) and coding mode (+ likelihoods, symbols, indexes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
import torch
from b200vc import ops
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
table = torch.exp(torch.linspace(torch.log(torch.tensor(0.11)), torch.log(torch.tensor(256.0)), 64)).cuda()

def timeit(name, fn, nbytes):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.add_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:44s} {ms*1e3:8.1f} us {nbytes/ms/1e6:6.0f} GB/s ({nbytes/ms/1e6/6539.2:.1%})", flush=True)

for N in (1, 2, 4, 8, 16, 32):
    C, H, W = 128, 68, 120
    y = (3.0 * torch.randn(N, C, H, W, generator=g)).cuda()
    sc = (torch.rand(N, C, H, W, generator=g) * 4.0).cuda()
    mu = torch.randn(N, C, H, W, generator=g).cuda()
    n = y.numel()
    timeit(f"gauss_cond bits-only [{N},{C},{H},{W}]",
           lambda: ops.gauss_cond(y, sc, mu, want_lik=False, want_bits=True), 16 * n)
    if N in (1, 4, 16):
        timeit(f"gauss_cond coding    [{N},{C},{H},{W}]",
               lambda: ops.gauss_cond(y, sc, mu, want_lik=True, want_bits=True, want_symbols=True, scale_table=table), 28 * n)
