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

# ---- structural diagnostics: gamma = I, beta = 0, debug mode 2 returns norm = gamma @ x^2 ------------------
def manual_params(gam):
    Cc = gam.shape[0]
    hi = gam.clone()
    return torch.cat([torch.zeros(Cc), gam.flatten(), gam.t().contiguous().flatten(), hi.flatten(),
                      torch.zeros(Cc * Cc)]).cuda()


def raw_norm(x, prm, impl):
    out = torch.empty_like(x)
    from b200vc import _lib
    lib = _lib.load()
    rc = lib.b200vc_gdn_f32(x.data_ptr(), prm.data_ptr(), None, out.data_ptr(), x.shape[0], x.shape[1],
                            x.shape[2] * x.shape[3], 2, impl, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    return out


eye = manual_params(torch.eye(C))
xs = torch.zeros(1, C, 8, 16).cuda()  # 128 positions = 2 tiles
probes = [(0, 0), (1, 0), (5, 3), (9, 37), (77, 64), (127, 127)]
for (j0, p0) in probes:
    xs.zero_()
    xs.view(C, 128)[j0, p0] = 2.0
    nz = raw_norm(xs, eye, 2).view(C, 128).nonzero().tolist()
    print(f"one-hot x[j={j0}, p={p0}] with gamma=I -> norm nonzero at {nz[:6]} (expect [[{j0}, {p0}]])", flush=True)
perm = torch.randperm(C, generator=g)
pm = torch.zeros(C, C)
pm[torch.arange(C), perm] = 1.0  # norm[i] = x^2[perm[i]]
xr = torch.randn(1, C, 8, 16, generator=g).cuda()
got = raw_norm(xr, manual_params(pm), 2).view(C, 128)
want = (xr.view(C, 128) ** 2)[perm.cuda()]
print("gamma = permutation: max abs err", (got - want).abs().max().item(),
      "(tf32-rounded x^2 expected: ~1e-3 rel since lo image is zero)", flush=True)

for (N, H, W) in [(1, 8, 8), (1, 5, 8), (1, 16, 20), (2, 33, 44), (1, 68, 120), (3, 136, 240), (1, 544, 960)]:
    scale = torch.exp(torch.empty(1, C, 1, 1).uniform_(-2.3, 2.3, generator=g))
    x = (torch.randn(N, C, H, W, generator=g) * scale).cuda()
    skip = torch.randn(N, C, H, W, generator=g).cuda()
    for inverse in (False, True):
        ref = ops.gdn(x, params, inverse=inverse, impl=1)
        got = ops.gdn(x, params, inverse=inverse, impl=2)
        torch.cuda.synchronize()
        rel = ((got - ref).abs() / ref.abs().clamp(min=1e-6)).max().item()
        got2 = ops.gdn(x, params, inverse=inverse, addend=skip.clone(), impl=2)
        ref2 = ops.gdn(x, params, inverse=inverse, addend=skip.clone(), impl=1)
        torch.cuda.synchronize()
        abs2 = (got2 - ref2).abs().max().item()
        print(f"N={N} {H}x{W} inverse={inverse}: max rel err vs fp32 kernel {rel:.3e}; with addend max abs {abs2:.3e}; "
              f"finite={torch.isfinite(got).all().item()}", flush=True)

flush_buf = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
print("B200VC_GDN_TC_CFG =", os.environ.get("B200VC_GDN_TC_CFG", "0"))
for (N, H, W) in [(1, 544, 960), (4, 544, 960), (1, 272, 480), (1, 136, 240)]:
    x = torch.randn(N, C, H, W, device="cuda")
    skip = torch.randn(N, C, H, W, device="cuda")
    out_sep = torch.empty_like(x)
    from b200vc import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream

    def sep():  # addend read from a separate tensor
        lib.b200vc_gdn_f32(x.data_ptr(), params.data_ptr(), skip.data_ptr(), out_sep.data_ptr(), N, C, H * W, 0, 2, st)

    cases = [("impl=1 plain", lambda: ops.gdn(x, params, impl=1), 2), ("impl=2 plain", lambda: ops.gdn(x, params, impl=2), 2),
             ("impl=2 +addend in place (TMA reduce-add)", lambda: ops.gdn(x, params, addend=skip, impl=2), 3),
             ("impl=2 +addend separate", sep, 3)]
    for name, fn, words in cases:
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
        print(f"{name:42s} N={N} {H}x{W}: {ms*1e3:7.1f} us  {gb:5.0f} GB/s  ({gb/6539.2:.1%} of measured peak)", flush=True)
