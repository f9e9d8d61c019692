#!/usr/bin/env python
"""K-DCN vs torchvision.ops.deform_conv2d on the reference's layer shapes at 1080p feature resolutions (L2 flushed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
import torch
import torchvision.ops as tv
from b200vc import ops
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")


def timeit(fn):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(5):
        flush.add_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / 5


for name, (Cin, Cout, H, W, groups) in {
        "ICIP2024 fusion l1 128->64": (128, 64, 544, 960, 16), "ICIP2024 fusion l2 192->96": (192, 96, 272, 480, 16),
        "ICIP2024 fusion l3 256->128": (256, 128, 136, 240, 16), "ICIP2023 l1 32->32": (32, 32, 544, 960, 8),
        "ICIP2023 l2 64->64": (64, 64, 272, 480, 8), "ICIP2023 l3 96->96": (96, 96, 136, 240, 8)}.items():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, Cin, H, W, generator=g).cuda()
    w = (torch.randn(Cout, Cin // groups, 3, 3, generator=g) * 0.1).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    m = torch.sigmoid(torch.randn(1, groups * 9, H, W, generator=g)).cuda()
    nbytes = 4 * (x.numel() + 3 * m.numel() + Cout * H * W)
    for kind in ("white-noise offsets (sigma 2 px)", "smooth offsets (upsampled x8, sigma 2 px)"):
        if kind.startswith("white"):
            off = (2.0 * torch.randn(1, 2 * groups * 9, H, W, generator=g)).cuda()
        else:
            off = torch.nn.functional.interpolate(2.0 * torch.randn(1, 2 * groups * 9, H // 8, W // 8, generator=g).cuda(),
                                                  size=(H, W), mode="bilinear")
        t_tv = timeit(lambda: tv.deform_conv2d(x, off, w, b, padding=(1, 1), mask=m))
        t_me = timeit(lambda: ops.deform_conv2d(x, off, w, b, padding=(1, 1), mask=m))
        t_nc = timeit(lambda: ops.deform_conv2d(x, off, w, b, padding=(1, 1), mask=m, use_workspace=False))
        print(f"{name:28s} [{H}x{W}] {kind:42s}: torchvision {t_tv*1e3:7.1f} us | b200vc {t_me*1e3:7.1f} us "
              f"({nbytes/t_me/1e6:5.0f} GB/s) x{t_tv/t_me:.1f} | NCHW kernel {t_nc*1e3:7.1f} us", flush=True)
