#!/usr/bin/env python
"""Launches every hot-path kernel a few times at its 1080p shape -- the target of
`ncu --set full -k regex:b200vc` captures.  Not a benchmark (numbers under a profiler are never reported)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from b200vc import modules, ops  # noqa: E402
from gpu_util import gc_case  # noqa: E402

which = set(sys.argv[1:]) or {"gdn", "warp", "warp2", "blend", "gc", "eb", "spynet", "sse", "dcn"}
reps = int(os.environ.get("REPS", "1"))
g = torch.Generator().manual_seed(0)
N, H, W = 1, 1088, 1920


def smooth_flow(n, amp=3.0):
    f = amp * torch.randn(n, 2, H // 16, W // 16, generator=g)
    return torch.nn.functional.interpolate(f, size=(H, W), mode="bilinear", align_corners=False).cuda()


if "gdn" in which:
    p = modules.GDN(128).cuda().eval()
    params = modules.gdn_params(p)
    x = torch.randn(1, 128, 544, 960, device="cuda")
    skip = torch.randn_like(x)
    for _ in range(reps):
        ops.gdn(x, params)
        ops.gdn(x, params, inverse=True, addend=skip)  # in-place residual form (TMA reduce-add)
if "warp" in which:
    img = torch.rand(N, 3, H, W, generator=g).cuda()
    flow = smooth_flow(N)
    for _ in range(reps):
        ops.backwarp(img, flow, "lhbdc")
        ops.backwarp(img, flow, "flex")
if "warp2" in which:
    xb = torch.rand(N, 3, H, W, generator=g).cuda()
    xa = torch.rand(N, 3, H, W, generator=g).cuda()
    fh = (2.0 * torch.randn(N, 4, 320, 512, generator=g)).cuda()
    fab = (1.5 * torch.randn(N, 2, 320, 512, generator=g)).cuda()
    fba = (1.5 * torch.randn(N, 2, 320, 512, generator=g)).cuda()
    for _ in range(reps):
        ops.warp2_lhbdc(xb, xa, fh, fab, fba)
if "blend" in which:
    mask = torch.rand(N, 1, H, W, device="cuda")
    both = torch.rand(N, 6, H, W, device="cuda")
    x = torch.rand(N, 3, H, W, device="cuda")
    for _ in range(reps):
        ops.blend_residual("mask", mask, both[:, :3], both[:, 3:], x)
if "spynet" in which:
    x1 = torch.rand(4, 3, H, W, generator=g).cuda()
    x2 = torch.rand(4, 3, H, W, generator=g).cuda()
    fl = torch.nn.functional.interpolate(2.0 * torch.randn(4, 2, 34, 60, generator=g), size=(H // 2, W // 2),
                                         mode="bilinear").cuda()
    for _ in range(reps):
        p1, p2 = ops.spynet_pyramid(x1), ops.spynet_pyramid(x2)
        ops.spynet_level(p1[-1], p2[-1], fl)          # TMA-staged (>= 2 Mpx per launch)
        ops.spynet_level(p1[-2][:1], p2[-2][:1], fl[:1, :, ::2, ::2].contiguous())  # gather kernel
if "sse" in which:
    img = torch.rand(1, 3, 2160, 3840, generator=g).cuda()
    cur = torch.rand(1, 3, 2160, 3840, generator=g).cuda()
    f4k = torch.nn.functional.interpolate(3.0 * torch.randn(1, 2, 135, 240, generator=g), size=(2160, 3840),
                                          mode="bilinear").cuda()
    for _ in range(reps):
        ops.warp_sse(img, f4k, cur, "ac1")
if "dcn" in which:
    Cin, Cout, h, w, groups = 128, 64, 544, 960, 16
    xd = torch.randn(1, Cin, h, w, generator=g).cuda()
    wd = (torch.randn(Cout, Cin // groups, 3, 3, generator=g) * 0.1).cuda()
    od = torch.nn.functional.interpolate(2.0 * torch.randn(1, 2 * groups * 9, h // 8, w // 8, generator=g).cuda(),
                                         size=(h, w), mode="bilinear")
    md = torch.sigmoid(torch.randn(1, groups * 9, h, w, generator=g)).cuda()
    for _ in range(reps):
        ops.deform_conv2d(xd, od, wd, None, padding=(1, 1), mask=md)
if "gc" in which:
    y, s, m = gc_case(3, 4, 128, 68, 120)
    for _ in range(reps):
        ops.gauss_cond(y, s, m, want_lik=False)
if "eb" in which:
    pe = modules.EntropyBottleneck(128).cuda().eval()
    packed = modules.eb_packed(pe)
    z = 3 * torch.randn(4, 128, 17, 30, device="cuda")
    for _ in range(reps):
        ops.entropy_bottleneck(z, packed, want_lik=False)
torch.cuda.synchronize()
print("done")
