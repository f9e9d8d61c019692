#!/usr/bin/env python
"""Per-tile pipeline timeline of the tcgen05 GDN kernel (CTA 0), from clock64() stamps written by the kernel itself.
Slots: 0 load issued, 1 raw landed, 2 split starts, 3 MMA issue starts, 4 epilogue starts (accumulators complete),
5 store issued, 6 previous slot released."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from b200vc import _lib, modules, ops  # noqa: E402

lib = _lib.load()
if not hasattr(lib, "b200vc_debug_set_gdn_trace"):
    raise SystemExit("gdn_trace.py needs a debug build: NVCC_FLAGS=-DB200VC_ENABLE_GDN_TRACE python "
                     "video-compression_b200/build.py --force   (include/b200vc_debug.h)")
p = modules.GDN(128).cuda().eval()
params = modules.gdn_params(p)
x = torch.randn(1, 128, 544, 960, device="cuda")
skip = torch.randn_like(x)
names = ["load issued", "raw landed", "split starts", "MMA issue", "epilogue starts", "store issued", "slot released"]
for label, kw in (("plain", {}), ("residual (TMA reduce-add)", {"addend": skip})):
    ops.gdn(x, params, **kw)
    buf = torch.zeros(256 * 16, dtype=torch.int64, device="cuda")
    lib.b200vc_debug_set_gdn_trace(buf.data_ptr())
    ops.gdn(x, params, **kw)
    torch.cuda.synchronize()
    lib.b200vc_debug_set_gdn_trace(None)
    t = buf.view(256, 16).cpu().double()
    n = int((t[:, 3] > 0).sum())
    t = t[:n]
    print(f"== {label}: {n} tiles on CTA 0")
    period = (t[8:n - 4, 3][1:] - t[8:n - 4, 3][:-1])
    print(f"   tile period (MMA issue to MMA issue): mean {period.mean():.0f} clk, min {period.min():.0f}, max {period.max():.0f}")
    k = slice(8, n - 4)
    d = lambda a, b: (t[k, b] - t[k, a]).mean().item()
    print(f"   load issued -> raw landed      {d(0, 1):7.0f} clk")
    print(f"   raw landed  -> split starts    {d(1, 2):7.0f} clk   (waiting for the operand slot = previous MMA)")
    print(f"   split starts -> MMA issue      {d(2, 3):7.0f} clk   (square + hi/lo split + fence + barrier)")
    print(f"   MMA issue -> epilogue starts   {d(3, 4):7.0f} clk   (48 tcgen05.mma + commit)")
    print(f"   epilogue starts -> store issued{d(4, 5):7.0f} clk   (tcgen05.ld, rsqrt, in-place result, fence, bar)")
    print(f"   store issued -> slot released  {d(5, 6):7.0f} clk   (wait_group.read of the previous store)")
    print(f"   load issued -> store issued    {d(0, 5):7.0f} clk   (raw slot lifetime without the deferred release)")
    print("   -- detail")
    print(f"   epilogue: d_full seen -> accumulators in registers {d(4, 8):7.0f} clk")
    print(f"   epilogue: -> result written to raw slot            {d(8, 9):7.0f} clk")
    print(f"   epilogue: -> fence.proxy.async done                {d(9, 10):7.0f} clk")
    print(f"   epilogue: -> bar.sync passed, store issued         {d(10, 5):7.0f} clk")
    print(f"   split: starts -> stores issued                     {d(2, 12):7.0f} clk")
    print(f"   split: -> fence done                               {d(12, 13):7.0f} clk")
    print(f"   split fence done -> MMA issue                      {d(13, 3):7.0f} clk")
    print(f"   mma: accumulator stage free -> operands ready      {d(14, 3):7.0f} clk")
    print(f"   mma: issue of 48 MMAs + 2 commits                  {d(3, 15):7.0f} clk")
    print(f"   mma: issued -> epilogue sees d_full                {d(15, 4):7.0f} clk")
    ep = (t[8:n - 4, 4][1:] - t[8:n - 4, 4][:-1])
    print(f"   epilogue period mean {ep.mean():.0f} clk;  epilogue busy (d_full seen -> slot released) {d(4, 6):.0f} clk")
    nxt = (t[9:n - 3, 4] - t[8:n - 4, 6]).mean().item()
    print(f"   epilogue idle between tiles (slot released -> next d_full seen) {nxt:7.0f} clk")
