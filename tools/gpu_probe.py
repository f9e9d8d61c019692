#!/usr/bin/env python
"""GPU diagnostics (run under gpurun): (1) which coordinate-arithmetic form bit-matches torch's CUDA ops,
(2) per-kernel timing at the 1080p shapes of SURVEY.md 8a against the eager torch chain the reference runs.
Writes gpurun_out/probe.json.  Uses the oracle as the checker only."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from b200vc import modules, ops  # noqa: E402
from gpu_util import gc_case, warp_case  # noqa: E402
from oracle import cai  # noqa: E402
from oracle import warp as o_warp  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
OUT = {}
flush_buf = torch.empty(256 * 1024 * 1024 // 4, device="cuda")


def timeit(fn, reps=10, flush=True):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if flush:
            flush_buf.add_(1.0)  # 256 MB write: evicts L2
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / reps


def section_arith():
    res = {}
    fns = {"lhbdc": o_warp.backwarp_lhbdc, "flex": o_warp.backwarp_flex, "ac1": o_warp.warp_ac1}
    for variant, fn in fns.items():
        img, flow = warp_case(7, 1, 3, 1088, 1920, amp=4.0)
        want = fn(img, flow)
        for arith in (0, 1, 2, 3):
            got = ops.backwarp(img, flow, variant, arith=arith)
            res[f"{variant}_arith{arith}"] = {"max_abs": (got - want).abs().max().item(),
                                              "bit_exact": (got == want).float().mean().item()}
    # upsample form
    g = torch.Generator().manual_seed(11)
    N, H, W = 1, 1088, 1920
    hh, ww, h4, w4 = 272, 480, 320, 512
    xb = torch.rand(N, 3, H, W, generator=g).cuda()
    xa = torch.rand(N, 3, H, W, generator=g).cuda()
    fh = (2.0 * torch.randn(N, 4, h4, w4, generator=g)).cuda()
    fab = (1.5 * torch.randn(N, 2, h4, w4, generator=g)).cuda()
    fba = (1.5 * torch.randn(N, 2, h4, w4, generator=g)).cuda()
    cb, ca = o_warp.lhbdc_flow_glue(fh, fab, fba, hh, ww)
    wantf = torch.cat([cb, ca], 1)
    want = torch.cat([o_warp.backwarp_lhbdc(xb, cb), o_warp.backwarp_lhbdc(xa, ca)], 1)
    for arith in (0, 1):
        got, flows = ops.warp2_lhbdc(xb, xa, fh, fab, fba, return_flows=True, arith=arith)
        res[f"warp2_arith{arith}"] = {"flow_max_abs": (flows - wantf).abs().max().item(),
                                      "flow_bit_exact": (flows == wantf).float().mean().item(),
                                      "img_max_abs": (got - want).abs().max().item(),
                                      "img_bit_exact": (got == want).float().mean().item()}
    OUT["arith"] = res
    for k, v in res.items():
        print(k, v)


def section_timing():
    res = {}
    peak = 6539.2

    def rec(name, ms, nbytes, ref_ms=None):
        res[name] = {"ms": ms, "GBps": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak, "torch_ms": ref_ms}
        print(f"{name:34s} {ms*1e3:9.1f} us  {nbytes/ms/1e6:8.1f} GB/s ({nbytes/ms/1e6/peak:5.1%})"
              + (f"   torch chain {ref_ms*1e3:9.1f} us  x{ref_ms/ms:.1f}" if ref_ms else ""))

    for N in (1, 4):
        img, flow = warp_case(1, N, 3, 1088, 1920, amp=4.0)
        rec(f"warp_lhbdc N={N}", timeit(lambda: ops.backwarp(img, flow, "lhbdc")), 32 * N * 1088 * 1920,
            timeit(lambda: o_warp.backwarp_lhbdc(img, flow)))
    def smooth_flow(n, h, w, amp=3.0):
        f = amp * torch.randn(n, 2, h // 16, w // 16)
        return torch.nn.functional.interpolate(f, size=(h, w), mode="bilinear", align_corners=False).cuda()

    for N in (1, 4):
        img, _ = warp_case(1, N, 3, 1088, 1920, amp=4.0)
        flow = smooth_flow(N, 1088, 1920)
        for variant in ("lhbdc", "flex", "ac1"):
            rec(f"warp_{variant} smooth N={N}", timeit(lambda: ops.backwarp(img, flow, variant)), 32 * N * 1088 * 1920)
    img, flow = warp_case(2, 1, 64, 544, 960, amp=4.0)
    rec("warp_ac1 C=64 544x960", timeit(lambda: ops.backwarp(img, flow, "ac1")), 130 * 4 * 544 * 960,
        timeit(lambda: o_warp.warp_ac1(img, flow)))
    g = torch.Generator().manual_seed(11)
    for N in (1, 4):
        xb = torch.rand(N, 3, 1088, 1920, generator=g).cuda()
        xa = torch.rand(N, 3, 1088, 1920, generator=g).cuda()
        fh = (2.0 * torch.randn(N, 4, 320, 512, generator=g)).cuda()
        fab = (1.5 * torch.randn(N, 2, 320, 512, generator=g)).cuda()
        fba = (1.5 * torch.randn(N, 2, 320, 512, generator=g)).cuda()

        def chain():
            cb, ca = o_warp.lhbdc_flow_glue(fh, fab, fba, 272, 480)
            return torch.cat([o_warp.backwarp_lhbdc(xb, cb), o_warp.backwarp_lhbdc(xa, ca)], 1)
        rec(f"warp2_lhbdc N={N}", timeit(lambda: ops.warp2_lhbdc(xb, xa, fh, fab, fba)), 50 * N * 1088 * 1920,
            timeit(chain))
        mask = torch.rand(N, 1, 1088, 1920, device="cuda")
        both = torch.rand(N, 6, 1088, 1920, device="cuda")
        x = torch.rand(N, 3, 1088, 1920, device="cuda")
        rec(f"blend_residual N={N}", timeit(lambda: ops.blend_residual("mask", mask, both[:, :3], both[:, 3:], x)),
            64 * N * 1088 * 1920, timeit(lambda: o_warp.blend_residual_lhbdc(mask, both[:, :3], both[:, 3:], x)))
    for (N, H, W) in ((1, 544, 960), (1, 272, 480), (1, 136, 240), (4, 544, 960), (1, 160, 256)):
        o = cai.GDN(128).cuda().eval()
        p = modules.GDN(128).cuda().eval()
        x = torch.randn(N, 128, H, W, device="cuda")
        params = modules.gdn_params(p)
        with torch.no_grad():
            t_ref = timeit(lambda: o(x))
        rec(f"gdn_fp32 N={N} {H}x{W}", timeit(lambda: ops.gdn(x, params, impl=1)), 1024 * N * H * W, t_ref)
        rec(f"gdn_auto N={N} {H}x{W}", timeit(lambda: ops.gdn(x, params, impl=0)), 1024 * N * H * W, t_ref)
    og = cai.GaussianConditional(None).cuda().eval()
    for N in (1, 4, 16):
        y, s, m = gc_case(3, N, 128, 68, 120)
        with torch.no_grad():
            t_ref = timeit(lambda: torch.log(og(y, s, means=m)[1]).sum())
        rec(f"gauss_cond full N={N}", timeit(lambda: ops.gauss_cond(y, s, m)), 20 * y.numel(), t_ref)
        rec(f"gauss_cond bits-only N={N}", timeit(lambda: ops.gauss_cond(y, s, m, want_lik=False)), 16 * y.numel(), t_ref)
    oe = cai.EntropyBottleneck(128).cuda().eval()
    pe = modules.EntropyBottleneck(128).cuda().eval()
    packed = modules.eb_packed(pe)
    for N in (1, 16):
        z = 3 * torch.randn(N, 128, 17, 30, device="cuda")
        with torch.no_grad():
            t_ref = timeit(lambda: torch.log(oe(z)[1]).sum())
        rec(f"entropy_bottleneck N={N}", timeit(lambda: ops.entropy_bottleneck(z, packed)), 12 * z.numel(), t_ref)
    # rANS coder at the 1080p residual-latent shape [1,128,68,120] (1 044 480 symbols)
    from b200vc import coding
    gc = modules.GaussianConditional(None).cuda().eval()
    gc.update_scale_table(modules.get_scale_table())
    y, s_, m_ = gc_case(3, 1, 128, 68, 120)
    idx = gc.build_indexes(s_)
    sym = gc.quantize(y, "symbols", m_)
    tab = modules.gc_tables(gc)
    for sl in (4096, 1024, 256):
        data = coding.rans_encode(sym, idx, tab, stream_len=sl)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            data = coding.rans_encode(sym, idx, tab, stream_len=sl)
        te = (time.perf_counter() - t0) / 5
        t0 = time.perf_counter()
        for _ in range(5):
            back = coding.rans_decode(data, idx, tab)
        torch.cuda.synchronize()
        td = (time.perf_counter() - t0) / 5
        ok = bool((back.view_as(sym) == sym).all().item())
        est = ops.gauss_cond(y, s_, m_, want_y_hat=False, want_lik=False)["bits"].item()
        print(f"rans stream_len={sl:5d}: encode {te*1e3:7.2f} ms  decode {td*1e3:7.2f} ms (wall, incl. D2H/H2D + container) "
              f"{len(data)} B = {8*len(data)/est:.4f} x estimated bits, round trip {ok}")
        res[f"rans_{sl}"] = {"encode_ms": te * 1e3, "decode_ms": td * 1e3, "bytes": len(data), "ratio_to_estimate": 8 * len(data) / est}
    OUT["timing"] = res


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.__version__)
    section_arith()
    section_timing()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(OUT, f, indent=1)
