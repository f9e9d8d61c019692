"""Torch-tensor wrappers over the b200vc C-ABI (one function per reference call chain).

Torch is plumbing here: it owns the device memory and the stream; all arithmetic happens in
libb200vc.so.  Every wrapper validates device / dtype / layout in Python (the reference's error
convention is Python exceptions) and raises ``RuntimeError`` on a non-zero return code.
Non-CUDA inputs raise: there is no CPU fallback.
"""
import math
import os

import torch

from . import _lib
from ._lib import (ARITH_NO_FMA, ARITH_TRUE_DIV, BLEND_HALF, BLEND_MASK, BLEND_NORMW, WARP_AC1, WARP_FLEX,
                   WARP_LHBDC)

_VARIANTS = {"lhbdc": WARP_LHBDC, "flex": WARP_FLEX, "ac1": WARP_AC1}
_BLENDS = {"mask": BLEND_MASK, "normw": BLEND_NORMW, "half": BLEND_HALF}
_launches = 0  # kernels launched through this module (bench.py reports it as gpu_launches)


def launch_count():
    return _launches


_prof = None  # name -> list of (start event, end event, algorithmic bytes) while profiling


def profile_begin():
    """Start recording a CUDA-event pair around every kernel launch made through this module (on the
    launching stream).  bench.py uses it inside its timed region to get per-kernel durations live."""
    global _prof
    _prof = {}


def profile_end():
    """Stop recording; returns {kernel: {"launches", "ms", "bytes"}} (synchronises the device)."""
    global _prof
    rec, _prof = _prof or {}, None
    torch.cuda.synchronize()
    out = {}
    for name, items in rec.items():
        ms = sum(s.elapsed_time(e) for s, e, _, _ in items)
        out[name] = {"launches": len(items), "ms": ms, "bytes": float(sum(b for _, _, b, _ in items))}
        tags = sorted({t for _, _, _, t in items if t is not None})
        if tags:
            out[name]["by_shape"] = {
                t: {"launches": sum(1 for i in items if i[3] == t),
                    "ms": sum(i[0].elapsed_time(i[1]) for i in items if i[3] == t),
                    "bytes": float(sum(i[2] for i in items if i[3] == t))} for t in tags}
    return out


def _run(name, nbytes, call, tag=None):
    """Launch one kernel through the C-ABI: count it, check the return code, optionally time it."""
    global _launches
    _launches += 1
    if _prof is None:
        rc = call()
    else:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = call()
        e.record()
        _prof.setdefault(name, []).append((s, e, nbytes, tag))
    _lib.check(rc, name)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (b200vc has no CPU fallback), got device {t.device}")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype}")


def _planes(t, name):
    """Return (tensor, data_ptr, batch_stride) for an NCHW tensor whose (C,H,W) block is dense; a channel
    slice of a wider tensor qualifies.  Anything else is made contiguous."""
    _need_cuda_f32(t, name)
    if t.dim() != 4:
        raise RuntimeError(f"{name}: expected a 4-D NCHW tensor, got shape {tuple(t.shape)}")
    N, C, H, W = t.shape
    st = t.stride()
    dense = st[3] == 1 and st[2] == W and st[1] == H * W and (N == 1 or st[0] >= C * H * W)
    if not dense:
        t = t.contiguous()
        st = t.stride()
    return t, t.data_ptr(), (st[0] if N > 1 else C * H * W)


def _contig(t, name):
    _need_cuda_f32(t, name)
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------------------- warp
_tables = {}


def grid_tables(variant, H, W, device):
    """Base sampling grid, built exactly as the reference does (torch.linspace on the CPU), kept as two
    1-D tables instead of a [1,2,H,W] tensor (LHBDC/model/m.py:113-118; ICIP2024/src/model/m.py:264-266)."""
    key = (variant, H, W, str(device))
    tab = _tables.get(key)
    if tab is None:
        if variant == "lhbdc":
            tx = torch.linspace(-1.0 + (1.0 / W), 1.0 - (1.0 / W), W)
            ty = torch.linspace(-1.0 + (1.0 / H), 1.0 - (1.0 / H), H)
        elif variant == "ac1":
            tx = torch.linspace(-1.0, 1.0, W)
            ty = torch.linspace(-1.0, 1.0, H)
        else:
            raise ValueError(variant)
        tab = (tx.to(device), ty.to(device))
        _tables[key] = tab
    return tab


def backwarp(img, flow, variant="lhbdc", out=None, arith=0):
    """Bilinear backward warp.  ``variant``: 'lhbdc' (Model.backwarp / flow.backwarp), 'flex'
    (BidirFlowRef.backwarp) or 'ac1' (FlowGuidedB.warp / DMC.warp).  ``out`` may be a channel slice of a
    concat buffer."""
    if variant not in _VARIANTS:
        raise ValueError(f"backwarp: unknown variant {variant!r}")
    img, ip, ibs = _planes(img, "backwarp(img)")
    flow = _contig(flow, "backwarp(flow)")
    N, C, H, W = img.shape
    if tuple(flow.shape) != (N, 2, H, W):
        raise RuntimeError(f"backwarp: flow shape {tuple(flow.shape)} does not match image {tuple(img.shape)}")
    if out is None:
        out = torch.empty_like(img, memory_format=torch.contiguous_format)
    if img.numel() == 0:  # empty batch: nothing to launch (the reference's grid_sample returns an empty tensor too)
        return out
    o, op, obs = _planes(out, "backwarp(out)")
    if o is not out or tuple(out.shape) != (N, C, H, W):
        raise RuntimeError("backwarp: `out` must be a dense [N,C,H,W] block (channel slices are fine)")
    if variant == "flex":
        tx = ty = None
    else:
        tx, ty = grid_tables(variant, H, W, img.device)
    lib = _lib.load()
    _run("warp_f32", (2 * C + 2) * 4 * N * H * W, lambda: lib.b200vc_warp_f32(
        ip, ibs, flow.data_ptr(), tx.data_ptr() if tx is not None else None,
        ty.data_ptr() if ty is not None else None, op, obs, N, C, H, W, _VARIANTS[variant], arith, _stream()),
         tag=f"{N}x{C}x{H}x{W}")
    return out


def spynet_pyramid(frame, preprocess=True, max_poolings=5):
    """LHBDC/model/flow.py:80-88 in one launch: Preprocess (channel flip + ImageNet statistics) and the avg-pool
    pyramid.  Returns the levels coarsest first, as the reference's ``tenFirst`` list ends up."""
    import ctypes
    frame, fp, fbs = _planes(frame, "spynet_pyramid(frame)")
    N, C, H, W = frame.shape
    if C != 3:
        raise RuntimeError(f"spynet_pyramid: expected 3 channels, got {C}")
    sizes = [(H, W)]
    for _ in range(max_poolings):  # flow.py:84: pool while a side is larger than 32
        h, w = sizes[-1]
        if h > 32 or w > 32:
            sizes.append((h // 2, w // 2))
    n_levels = len(sizes) - 1
    if not preprocess and n_levels == 0:
        return [frame]
    levels = [torch.empty((N, 3, h, w), device=frame.device, dtype=frame.dtype) if (l > 0 or preprocess) else frame
              for l, (h, w) in enumerate(sizes)]
    ptrs = (ctypes.c_void_p * (n_levels + 1))(*[(t.data_ptr() if (l > 0 or preprocess) else None)
                                                for l, t in enumerate(levels)])
    lib = _lib.load()
    total = sum(h * w for h, w in sizes[1:]) + (H * W if preprocess else 0)
    _run("spynet_pyramid_f32", 4 * 3 * N * (H * W + total), lambda: lib.b200vc_spynet_pyramid_f32(
        fp, fbs, ptrs, N, H, W, n_levels, int(bool(preprocess)), _stream()), tag=f"{N}x3x{H}x{W}")
    return levels[::-1]


def spynet_level(first, second, flow_prev=None):
    """One SPyNet level's 8-channel conv input (flow.py:93-98): cat(first, backwarp(second, up), up) with
    up = 2 * bilinear x2 (align_corners=True) of ``flow_prev``, replicate-padded to the level's size.
    ``flow_prev=None`` is the all-zero initial flow."""
    first, p1, bs1 = _planes(first, "spynet_level(first)")
    second, p2, bs2 = _planes(second, "spynet_level(second)")
    N, C, H, W = first.shape
    if C != 3 or tuple(second.shape) != (N, 3, H, W):
        raise RuntimeError(f"spynet_level: expected two [N,3,H,W] images, got {tuple(first.shape)} and {tuple(second.shape)}")
    hp = wp = 0
    if flow_prev is not None:
        flow_prev = _contig(flow_prev, "spynet_level(flow_prev)")
        if flow_prev.dim() != 4 or flow_prev.shape[0] != N or flow_prev.shape[1] != 2:
            raise RuntimeError(f"spynet_level: flow_prev must be [N,2,h,w], got {tuple(flow_prev.shape)}")
        hp, wp = flow_prev.shape[2:]
        if H not in (2 * hp, 2 * hp + 1) or W not in (2 * wp, 2 * wp + 1):
            raise RuntimeError(f"spynet_level: a {hp}x{wp} flow does not upsample to {H}x{W}")
    tx, ty = grid_tables("lhbdc", H, W, first.device)
    feat = torch.empty((N, 8, H, W), device=first.device, dtype=first.dtype)
    lib = _lib.load()
    _run("spynet_level_f32", 4 * N * (14 * H * W + 2 * hp * wp), lambda: lib.b200vc_spynet_level_f32(
        p1, bs1, p2, bs2, flow_prev.data_ptr() if flow_prev is not None else None, tx.data_ptr(), ty.data_ptr(),
        feat.data_ptr(), N, H, W, hp, wp, _stream()), tag=f"{N}x8x{H}x{W}")
    return feat


def warp2_lhbdc(x_before, x_after, flow_hat, flow_ab, flow_ba, return_flows=False, arith=0):
    """Fused LHBDC/model/m.py:55-63: flow glue + two warps + concat -> [N,6,H,W] (and optionally the two
    full-resolution flows [N,4,H,W])."""
    xb = _contig(x_before, "warp2(x_before)")
    xa = _contig(x_after, "warp2(x_after)")
    fh = _contig(flow_hat, "warp2(flow_hat)")
    fab = _contig(flow_ab, "warp2(flow_ab)")
    fba = _contig(flow_ba, "warp2(flow_ba)")
    N, C, H, W = xb.shape
    if C != 3 or xa.shape != xb.shape:
        raise RuntimeError("warp2_lhbdc: references must both be [N,3,H,W]")
    h4, w4 = fh.shape[2], fh.shape[3]
    if tuple(fh.shape) != (N, 4, h4, w4) or tuple(fab.shape) != (N, 2, h4, w4) or tuple(fba.shape) != (N, 2, h4, w4):
        raise RuntimeError("warp2_lhbdc: flow_hat must be [N,4,h,w] and the priors [N,2,h,w]")
    tx, ty = grid_tables("lhbdc", H, W, xb.device)
    out = torch.empty((N, 6, H, W), device=xb.device, dtype=torch.float32)
    flows = torch.empty((N, 4, H, W), device=xb.device, dtype=torch.float32) if return_flows else None
    if N == 0:
        return (out, flows) if return_flows else out
    lib = _lib.load()
    # algorithmic bytes: 2 references in, 6-channel concat out, quarter-res flows (4+2+2 channels over HW/16)
    nbytes = N * H * W * (12 * 4 + 8 * 4 / 16 + (16 if return_flows else 0))
    _run("warp2_lhbdc_f32", nbytes, lambda: lib.b200vc_warp2_lhbdc_f32(
        xb.data_ptr(), xa.data_ptr(), fh.data_ptr(), fab.data_ptr(), fba.data_ptr(), tx.data_ptr(), ty.data_ptr(),
        out.data_ptr(), flows.data_ptr() if flows is not None else None, N, H, W, h4, w4, arith, _stream()))
    return (out, flows) if return_flows else out


def warp2_flex(x0, x1, fa, fb, mode, t=0.5):
    """Fused Flex-Rate motion compensation (b_model.py:34-45 and :58-66) -> the 16-channel buffer
    cat(ft0, ft1, x0, x1, backwarp(x0, ft0), backwarp(x1, ft1)).
    mode 'linear': fa = Flow_0_1, fb = Flow_1_0 ([N,2,H,W] each; channel slices of the predictor output are fine),
    ft0 = -(1-t)t*fa + t*t*fb, ft1 = (1-t)(1-t)*fa - t(1-t)*fb.
    mode 'refine': fa = cat(mv_before, mv_after) (e.g. channels 0:4 of the previous buffer), fb = flow_hat[:, 0:4]."""
    x0 = _contig(x0, "warp2_flex(x0)")
    x1 = _contig(x1, "warp2_flex(x1)")
    N, C, H, W = x0.shape
    cf = 2 if mode == "linear" else 4
    if mode not in ("linear", "refine"):
        raise ValueError(f"warp2_flex: unknown mode {mode!r}")
    if C != 3 or x1.shape != x0.shape:
        raise RuntimeError("warp2_flex: references must both be [N,3,H,W]")
    fa, pa, abs_ = _planes(fa, "warp2_flex(fa)")
    fb, pb, bbs = _planes(fb, "warp2_flex(fb)")
    if tuple(fa.shape) != (N, cf, H, W) or tuple(fb.shape) != (N, cf, H, W):
        raise RuntimeError(f"warp2_flex: flow inputs must be [N,{cf},H,W] for mode {mode!r}")
    out = torch.empty((N, 16, H, W), device=x0.device, dtype=torch.float32)
    if N == 0:
        return out
    # the reference multiplies fp32 tensors by python floats: the scalar is cast to fp32, each product rounded once
    a0, b0, a1, b1 = -(1 - t) * t, t * t, (1 - t) * (1 - t), -(t * (1 - t))
    lib = _lib.load()
    _run("warp2_flex_f32", N * H * W * 4 * (6 + 2 * cf + 16), lambda: lib.b200vc_warp2_flex_f32(
        x0.data_ptr(), x1.data_ptr(), pa, abs_, pb, bbs, 0 if mode == "linear" else 1, a0, b0, a1, b1, out.data_ptr(),
        N, H, W, _stream()), tag=f"{N}x16x{H}x{W}")
    return out


def warp2_half_sse(x1, x2, flow1, flow2, x_cur, variant="ac1", want_pred=False):
    """Search form of ICIP2024/src/opt_helpers.py:23-51: warp both references, 0.5/0.5 blend, clamp, squared error
    against ``x_cur`` -- one kernel.  Returns (sse[N] float64, pred or None); MSE = sse / (3*H*W)."""
    if variant not in _VARIANTS:
        raise ValueError(f"warp2_half_sse: unknown variant {variant!r}")
    x1, x2, xc = (_contig(t, "warp2_half_sse(img)") for t in (x1, x2, x_cur))
    f1, f2 = _contig(flow1, "warp2_half_sse(flow1)"), _contig(flow2, "warp2_half_sse(flow2)")
    N, C, H, W = x1.shape
    if C != 3 or x2.shape != x1.shape or xc.shape != x1.shape or tuple(f1.shape) != (N, 2, H, W) or f2.shape != f1.shape:
        raise RuntimeError("warp2_half_sse: expected [N,3,H,W] images and [N,2,H,W] flows")
    tx, ty = (None, None) if variant == "flex" else grid_tables(variant, H, W, x1.device)
    lib = _lib.load()
    nb = lib.b200vc_warp2_half_sse_blocks(H, W)
    part = torch.empty(N * nb, device=x1.device, dtype=torch.float64)
    pred = torch.empty_like(x1) if want_pred else None
    p = lambda t: t.data_ptr() if t is not None else None
    tot, cnt = _finish(x1.device, N)
    _run("warp2_half_sse_f32", (13 + 3 * int(want_pred)) * 4 * N * H * W, lambda: lib.b200vc_warp2_half_sse_f32(
        x1.data_ptr(), x2.data_ptr(), f1.data_ptr(), f2.data_ptr(), xc.data_ptr(), p(tx), p(ty), p(pred),
        part.data_ptr(), p(tot), p(cnt), N, H, W, _VARIANTS[variant], _stream()))
    return (tot if tot is not None else sum_partials(part, nb, N)), pred


def warp_sse(img, flow, x_cur, variant="ac1", want_pred=False):
    """Single-reference search form (OJSP2025/video_model.py:643-645): ``x_hat = warp(img, flow)`` and the squared
    error against ``x_cur`` in one kernel.  Returns (sse[N] float64, x_hat or None); MSE = sse / (3*H*W)."""
    if variant not in _VARIANTS:
        raise ValueError(f"warp_sse: unknown variant {variant!r}")
    img, xc = _contig(img, "warp_sse(img)"), _contig(x_cur, "warp_sse(x_cur)")
    flow = _contig(flow, "warp_sse(flow)")
    N, C, H, W = img.shape
    if C != 3 or xc.shape != img.shape or tuple(flow.shape) != (N, 2, H, W):
        raise RuntimeError("warp_sse: expected [N,3,H,W] images and an [N,2,H,W] flow")
    tx, ty = (None, None) if variant == "flex" else grid_tables(variant, H, W, img.device)
    lib = _lib.load()
    nb = lib.b200vc_warp2_half_sse_blocks(H, W)
    part = torch.empty(N * nb, device=img.device, dtype=torch.float64)
    pred = torch.empty_like(img) if want_pred else None
    p = lambda t: t.data_ptr() if t is not None else None
    tot, cnt = _finish(img.device, N)
    _run("warp_sse_f32", (8 + 3 * int(want_pred)) * 4 * N * H * W, lambda: lib.b200vc_warp_sse_f32(
        img.data_ptr(), flow.data_ptr(), xc.data_ptr(), p(tx), p(ty), p(pred), part.data_ptr(), p(tot), p(cnt), N, H, W,
        _VARIANTS[variant], _stream()), tag=f"{N}x3x{H}x{W}")
    return (tot if tot is not None else sum_partials(part, nb, N)), pred


# --------------------------------------------------------------------- checkerboard context glue
def round_checker(y, want_half=True):
    """``ste_round(y)`` and its anchor-zeroed copy (compression_bottlenecks.py:238-242) for the whole latent at once.
    Returns (y_hat, y_half or None)."""
    y = _contig(y, "round_checker(y)")
    if y.dim() != 4:
        raise RuntimeError(f"round_checker: expected [N,C,H,W], got {tuple(y.shape)}")
    N, C, H, W = y.shape
    y_hat = torch.empty_like(y)
    y_half = torch.empty_like(y) if want_half else None
    if y.numel() == 0:
        return y_hat, y_half
    lib = _lib.load()
    _run("round_checker_f32", 4 * y.numel() * (2 + int(want_half)), lambda: lib.b200vc_round_checker_f32(
        y.data_ptr(), y_hat.data_ptr(), y_half.data_ptr() if want_half else None, N, C, H, W, _stream()))
    return y_hat, y_half


def checker_mask(src, out=None, zero_parity=1):
    """``x[:, :, 0::2, 1::2] = 0; x[:, :, 1::2, 0::2] = 0`` (zero_parity=1; compression_bottlenecks.py:245-246) as one
    pass; ``out`` may be a channel slice of a concat buffer, or ``src`` itself."""
    src, sp, sbs = _planes(src, "checker_mask(src)")
    N, C, H, W = src.shape
    if out is None:
        out = torch.empty_like(src)
    o, op, obs = _planes(out, "checker_mask(out)")
    if o is not out or tuple(out.shape) != (N, C, H, W):
        raise RuntimeError("checker_mask: `out` must be a dense [N,C,H,W] block (channel slices are fine)")
    if src.numel() == 0:
        return out
    lib = _lib.load()
    _run("checker_mask_f32", 8 * src.numel(), lambda: lib.b200vc_checker_mask_f32(
        sp, sbs, op, obs, N, C, H, W, int(zero_parity), _stream()))
    return out


# ------------------------------------------------------------------------- deformable convolution
def deform_conv2d(input, offset, weight, bias=None, stride=(1, 1), padding=(0, 0), dilation=(1, 1), mask=None,
                  use_workspace=True):
    """``torchvision.ops.deform_conv2d`` (same signature and layouts; ICIP2023/src/model/m.py:29-34,
    ICIP2024/src/model/helpers.py:40,57) without the im2col matrix.  ``use_workspace=False`` forces the NCHW gather
    kernel (no scratch copy of the input)."""
    pair = lambda v: (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))
    (sh, sw), (ph, pw), (dh, dw) = pair(stride), pair(padding), pair(dilation)
    x, off, w = _contig(input, "deform_conv2d(input)"), _contig(offset, "deform_conv2d(offset)"), \
        _contig(weight, "deform_conv2d(weight)")
    if x.dim() != 4 or off.dim() != 4 or w.dim() != 4:
        raise RuntimeError("deform_conv2d: input, offset and weight must be 4-D")
    N, Cin, H, W = x.shape
    Cout, cin_g, kh, kw = w.shape
    if cin_g == 0 or Cin % cin_g != 0:
        raise RuntimeError(f"deform_conv2d: weight expects {cin_g} channels per group, input has {Cin}")
    groups = Cin // cin_g
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    K = kh * kw
    if off.shape[0] != N or off.shape[1] % (2 * K) != 0 or tuple(off.shape[2:]) != (Ho, Wo):
        raise RuntimeError(f"deform_conv2d: offset shape {tuple(off.shape)} does not match [N, 2*og*{K}, {Ho}, {Wo}]")
    og = off.shape[1] // (2 * K)
    if og == 0 or Cin % og != 0:
        raise RuntimeError(f"deform_conv2d: {og} offset groups do not divide {Cin} input channels")
    m = None
    if mask is not None:
        m = _contig(mask, "deform_conv2d(mask)")
        if tuple(m.shape) != (N, og * K, Ho, Wo):
            raise RuntimeError(f"deform_conv2d: mask shape {tuple(m.shape)} does not match [N, {og * K}, {Ho}, {Wo}]")
    b = None
    if bias is not None:
        b = _contig(bias, "deform_conv2d(bias)")
        if tuple(b.shape) != (Cout,):
            raise RuntimeError("deform_conv2d: bias must be [Cout]")
    out = torch.empty((N, Cout, Ho, Wo), device=x.device, dtype=x.dtype)
    if out.numel() == 0:
        return out
    lib = _lib.load()
    p = lambda t: t.data_ptr() if t is not None else None
    nbytes = 4 * (x.numel() + off.numel() + (m.numel() if m is not None else 0) + out.numel())
    cout_g = Cout // groups
    fast = (use_workspace and cin_g in (4, 8, 12, 16) and cout_g in (4, 6, 8, 12, 16) and (Cin // og) % cin_g == 0
            and K <= 9)
    ws = torch.empty(x.numel(), device=x.device, dtype=x.dtype) if fast else None
    _run("deform_conv2d_f32", nbytes, lambda: lib.b200vc_deform_conv2d_f32(
        x.data_ptr(), off.data_ptr(), p(m), w.data_ptr(), p(b), out.data_ptr(), p(ws), N, Cin, H, W, Cout, kh, kw, sh, sw,
        ph, pw, dh, dw, groups, og, _stream()), tag=f"{N}x{Cin}->{Cout}x{H}x{W}g{groups}")
    return out


# ------------------------------------------------------------------------------- blend / residual
def reduce_blocks(elems_per_sample):
    return _lib.load().b200vc_reduce_blocks(int(elems_per_sample))


_FUSED_FINISH = os.environ.get("B200VC_FUSED_FINISH", "1") != "0"
_counters = {}


def _finish(dev, n_out):
    """(totals[n_out] float64, counters int32) for the kernels' fused last-CTA-done reduction, or (None, None) when
    switched off (B200VC_FUSED_FINISH=0: the two-launch form through ``sum_partials``).  The ticket counters are one
    zeroed buffer per (device, stream): launches on a stream are ordered and every kernel leaves them zero."""
    if not _FUSED_FINISH:
        return None, None
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    buf = _counters.get(key)
    if buf is None or buf.numel() < n_out:
        buf = torch.zeros(max(1024, n_out), device=dev, dtype=torch.int32)
        _counters[key] = buf
    return torch.empty(n_out, device=dev, dtype=torch.float64), buf


def sum_partials(partials, n_per, n_out):
    out = torch.empty(n_out, device=partials.device, dtype=torch.float64)
    lib = _lib.load()
    _run("sum_partials_f64", 8 * (n_per + 1) * n_out, lambda: lib.b200vc_sum_partials_f64(
        partials.data_ptr(), n_per, n_out, out.data_ptr(), _stream()))
    return out


def blend_residual(mode, mask, a, b, x_cur, want_pred=True, want_res=True, want_sse=False):
    """mode 'mask' (LHBDC m.py:63-67), 'normw' (Flex b_model.py:68-73; mask = raw 2-ch logits) or 'half'
    (ICIP2024 opt_helpers.py:35-45).  Returns (pred, res, sse[N] float64) with None for parts not requested."""
    if mode not in _BLENDS:
        raise ValueError(f"blend_residual: unknown mode {mode!r}")
    a, ap, abs_ = _planes(a, "blend(a)")
    b, bp, bbs = _planes(b, "blend(b)")
    x = _contig(x_cur, "blend(x_cur)")
    N, C, H, W = x.shape
    if C != 3 or tuple(a.shape) != (N, 3, H, W) or tuple(b.shape) != (N, 3, H, W):
        raise RuntimeError("blend_residual: a, b, x_cur must all be [N,3,H,W]")
    mp = None
    if mode != "half":
        mask = _contig(mask, "blend(mask)")
        want_c = 2 if mode == "normw" else 1
        if tuple(mask.shape) != (N, want_c, H, W):
            raise RuntimeError(f"blend_residual: mask must be [N,{want_c},H,W] for mode {mode!r}")
        mp = mask.data_ptr()
    pred = torch.empty_like(x) if want_pred else None
    res = torch.empty_like(x) if want_res else None
    if N == 0:
        return pred, res, (torch.zeros(0, device=x.device, dtype=torch.float64) if want_sse else None)
    nb = reduce_blocks(H * W)
    part = torch.empty(N * nb, device=x.device, dtype=torch.float64) if want_sse else None
    lib = _lib.load()
    planes = 9 + (0 if mode == "half" else (2 if mode == "normw" else 1)) + 3 * int(want_pred) + 3 * int(want_res)
    tot, cnt = _finish(x.device, N) if want_sse else (None, None)
    p = lambda t: t.data_ptr() if t is not None else None
    _run("blend_residual_f32", planes * 4 * N * H * W, lambda: lib.b200vc_blend_residual_f32(
        _BLENDS[mode], mp, ap, abs_, bp, bbs, x.data_ptr(), pred.data_ptr() if want_pred else None,
        res.data_ptr() if want_res else None, part.data_ptr() if want_sse else None, nb, p(tot), p(cnt), N, H, W,
        _stream()))
    sse = (tot if tot is not None else sum_partials(part, nb, N)) if want_sse else None
    return pred, res, sse


def sse_u8(a, b, h, w):
    """Sum over the crop [:h,:w] of (uint8(a) - uint8(b))^2 with float_to_uint8 = round(clip(x,0,1)*255)
    (LHBDC/test/testing.py:176-182); returns a float64 device tensor [N] of per-sample sums (exact integers)."""
    a = _contig(a, "sse_u8(a)")
    b = _contig(b, "sse_u8(b)")
    if a.shape != b.shape or a.dim() != 4:
        raise RuntimeError("sse_u8: a and b must be NCHW tensors of the same shape")
    N, C, H, W = a.shape
    nb = reduce_blocks(C * h * w)
    part = torch.empty(N * nb, device=a.device, dtype=torch.float64)
    lib = _lib.load()
    tot, cnt = _finish(a.device, N)
    p = lambda t: t.data_ptr() if t is not None else None
    _run("sse_u8_f32", 8 * N * C * h * w, lambda: lib.b200vc_sse_u8_f32(
        a.data_ptr(), b.data_ptr(), part.data_ptr(), nb, p(tot), p(cnt), N, C, H, W, h, w, _stream()))
    return tot if tot is not None else sum_partials(part, nb, N)


# --------------------------------------------------------------------------------------- GDN
def gdn_prepare(beta, gamma, beta_bound, gamma_bound, pedestal):
    """NonNegativeParametrizer + operand images, once per weight version.  Returns the params tensor."""
    _need_cuda_f32(beta, "gdn_prepare(beta)")
    _need_cuda_f32(gamma, "gdn_prepare(gamma)")
    C = beta.numel()
    if tuple(gamma.shape) != (C, C):
        raise RuntimeError(f"gdn_prepare: gamma must be [{C},{C}]")
    lib = _lib.load()
    params = torch.empty(lib.b200vc_gdn_params_floats(C), device=beta.device, dtype=torch.float32)
    b, g = beta.detach().contiguous(), gamma.detach().contiguous()
    _run("gdn_prepare_f32", 4 * (C + 5 * C * C), lambda: lib.b200vc_gdn_prepare_f32(
        b.data_ptr(), g.data_ptr(), float(beta_bound), float(gamma_bound), float(pedestal), params.data_ptr(), C,
        _stream()))
    return params


_GDN_IMPL = int(os.environ.get("B200VC_GDN_IMPL", "0"))  # 0 auto (tcgen05 when C == 128), 1 exact fp32, 2 tcgen05
_GDN_INPLACE_ADD = os.environ.get("B200VC_GDN_INPLACE_ADD", "1") != "0"


def gdn(x, params, inverse=False, addend=None, impl=0, inplace=None):
    """out = x * rsqrt(beta + gamma @ x^2) (IGDN: sqrt) [+ addend]; x is NCHW.  With ``addend`` the sum is
    written into the addend's own storage (it is consumed) and that tensor is returned, unless ``inplace=False``
    (a fresh output tensor; required when the addend must survive, e.g. when it is the block's own input).
    impl 0 = auto: the tcgen05 3xTF32 kernel for C == 128 (~1e-6 relative), else the exact-fp32 CUDA-core
    kernel; B200VC_GDN_IMPL=1 forces the exact kernel everywhere (bit-parity experiments)."""
    if impl == 0:
        impl = _GDN_IMPL
    x = _contig(x, "gdn(x)")
    N, C, H, W = x.shape
    if addend is not None:
        addend = _contig(addend, "gdn(addend)")
        if addend.shape != x.shape:
            raise RuntimeError("gdn: addend shape mismatch")
    # residual form: accumulate into the addend's storage (the reference's `out += identity`); the tcgen05
    # kernel then needs no addend traffic inside the SM -- the result tile leaves through a TMA reduce-add
    if inplace is None:
        inplace = _GDN_INPLACE_ADD
    out = addend if (addend is not None and inplace) else torch.empty_like(x)
    if x.numel() == 0:
        return out
    lib = _lib.load()
    nbytes = (2 + (addend is not None)) * C * 4 * N * H * W
    _run("gdn_f32", nbytes, lambda: lib.b200vc_gdn_f32(
        x.data_ptr(), params.data_ptr(), addend.data_ptr() if addend is not None else None, out.data_ptr(), N, C,
        H * W, 1 if inverse else 0, impl, _stream()), tag=f"{N}x{C}x{H}x{W}")
    return out


# ----------------------------------------------------------------------------- entropy models
def gauss_cond(y, scales, means, scale_bound=0.11, lik_bound=1e-9, inv_gain=None, want_y_hat=True,
               want_lik=True, want_bits=True, want_symbols=False, scale_table=None):
    """GaussianConditional.forward (eval) [+ quantize('symbols') + build_indexes] + bit sums in one pass.
    Returns dict(y_hat, lik, bits[N] float64 (= sum -log2 lik), symbols, indexes)."""
    y = _contig(y, "gauss_cond(y)")
    N, C, H, W = y.shape
    scales, sp, sbs = _planes(scales, "gauss_cond(scales)")
    means, mp, mbs = _planes(means, "gauss_cond(means)")
    if tuple(scales.shape) != (N, C, H, W) or tuple(means.shape) != (N, C, H, W):
        raise RuntimeError("gauss_cond: scales / means shape mismatch")
    if sbs != mbs:
        means = means.contiguous()
        scales = scales.contiguous()
        sp, mp, sbs = scales.data_ptr(), means.data_ptr(), C * H * W
    dev = y.device
    y_hat = torch.empty_like(y) if want_y_hat else None
    lik = torch.empty_like(y) if want_lik else None
    sym = torch.empty(y.shape, device=dev, dtype=torch.int32) if want_symbols else None
    idx = torch.empty(y.shape, device=dev, dtype=torch.int32) if want_symbols else None
    if want_symbols:
        if scale_table is None or scale_table.numel() < 2:
            raise RuntimeError("gauss_cond: symbols/indexes need the scale table (call update() first)")
        scale_table = _contig(scale_table, "gauss_cond(scale_table)")
    if inv_gain is not None:
        inv_gain = _contig(inv_gain.reshape(-1), "gauss_cond(inv_gain)")
        if inv_gain.numel() != C:
            raise RuntimeError("gauss_cond: inv_gain must have C entries")
    if y.numel() == 0:
        return {"y_hat": y_hat, "lik": lik, "bits": torch.zeros(N, device=dev, dtype=torch.float64) if want_bits else None,
                "symbols": sym, "indexes": idx}
    # one partial slot per 4096 elements: the persistent kernel needs no slot-per-CTA parallelism, and four units
    # per lane and slot amortise the slot's log2 and warp reduction
    nb = reduce_blocks((C * H * W + 3) // 4)
    part = torch.empty(N * nb, device=dev, dtype=torch.float64) if want_bits else None
    p = lambda t: t.data_ptr() if t is not None else None
    lib = _lib.load()
    words = 3 + int(want_y_hat) + int(want_lik) + 2 * int(want_symbols)
    tot, cnt = _finish(dev, N) if want_bits else (None, None)
    _run("gauss_cond_f32", words * 4 * y.numel(), lambda: lib.b200vc_gauss_cond_f32(
        y.data_ptr(), sp, mp, sbs, p(inv_gain), p(y_hat), p(lik), p(sym), p(idx),
        p(scale_table) if want_symbols else None, scale_table.numel() if want_symbols else 0, float(scale_bound),
        float(lik_bound), p(part), nb, p(tot), p(cnt), N, C, H * W, _stream()), tag=f"{N}x{C}x{H}x{W}")
    bits = (tot if tot is not None else sum_partials(part, nb, N)) if want_bits else None
    return {"y_hat": y_hat, "lik": lik, "bits": bits, "symbols": sym, "indexes": idx}


def eb_prepare(matrices, biases, factors, quantiles):
    """Pack softplus(_matrix*), _bias*, tanh(_factor*), median per channel -> [C,59]."""
    import ctypes
    C = quantiles.shape[0]
    keep = [_contig(t.detach(), "eb_prepare") for t in list(matrices) + list(biases) + list(factors)]
    q = _contig(quantiles.detach(), "eb_prepare(quantiles)")
    arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    packed = torch.empty((C, _lib.EB_PARAMS_PER_CHANNEL), device=q.device, dtype=torch.float32)
    lib = _lib.load()
    a_m, a_b, a_f = arr(keep[0:5]), arr(keep[5:10]), arr(keep[10:14])
    _run("eb_prepare_f32", 4 * C * 2 * _lib.EB_PARAMS_PER_CHANNEL, lambda: lib.b200vc_eb_prepare_f32(
        a_m, a_b, a_f, q.data_ptr(), packed.data_ptr(), C, _stream()))
    return packed


def entropy_bottleneck(z, packed, lik_bound=1e-9, gain=None, inv_gain=None, want_z_hat=True, want_lik=True,
                       want_bits=True, want_symbols=False):
    """EntropyBottleneck.forward (eval) + bit sums.  Returns dict(z_hat, lik, bits[N], symbols)."""
    z = _contig(z, "entropy_bottleneck(z)")
    N, C, H, W = z.shape
    if tuple(packed.shape) != (C, _lib.EB_PARAMS_PER_CHANNEL):
        raise RuntimeError("entropy_bottleneck: packed parameter shape mismatch")
    dev = z.device
    z_hat = torch.empty_like(z) if want_z_hat else None
    lik = torch.empty_like(z) if want_lik else None
    sym = torch.empty(z.shape, device=dev, dtype=torch.int32) if want_symbols else None
    for g, nm in ((gain, "gain"), (inv_gain, "inv_gain")):
        if g is not None and g.numel() != C:
            raise RuntimeError(f"entropy_bottleneck: {nm} must have C entries")
    gain = _contig(gain.reshape(-1), "eb(gain)") if gain is not None else None
    inv_gain = _contig(inv_gain.reshape(-1), "eb(inv_gain)") if inv_gain is not None else None
    if z.numel() == 0:
        return {"z_hat": z_hat, "lik": lik, "bits": torch.zeros(N, device=dev, dtype=torch.float64) if want_bits else None,
                "symbols": sym}
    nb = reduce_blocks(4 * C * H * W)  # scalar kernel (heavy per-element math): one element per thread
    part = torch.empty(N * nb, device=dev, dtype=torch.float64) if want_bits else None
    p = lambda t: t.data_ptr() if t is not None else None
    lib = _lib.load()
    words = 1 + int(want_z_hat) + int(want_lik) + int(want_symbols)
    tot, cnt = _finish(dev, N) if want_bits else (None, None)
    _run("entropy_bottleneck_f32", words * 4 * z.numel(), lambda: lib.b200vc_entropy_bottleneck_f32(
        z.data_ptr(), packed.data_ptr(), p(gain), p(inv_gain), p(z_hat), p(lik), p(sym), float(lik_bound), p(part),
        nb, p(tot), p(cnt), N, C, H * W, _stream()))
    bits = (tot if tot is not None else sum_partials(part, nb, N)) if want_bits else None
    return {"z_hat": z_hat, "lik": lik, "bits": bits, "symbols": sym}
