"""ctypes binding of libb200vc.so (the C-ABI declared in include/b200vc.h).

This is the stub a maintainer of the reference would add on their side: the reference
has no FFI layer, so the binding is new, but every entry point stands in for a reference torch/CompressAI
call chain cited in the header.  The library is REQUIRED: there is no CPU / eager fallback.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200VC_LIB") or os.path.join(_HERE, "libb200vc.so")   # B200VC_LIB: A/B timing of two builds

WARP_LHBDC, WARP_FLEX, WARP_AC1 = 0, 1, 2
ARITH_NO_FMA, ARITH_TRUE_DIV = 1, 2
BLEND_MASK, BLEND_NORMW, BLEND_HALF = 0, 1, 2
EB_PARAMS_PER_CHANNEL = 59

_fp = c_void_p  # device pointers travel as integers

_SIGNATURES = {
    "b200vc_version": (c_int, []),
    "b200vc_last_error": (c_char_p, []),
    "b200vc_sm_count": (c_int, []),
    "b200vc_reduce_blocks": (c_int, [c_int64]),
    "b200vc_warp_f32": (c_int, [_fp, c_int64, _fp, _fp, _fp, _fp, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200vc_warp2_lhbdc_f32": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200vc_warp2_flex_f32": (c_int, [_fp, _fp, _fp, c_int64, _fp, c_int64, c_int, c_float, c_float, c_float, c_float, _fp, c_int, c_int, c_int, c_void_p]),
    "b200vc_warp2_half_sse_blocks": (c_int, [c_int, c_int]),
    "b200vc_warp2_half_sse_f32": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, c_int, c_int, c_int, c_int, c_void_p]),
    "b200vc_warp_sse_f32": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, c_int, c_int, c_int, c_int, c_void_p]),
    "b200vc_deform_conv2d_f32": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp] + [c_int] * 15 + [c_void_p]),
    "b200vc_round_checker_f32": (c_int, [_fp, _fp, _fp, c_int, c_int, c_int, c_int, c_void_p]),
    "b200vc_checker_mask_f32": (c_int, [_fp, c_int64, _fp, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200vc_spynet_pyramid_f32": (c_int, [_fp, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200vc_spynet_level_f32": (c_int, [_fp, c_int64, _fp, c_int64, _fp, _fp, _fp, _fp, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200vc_blend_residual_f32": (c_int, [c_int, _fp, _fp, c_int64, _fp, c_int64, _fp, _fp, _fp, _fp, c_int, _fp, _fp, c_int, c_int, c_int, c_void_p]),
    "b200vc_gdn_params_floats": (c_int64, [c_int]),
    "b200vc_gdn_prepare_f32": (c_int, [_fp, _fp, c_float, c_float, c_float, _fp, c_int, c_void_p]),
    "b200vc_gdn_f32": (c_int, [_fp, _fp, _fp, _fp, c_int, c_int, c_int64, c_int, c_int, c_void_p]),
    "b200vc_gauss_cond_f32": (c_int, [_fp, _fp, _fp, c_int64, _fp, _fp, _fp, _fp, _fp, _fp, c_int, c_float, c_float, _fp, c_int, _fp, _fp, c_int, c_int, c_int64, c_void_p]),
    "b200vc_eb_prepare_f32": (c_int, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), _fp, _fp, c_int, c_void_p]),
    "b200vc_entropy_bottleneck_f32": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, c_float, _fp, c_int, _fp, _fp, c_int, c_int, c_int64, c_void_p]),
    "b200vc_sum_partials_f64": (c_int, [_fp, c_int, c_int, _fp, c_void_p]),
    "b200vc_sse_u8_f32": (c_int, [_fp, _fp, _fp, c_int, _fp, _fp, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200vc_rans_scratch_words": (c_int, [c_int]),
    "b200vc_rans_encode": (c_int, [_fp, _fp, _fp, _fp, _fp, c_int, c_int64, c_int, _fp, _fp, c_void_p]),
    "b200vc_rans_compact": (c_int, [_fp, c_int, _fp, _fp, c_int, _fp, c_void_p]),
    "b200vc_rans_decode": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, c_int, c_int64, c_int, _fp, _fp, c_void_p]),
}

EXPORTS = tuple(_SIGNATURES)
_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built: no fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"b200vc: {LIB_PATH} is missing -- build it with `python video-compression_b200/build.py` "
            "(nvcc, sm_100a).  There is no CPU or eager-PyTorch fallback for the hot path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if hasattr(lib, "b200vc_debug_set_gdn_trace"):  # debug builds only (include/b200vc_debug.h)
        lib.b200vc_debug_set_gdn_trace.restype = None
        lib.b200vc_debug_set_gdn_trace.argtypes = [_fp]
    _lib = lib
    return lib


def last_error():
    msg = load().b200vc_last_error()
    return msg.decode() if msg else ""


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"b200vc.{what} failed (code {rc}): {last_error()}")
