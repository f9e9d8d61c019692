"""``patch(model)``: make an *existing* model -- built from the reference's own classes on top of a real
CompressAI install (or from the test oracle) -- run the b200vc kernels, leaving its module tree, attribute
names and checkpoint keys untouched.  Dispatch is by class name, so nothing from the reference
or from CompressAI is imported here.
"""
import sys
import types

from . import icip, lhbdc, modules, ops

_WARP_BY_CLASS = {
    "Model": ("backwarp", "lhbdc"),            # LHBDC/model/m.py:111
    "BidirFlowRef": ("backwarp", "flex"),      # Flex-Rate.../b_model/b_model.py:99
    "FlowGuidedB": ("warp", "ac1"),            # ICIP2024/src/model/m.py:262
    "OffsetDiversity": ("warp", "ac1"),        # ICIP2024/src/model/helpers.py:61
    "DMC": ("warp", "ac1"),                    # OJSP2025/video_model.py:668
}


def _bind(obj, name, fn):
    setattr(obj, name, types.MethodType(fn, obj))


def patch(model, fuse=True):
    """Returns ``model`` with its hot operators re-bound.  ``fuse=True`` additionally replaces the LHBDC
    ``Model.forward`` by the fused pipeline (warp2 + blend kernel + bits-only likelihood passes); with
    ``fuse=False`` the reference's own ``forward`` keeps running and only the operator calls are swapped."""
    for mod in model.modules():
        cls = type(mod).__name__
        if cls == "GDN":
            _bind(mod, "forward", lambda self, x: modules.gdn_forward(self, x))
        elif cls == "ResidualBlockWithStride" and fuse:
            _bind(mod, "forward", modules.res_stride_forward)
        elif cls == "ResidualBlockUpsample" and fuse:
            _bind(mod, "forward", modules.res_upsample_forward)
        elif cls == "EntropyBottleneck":
            _bind(mod, "forward", modules.eb_forward)
            _bind(mod, "compress", modules.eb_compress)
            _bind(mod, "decompress", modules.eb_decompress)
        elif cls == "GaussianConditional":
            _bind(mod, "forward", modules.gc_forward)
            _bind(mod, "build_indexes", modules.gc_build_indexes)
            _bind(mod, "quantize", modules.gc_quantize)
            _bind(mod, "compress", modules.gc_compress)
            _bind(mod, "decompress", modules.gc_decompress)
        elif cls == "DeformConv2d" and hasattr(mod, "weight"):     # torchvision.ops.DeformConv2d
            _bind(mod, "forward", icip.deform_conv_forward)
        if hasattr(mod, "entropy_bottleneck") and hasattr(mod, "gaussian_conditional") and hasattr(mod, "g_a"):
            _bind(mod, "forward_bits", modules.hyperprior_forward_bits)
            _bind(mod, "symbols", modules.hyperprior_symbols)
        if cls in _WARP_BY_CLASS:
            name, variant = _WARP_BY_CLASS[cls]
            _bind(mod, name, lambda self, img, flow, _v=variant: ops.backwarp(img, flow, _v))
        if cls == "Network" and hasattr(mod, "netBasic"):
            if fuse:  # pyramid + per-level glue kernels (K-SPYPYR / K-SPYLEVEL), convolutions untouched
                _bind(mod, "forward", lhbdc.Network.forward)
            # SPyNet calls the module-level ``backwarp`` of LHBDC/model/flow.py (flow.py:98): swap the global.
            host = sys.modules.get(type(mod).__module__)
            if host is not None and hasattr(host, "backwarp"):
                host.backwarp = lambda tenInput, tenFlow: ops.backwarp(tenInput, tenFlow, "lhbdc")
            elif hasattr(mod, "_backwarp"):
                mod._backwarp = lambda tenInput, tenFlow: ops.backwarp(tenInput, tenFlow, "lhbdc")
        if cls == "Model" and fuse and all(hasattr(mod, a) for a in ("FlowNet", "mv_compressor", "masknet")):
            _bind(mod, "motion", lhbdc.Model.motion)
            _bind(mod, "forward_device", lhbdc.Model.forward_device)
            _bind(mod, "forward", lhbdc.Model.forward)
    return model
