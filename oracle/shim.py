"""TEST INFRASTRUCTURE (oracle).  Installs a ``compressai`` stand-in built from ``oracle/cai.py`` into
``sys.modules`` so the reference's own model files (``LHBDC/model/{m,layers}.py``,
``Flex-Rate.../b_model/{b_model,layers}.py``, ``ICIP2024/src/model/m.py``) import **verbatim** in the build
container.  Used only by ``oracle/make_golden.py`` (which needs ``/root/reference``; the GPU box never
runs it) to pin the restated warps / models against the reference's code.
"""
import importlib
import sys
import types

from . import cai


def install():
    if "compressai" in sys.modules and getattr(sys.modules["compressai"], "__oracle_shim__", False):
        return sys.modules["compressai"]
    root = types.ModuleType("compressai")
    root.__oracle_shim__ = True
    root.__path__ = []  # mark as package

    def sub(name, **attrs):
        m = types.ModuleType(f"compressai.{name}")
        m.__dict__.update(attrs)
        sys.modules[f"compressai.{name}"] = m
        setattr(root, name.split(".")[0], sys.modules[f"compressai.{name.split('.')[0]}"])
        return m

    models = sub(
        "models",
        MeanScaleHyperprior=cai.MeanScaleHyperprior,
        ScaleHyperprior=cai.ScaleHyperprior,
        CompressionModel=cai.CompressionModel,
        JointAutoregressiveHierarchicalPriors=cai.JointAutoregressiveHierarchicalPriors,
        Cheng2020Anchor=cai.Cheng2020Anchor,
    )
    models.__path__ = []
    mutils = sub("models.utils", conv=cai._conv5, deconv=cai._deconv5)
    models.utils = mutils
    sub(
        "entropy_models",
        EntropyBottleneck=cai.EntropyBottleneck,
        GaussianConditional=cai.GaussianConditional,
        EntropyModel=cai.EntropyModel,
    )
    sub(
        "layers",
        GDN=cai.GDN,
        MaskedConv2d=cai.MaskedConv2d,
        AttentionBlock=cai.AttentionBlock,
        ResidualBlock=cai.ResidualBlock,
        ResidualBlockUpsample=cai.ResidualBlockUpsample,
        ResidualBlockWithStride=cai.ResidualBlockWithStride,
        conv3x3=cai.conv3x3,
        conv1x1=cai.conv1x1,
        subpel_conv3x3=cai.subpel_conv3x3,
    )
    sub("ops", LowerBound=cai.LowerBound, NonNegativeParametrizer=cai.NonNegativeParametrizer)

    class _NoCoder:  # compressai.ans: the C++ rANS coder is imported by the ICIP model files, never built offline
        def __init__(self, *a, **k):
            raise NotImplementedError("compressai.ans is not available offline (oracle shim)")

    sub("ans", BufferedRansEncoder=_NoCoder, RansDecoder=_NoCoder)
    sys.modules["compressai"] = root
    return root


def import_reference(pkg_dir, module):
    """Import ``module`` (dotted, e.g. ``model.m``) with ``pkg_dir`` first on sys.path, through the shim."""
    install()
    sys.path.insert(0, pkg_dir)
    try:
        for k in [k for k in sys.modules if k == module.split(".")[0] or k.startswith(module.split(".")[0] + ".")]:
            del sys.modules[k]
        return importlib.import_module(module)
    finally:
        sys.path.remove(pkg_dir)
