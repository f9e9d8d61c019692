"""b200vc -- B200 (sm_100a) native hot path for the KUIS-AI LHBDC / Flex-Rate hierarchical B-frame codecs:
flow-driven bilinear backward warp, GDN/IGDN, quantise + Gaussian-conditional / factorised-prior likelihood.

Layout
  csrc/ (one level up)  hand-written CUDA kernels + the C-ABI of include/b200vc.h  -> libb200vc.so
  _lib.py               ctypes binding (the only way into the kernels; no fallback)
  ops.py                tensor-level wrappers (one per reference call chain)
  modules.py            CompressAI-interface mirror (GDN, EntropyBottleneck, GaussianConditional, blocks)
  lhbdc.py              LHBDC model mirror (Model, Network, MVCompressor, ResidualCompressor, Mask)
  icip.py               ICIP2024 down-ratio search, deformable alignment, ELIC checkerboard context loop
  flowguided.py         ICIP2024 model mirror (FlowGuidedB, Offset_ELIC, Res_ELIC, ...) + its sequence coding loop
  ojsp.py               OJSP2025 down-sampling-ratio search (optimize_down_sampling_ratio, warp_psnr)
  flexrate.py           Flex-Rate model mirror (BidirFlowRef, Gain_Module, FlowCompressor, ResidualCompressor, UNet)
  patch.py              patch(model): swap the kernels into a model built from the reference's own classes
  coding.py             CDF tables, GPU rANS coder, LHBDC .bin container (compress / decompress / encode_B / decode_B)
  gop.py / dist.py      hierarchical-GOP schedule, GOP sharding over ranks, record gathering
  synthetic.py          seeded synthetic video + weight calibration (no datasets / checkpoints offline)
"""
from . import _lib, coding, dist, flexrate, flowguided, gop, icip, lhbdc, modules, ojsp, ops, synthetic  # noqa: F401
from .flexrate import BidirFlowRef  # noqa: F401
from .lhbdc import Model, decode_B, encode_B, encode_B_symbols  # noqa: F401
from .patch import patch  # noqa: F401

__version__ = "0.1.0"
