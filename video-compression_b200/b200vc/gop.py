"""Hierarchical-GOP scheduling, GOP sharding and the per-GOP coding loop.

Mirrors the evaluation contract of the reference drivers (not their I/O / pandas bookkeeping):
``LHBDC/test/testing.py:70-74`` (GOP-8 coding order, reference pairs, hierarchy levels) and
``Flex-Rate.../test/testing.py:71-77`` (GOP-16).  GOPs only depend on their two anchor frames, so (a) frames of
the same hierarchy level -- within a GOP and across GOPs -- are coded as ONE batched ``Model`` call, and
(b) GOPs are the unit of multi-GPU sharding: contiguous blocks of GOPs per rank, no frame data
ever crosses GPUs, only the per-frame (bits, SSE) records are gathered.
"""
from dataclasses import dataclass

import torch

from . import ops


@dataclass(frozen=True)
class Schedule:
    gop: int
    refs: dict      # frame -> (past reference, future reference)
    levels: dict    # frame -> hierarchy level (0 = coded first)

    @property
    def order(self):
        return sorted(self.refs, key=lambda f: (self.levels[f], f))

    def by_level(self):
        out = {}
        for f in self.order:
            out.setdefault(self.levels[f], []).append(f)
        return [out[k] for k in sorted(out)]


# LHBDC/test/testing.py:70-74
LHBDC_GOP8 = Schedule(
    gop=8,
    refs={4: (0, 8), 2: (0, 4), 1: (0, 2), 3: (2, 4), 6: (4, 8), 5: (4, 6), 7: (6, 8)},
    levels={4: 0, 2: 1, 1: 2, 3: 2, 6: 1, 5: 2, 7: 2},
)
# Flex-Rate-Hier-Bidir-Video-Compression/test/testing.py:71-77
FLEX_GOP16 = Schedule(
    gop=16,
    refs={8: (0, 16), 4: (0, 8), 2: (0, 4), 1: (0, 2), 3: (2, 4), 6: (4, 8), 5: (4, 6), 7: (6, 8), 12: (8, 16),
          10: (8, 12), 9: (8, 10), 11: (10, 12), 14: (12, 16), 13: (12, 14), 15: (14, 16)},
    levels={8: 0, 4: 1, 2: 2, 1: 3, 3: 3, 6: 2, 5: 3, 7: 3, 12: 1, 10: 2, 9: 3, 11: 3, 14: 2, 13: 3, 15: 3},
)


# Flex-Rate.../test/testing.py:86-89: (I-frame quality, {hierarchy level: (gain row n, interpolation l)})
FLEX_QUALITIES = (
    (5, {0: (1, 1.), 1: (0, 0.33), 2: (0, 0.66), 3: (0, 1.)}), (6, {0: (1, 0.66), 1: (1, 1.), 2: (0, 0.33), 3: (0, 0.66)}),
    (6, {0: (1, 0.33), 1: (1, 0.66), 2: (1, 1.), 3: (0, 0.33)}), (6, {0: (2, 1.), 1: (1, 0.33), 2: (1, 0.66), 3: (1, 1.)}),
    (7, {0: (2, 0.66), 1: (2, 1.), 2: (1, 0.33), 3: (1, 0.66)}), (7, {0: (2, 0.33), 1: (2, 0.66), 2: (2, 1.), 3: (1, 0.33)}),
    (7, {0: (3, 1.), 1: (2, 0.33), 2: (2, 0.66), 3: (2, 1.)}), (8, {0: (3, 1.), 1: (3, 1.), 2: (3, 1.), 3: (2, 0.33)}),
)


def num_gops(num_frames, gop):
    """GOP k covers frames [k*gop, (k+1)*gop] (anchors shared); incomplete tails are dropped
    (``drop_last=True``, LHBDC/test/testing.py:117-120)."""
    return max(0, (num_frames - 1) // gop)


def shard_units(num_units, world_size, rank):
    """Contiguous block assignment: unit u -> rank u*R//U.  Returns range(lo, hi)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    lo = (rank * num_units) // world_size
    hi = ((rank + 1) * num_units) // world_size
    return range(lo, hi)


def psnr_from_sse(sse, n_values, data_range=255.0):
    """PSNR from a sum of squared uint8 differences (LHBDC/test/testing.py:176-182, test/utils.py PSNR/MSE)."""
    return 10.0 * torch.log10(torch.as_tensor(data_range ** 2, dtype=torch.float64) * n_values / sse)


class GopCoder:
    """Codes GOPs of already-resident frames with a B-frame ``model`` exposing
    ``forward_device(x_before, x_current, x_after) -> (x_hat, bits[N], ...)`` (LHBDC ``Model``) or, with
    ``level_quality = {hierarchy level: (n, l)}`` (one entry of ``FLEX_QUALITIES``; Flex-Rate.../test/testing.py:
    193-200 passes ``n=[n], l=l`` per frame level), ``forward_device(x_before, x_current, x_after, [n], l)``
    (``BidirFlowRef``).

    ``code(frames, crop)``: frames [G, gop+1, 3, H, W] (G independent GOPs, padded size).  Without an
    ``anchor_codec`` the anchors are used as-is (uncoded: I-frame coding is outside the B-frame hot path); with one
    (``b200vc.modules.mbt2018_mean(q)`` or any module with ``forward_bits``) both anchors of every GOP go through
    ``image_compress`` first, as the reference's evaluation loop does (LHBDC/test/testing.py:127-152: ``decoded[0]``,
    ``decoded[8]`` are the I-codec's reconstructions and their sizes are booked as I-frames) -- a GOP's last anchor
    is coded again as the next GOP's first one, which is what keeps the GOPs independent units.  Returns device tensors
    ``bits[G, gop+1]`` (0 for uncoded anchors) and ``sse[G, gop+1]`` (uint8-domain, over the unpadded crop) without any
    host synchronisation.
    """

    def __init__(self, model, schedule=LHBDC_GOP8, level_quality=None, anchor_codec=None):
        self.model = model
        self.schedule = schedule
        self.level_quality = level_quality
        self.anchor_codec = anchor_codec

    @torch.no_grad()
    def code(self, frames, crop, want_decoded=False):
        sch = self.schedule
        G, T, C, H, W = frames.shape
        if T != sch.gop + 1:
            raise ValueError(f"expected GOPs of {sch.gop + 1} frames, got {T}")
        h, w = crop
        dev = frames.device
        decoded = {0: frames[:, 0], sch.gop: frames[:, sch.gop]}
        bits = torch.zeros((G, T), device=dev, dtype=torch.float64)
        sse = torch.zeros((G, T), device=dev, dtype=torch.float64)
        if self.anchor_codec is not None:
            from .modules import image_compress
            both = torch.cat([frames[:, 0], frames[:, sch.gop]], 0)       # first and last anchors of all GOPs: one call
            dec, size = image_compress(both, self.anchor_codec)
            a_sse = ops.sse_u8(dec, both, h, w)
            decoded = {0: dec[:G], sch.gop: dec[G:]}
            bits[:, 0], bits[:, sch.gop] = size[:G], size[G:]
            sse[:, 0], sse[:, sch.gop] = a_sse[:G], a_sse[G:]
        for level, level_frames in enumerate(sch.by_level()):
            xb = torch.cat([decoded[sch.refs[f][0]] for f in level_frames], 0)
            xa = torch.cat([decoded[sch.refs[f][1]] for f in level_frames], 0)
            xc = torch.cat([frames[:, f] for f in level_frames], 0)
            if self.level_quality is None:
                x_hat, b = self.model.forward_device(xb, xc, xa)[:2]
            else:
                n, l = self.level_quality[level]
                x_hat, b = self.model.forward_device(xb, xc, xa, [n], l)[:2]
            level_sse = ops.sse_u8(x_hat, xc, h, w)   # one launch scores every frame of the level
            for k, f in enumerate(level_frames):
                sl = slice(k * G, (k + 1) * G)
                decoded[f] = x_hat[sl]
                bits[:, f] = b[sl]
                sse[:, f] = level_sse[sl]
        if want_decoded:
            return bits, sse, torch.stack([decoded[t] for t in range(T)], 1)
        return bits, sse
