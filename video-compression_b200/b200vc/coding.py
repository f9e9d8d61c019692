"""Entropy-coding layer behind ``.compress()`` / ``.decompress()``.

Replaces what the reference gets from ``compressai.ans`` + ``EntropyModel.update`` (``LHBDC/encode_B.py:34-35,96,104``,
``LHBDC/model/layers.py:93-117``): quantised CDF tables, the range-ANS coder and the ``.bin`` container of
``LHBDC/encode_B.py:114-126`` / ``decode_B.py:88-104``.

* Table construction follows CompressAI 1.1.x (``GaussianConditional.update_scale_table``,
  ``EntropyBottleneck.update``, C++ ``pmf_to_quantized_cdf``) with the pmf evaluated by the b200vc likelihood kernels;
  it is a once-per-model preparation step (numpy on the host, as in CompressAI), not part of the per-frame path.
* The coder itself runs on the GPU (``csrc/rans.cu``).  Its container ("b2r1") is b200vc's own: CompressAI's
  byte-exact stream cannot be verified offline, so parity here is a self round trip + a numpy oracle of the same format.
"""
import struct

import numpy as np
import torch

from . import _lib, ops

PRECISION = 16
STREAM_LEN = 4096      # symbols per independent rANS stream (0.016 bit/symbol of framing overhead)
MAGIC = b"b2r1"


def pmf_to_quantized_cdf(pmf, precision=PRECISION):
    """CompressAI's C++ ``pmf_to_quantized_cdf``: round to ``precision`` bits, renormalise, make every bin non-empty
    by stealing from the smallest bin with frequency > 1."""
    pmf = np.asarray(pmf, dtype=np.float32)
    cdf = np.zeros(pmf.size + 1, dtype=np.int64)
    cdf[1:] = np.floor(pmf.astype(np.float64) * (1 << precision) + 0.5).astype(np.int64)  # std::round, values >= 0
    total = int(cdf.sum())
    if total == 0:
        raise ValueError("pmf_to_quantized_cdf: empty pmf")
    cdf = ((1 << precision) * cdf) // total
    cdf = np.cumsum(cdf)
    cdf[-1] = 1 << precision
    n = cdf.size - 1
    for i in range(n):
        if cdf[i] == cdf[i + 1]:
            freqs = np.diff(cdf)
            cand = np.where(freqs > 1, freqs, np.iinfo(np.int64).max)
            best = int(np.argmin(cand))
            if cand[best] == np.iinfo(np.int64).max:
                raise ValueError("pmf_to_quantized_cdf: cannot make all bins non-empty")
            if best < i:
                cdf[best + 1:i + 1] -= 1
            else:
                cdf[i + 1:best + 1] += 1
    return cdf.astype(np.int32)


def _pmf_to_cdf(pmf, tail_mass, pmf_length, max_length):
    cdf = np.zeros((pmf.shape[0], max_length + 2), dtype=np.int32)
    for i in range(pmf.shape[0]):
        row = np.concatenate([pmf[i, :pmf_length[i]], tail_mass[i:i + 1]])
        q = pmf_to_quantized_cdf(row)
        cdf[i, :q.size] = q
    return cdf


class Tables:
    """Device-resident CDF rows: cdf [rows, stride] int32, cdf_length [rows], offset [rows]."""

    def __init__(self, cdf, cdf_length, offset, device):
        self.cdf = torch.as_tensor(cdf, dtype=torch.int32, device=device).contiguous()
        self.cdf_length = torch.as_tensor(cdf_length, dtype=torch.int32, device=device).contiguous()
        self.offset = torch.as_tensor(offset, dtype=torch.int32, device=device).contiguous()

    @property
    def stride(self):
        return self.cdf.shape[1]


def gaussian_tables(scale_table, device, tail_mass=1e-9, scale_bound=0.11):
    """GaussianConditional.update_scale_table: pmf over |k - centre| for every table scale."""
    from scipy.stats import norm
    table = torch.as_tensor(scale_table, dtype=torch.float32)
    multiplier = -float(norm.ppf(tail_mass / 2))
    center = torch.ceil(table * multiplier).int()
    length = 2 * center + 1
    max_length = int(length.max())
    samples = (torch.arange(max_length).int()[None, :] - center[:, None]).abs().float()    # [T, L]
    scales = table[:, None].expand_as(samples).contiguous()
    r = ops.gauss_cond(samples.view(1, 1, *samples.shape).to(device), scales.view(1, 1, *scales.shape).to(device),
                       torch.zeros(1, 1, *samples.shape, device=device), scale_bound=min(scale_bound, float(table[0])),
                       lik_bound=0.0, want_y_hat=False, want_bits=False)
    pmf = r["lik"].view(samples.shape).cpu().numpy()
    length_np = length.numpy()
    for i in range(pmf.shape[0]):
        pmf[i, length_np[i]:] = 0.0
    tail = np.clip(1.0 - pmf.astype(np.float64).sum(axis=1), 0.0, None).astype(np.float32)   # = 2 * lower tail
    cdf = _pmf_to_cdf(pmf, tail, length_np, max_length)
    return Tables(cdf, length_np + 2, -center.numpy(), device)


def bottleneck_tables(packed, quantiles, device):
    """EntropyBottleneck.update: per-channel pmf on the integer support spanned by the learned quantiles."""
    q = quantiles.detach().float().cpu()
    medians = q[:, 0, 1]
    minima = torch.clamp(torch.ceil(medians - q[:, 0, 0]).int(), min=0)
    maxima = torch.clamp(torch.ceil(q[:, 0, 2] - medians).int(), min=0)
    start = medians - minima.float()
    length = (maxima + minima + 1).numpy()
    max_length = int(length.max())
    samples = torch.arange(max_length).float()[None, :] + start[:, None]                      # [C, L]
    C = samples.shape[0]
    r = ops.entropy_bottleneck(samples.view(1, C, 1, max_length).to(device).contiguous(), packed, lik_bound=0.0,
                               want_z_hat=False, want_bits=False)
    pmf = r["lik"].view(C, max_length).cpu().numpy()
    for i in range(C):
        pmf[i, length[i]:] = 0.0
    tail = np.clip(1.0 - pmf.astype(np.float64).sum(axis=1), 0.0, None).astype(np.float32)
    cdf = _pmf_to_cdf(pmf, tail, length, max_length)
    return Tables(cdf, length + 2, -minima.numpy(), device)


# ------------------------------------------------------------------------------------------ coder
def rans_encode(symbols, indexes, tables, stream_len=STREAM_LEN):
    """int32 symbols + CDF-row indexes (flattened, same length) -> bytes ("b2r1" container)."""
    sym = symbols.reshape(-1).to(torch.int32).contiguous()
    idx = indexes.reshape(-1).to(torch.int32).contiguous()
    n = sym.numel()
    if idx.numel() != n:
        raise RuntimeError("rans_encode: symbols / indexes length mismatch")
    if n == 0:
        return MAGIC + struct.pack("<III", 0, stream_len, 0)
    if not sym.is_cuda:
        raise RuntimeError("rans_encode: expected CUDA tensors (b200vc has no CPU fallback)")
    lib = _lib.load()
    S = (n + stream_len - 1) // stream_len
    words = lib.b200vc_rans_scratch_words(stream_len)
    scratch = torch.empty(S * words, dtype=torch.int16, device=sym.device)
    sizes = torch.empty(S, dtype=torch.int32, device=sym.device)
    st = torch.cuda.current_stream().cuda_stream
    ops._run("rans_encode", 8 * n, lambda: lib.b200vc_rans_encode(
        sym.data_ptr(), idx.data_ptr(), tables.cdf.data_ptr(), tables.cdf_length.data_ptr(), tables.offset.data_ptr(),
        tables.stride, n, stream_len, scratch.data_ptr(), sizes.data_ptr(), st))
    sizes64 = sizes.to(torch.int64)
    offsets = torch.cumsum(sizes64, 0) - sizes64
    total = int(sizes64.sum().item())
    out = torch.empty(total, dtype=torch.int16, device=sym.device)
    ops._run("rans_compact", 4 * total, lambda: lib.b200vc_rans_compact(
        scratch.data_ptr(), stream_len, sizes.data_ptr(), offsets.data_ptr(), S, out.data_ptr(), st))
    header = MAGIC + struct.pack("<III", n, stream_len, S) + sizes.cpu().numpy().astype("<u4").tobytes()
    return header + out.cpu().numpy().astype("<i2").tobytes()


def parse_container(data):
    """Header of an (untrusted) b2r1 container -> (n, stream_len, sizes[S], payload words).  Everything the GPU
    decoder will index with is checked here: the stream count must be the one ``n`` and ``stream_len`` imply, every
    stream holds at least its two state words and at most the encoder's worst case, and the payload length must
    equal the sum of the stream sizes."""
    if len(data) < 16 or data[:4] != MAGIC:
        raise ValueError("rans_decode: not a b2r1 stream")
    n, stream_len, S = struct.unpack_from("<III", data, 4)
    if n == 0:
        if S != 0 or len(data) != 16:
            raise ValueError("rans_decode: malformed empty stream")
        return 0, stream_len, np.zeros(0, dtype=np.int64), np.zeros(0, dtype="<u2")
    if stream_len == 0 or S != (n + stream_len - 1) // stream_len:
        raise ValueError(f"rans_decode: header mismatch ({n} symbols in streams of {stream_len} need "
                         f"{(n + stream_len - 1) // max(stream_len, 1)} streams, header says {S})")
    if len(data) < 16 + 4 * S or (len(data) - 16 - 4 * S) % 2:
        raise ValueError("rans_decode: truncated header")
    sizes = np.frombuffer(data, dtype="<u4", count=S, offset=16).astype(np.int64)
    payload = np.frombuffer(data, dtype="<u2", offset=16 + 4 * S)
    if sizes.min() < 2 or sizes.max() > 3 * stream_len + 2:
        raise ValueError("rans_decode: a stream size is outside [2, 3 * stream_len + 2] words")
    if payload.size != int(sizes.sum()):
        raise ValueError("rans_decode: truncated payload")
    return n, stream_len, sizes, payload


def rans_decode(data, indexes, tables):
    """bytes + CDF-row indexes (flattened) -> int32 symbols on the indexes' device.  Raises ``ValueError`` for a
    malformed or corrupt container (never an out-of-bounds device access)."""
    idx = indexes.reshape(-1).to(torch.int32).contiguous()
    n, stream_len, sizes, payload = parse_container(data)
    if idx.numel() != n:
        raise RuntimeError(f"rans_decode: stream holds {n} symbols, {idx.numel()} indexes given")
    if n == 0:
        return torch.empty(0, dtype=torch.int32, device=idx.device)
    if not idx.is_cuda:
        raise RuntimeError("rans_decode: expected CUDA tensors (b200vc has no CPU fallback)")
    dev = idx.device
    offsets = torch.from_numpy(np.cumsum(sizes) - sizes).to(dev)
    sizes32 = torch.from_numpy(sizes.astype(np.int32)).to(dev)
    words = torch.from_numpy(payload.astype(np.int16)).to(dev)
    out = torch.empty(n, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    lib = _lib.load()
    ops._run("rans_decode", 8 * n, lambda: lib.b200vc_rans_decode(
        words.data_ptr(), offsets.data_ptr(), sizes32.data_ptr(), idx.data_ptr(), tables.cdf.data_ptr(),
        tables.cdf_length.data_ptr(), tables.offset.data_ptr(), tables.stride, n, stream_len, out.data_ptr(),
        status.data_ptr(), torch.cuda.current_stream().cuda_stream))
    if status.item() != 0:
        raise ValueError("rans_decode: corrupt stream (a stream over-ran, stopped short of its end or did not "
                         "return to the coder's initial state)")
    return out


# ------------------------------------------------------------------- .bin container (LHBDC/encode_B.py:114-126)
def write_bin(path, lam, mv_bits, res_bits):
    """uint32 lambda | uint16x2 mv z-shape | uint32 len(mv y) | uint32 len(mv z) | uint16x2 res z-shape |
    uint32 len(res y) | mv y | mv z | res y | res z (last length implicit)."""
    with open(path, "wb") as f:
        f.write(np.array(lam, dtype=np.uint32).tobytes())
        f.write(np.array(tuple(mv_bits["shape"]), dtype=np.uint16).tobytes())
        f.write(np.array(len(mv_bits["strings"][0][0]), dtype=np.uint32).tobytes())
        f.write(np.array(len(mv_bits["strings"][1][0]), dtype=np.uint32).tobytes())
        f.write(np.array(tuple(res_bits["shape"]), dtype=np.uint16).tobytes())
        f.write(np.array(len(res_bits["strings"][0][0]), dtype=np.uint32).tobytes())
        for s in (mv_bits["strings"][0][0], mv_bits["strings"][1][0], res_bits["strings"][0][0],
                  res_bits["strings"][1][0]):
            f.write(s)


def read_bin(path):
    """Inverse of ``write_bin`` (LHBDC/decode_B.py:88-104): (lambda, mv strings, mv shape, res strings, res shape)."""
    with open(path, "rb") as f:
        lam = int(np.frombuffer(f.read(4), dtype=np.uint32)[0])
        shape_mv = torch.Size(np.frombuffer(f.read(4), dtype=np.uint16).astype(int).tolist())
        len0_mv = int(np.frombuffer(f.read(4), dtype=np.uint32)[0])
        len1_mv = int(np.frombuffer(f.read(4), dtype=np.uint32)[0])
        shape_res = torch.Size(np.frombuffer(f.read(4), dtype=np.uint16).astype(int).tolist())
        len0_res = int(np.frombuffer(f.read(4), dtype=np.uint32)[0])
        s0_mv, s1_mv, s0_res, s1_res = f.read(len0_mv), f.read(len1_mv), f.read(len0_res), f.read()
    return lam, [[s0_mv], [s1_mv]], shape_mv, [[s0_res], [s1_res]], shape_res
