"""TEST INFRASTRUCTURE (oracle).  Pure-Python restatement of b200vc's "b2r1" rANS container (csrc/rans.cu): the
model side follows CompressAI's coder interface (``BufferedRansEncoder.encode_with_indexes`` /
``RansDecoder.decode_with_indexes``: per-symbol CDF row, offset, tail-bin escape), which the reference reaches through
``.compress()`` / ``.decompress()`` (LHBDC/model/layers.py:93-117).  parity unpinned against CompressAI's own byte
stream (not installable offline); pinned here: GPU bytes == these bytes, and decode(encode(x)) == x.
Small inputs only (Python loops)."""
import struct

import numpy as np

L = 1 << 16
MAGIC = b"b2r1"


def encode(symbols, indexes, cdf, cdf_len, offset, stream_len):
    symbols = [int(v) for v in np.asarray(symbols).reshape(-1)]
    indexes = [int(v) for v in np.asarray(indexes).reshape(-1)]
    n = len(symbols)
    if n == 0:
        return MAGIC + struct.pack("<III", 0, stream_len, 0)
    streams = []
    for lo in range(0, n, stream_len):
        hi = min(lo + stream_len, n)
        x, words = L, []  # words are produced in reverse order of reading

        def put(start, freq):
            nonlocal x
            if x >= (freq << 16):
                words.append(x & 0xFFFF)
                x >>= 16
            x = ((x // freq) << 16) + (x % freq) + start

        for i in range(hi - 1, lo - 1, -1):
            row = indexes[i]
            max_value = int(cdf_len[row]) - 2
            value = symbols[i] - int(offset[row])
            if value < 0 or value >= max_value:
                raw = (-2 * value - 1) if value < 0 else 2 * (value - max_value)
                put(raw >> 16, 1)
                put(raw & 0xFFFF, 1)
                value = max_value
            start = int(cdf[row][value])
            put(start, int(cdf[row][value + 1]) - start)
        words.append(x & 0xFFFF)
        words.append(x >> 16)
        streams.append(words[::-1])
    sizes = np.array([len(w) for w in streams], dtype="<u4")
    payload = np.array([w for s in streams for w in s], dtype="<u2")
    return MAGIC + struct.pack("<III", n, stream_len, len(streams)) + sizes.tobytes() + payload.tobytes()


def decode(data, indexes, cdf, cdf_len, offset):
    assert data[:4] == MAGIC
    n, stream_len, S = struct.unpack_from("<III", data, 4)
    sizes = np.frombuffer(data, dtype="<u4", count=S, offset=16).astype(np.int64)
    payload = np.frombuffer(data, dtype="<u2", offset=16 + 4 * S)
    indexes = [int(v) for v in np.asarray(indexes).reshape(-1)]
    out = np.zeros(n, dtype=np.int32)
    pos = 0
    for s in range(S):
        words = [int(w) for w in payload[pos:pos + sizes[s]]]
        pos += int(sizes[s])
        x = (words[0] << 16) | words[1]
        p = 2

        def advance(start, freq):
            nonlocal x, p
            x = freq * (x >> 16) + (x & 0xFFFF) - start
            if x < L:
                x = (x << 16) | words[p]
                p += 1

        for i in range(s * stream_len, min((s + 1) * stream_len, n)):
            row = indexes[i]
            max_value = int(cdf_len[row]) - 2
            slot = x & 0xFFFF
            v = int(np.searchsorted(np.asarray(cdf[row][:max_value + 1]), slot, side="right")) - 1
            start = int(cdf[row][v])
            advance(start, int(cdf[row][v + 1]) - start)
            value = v
            if v == max_value:
                lo16 = x & 0xFFFF
                advance(lo16, 1)
                hi16 = x & 0xFFFF
                advance(hi16, 1)
                raw = (hi16 << 16) | lo16
                value = -((raw + 1) >> 1) if raw & 1 else max_value + (raw >> 1)
            out[i] = value + int(offset[row])
        assert p == len(words), "stream not fully consumed"
    return out
