"""CPU: CDF-table quantiser and the numpy oracle of the b2r1 rANS container (round trips, escapes, ragged tails)."""
import numpy as np
import pytest

from b200vc import coding
from oracle import rans as o_rans


def _tables(rng, rows=5, max_len=40):
    cdf = np.zeros((rows, max_len + 2), dtype=np.int32)
    cdf_len, offset = np.zeros(rows, np.int32), np.zeros(rows, np.int32)
    for r in range(rows):
        n = int(rng.integers(1, max_len + 1))
        pmf = rng.random(n).astype(np.float32) ** 4 + 1e-9
        pmf /= pmf.sum() * 1.02
        q = coding.pmf_to_quantized_cdf(np.concatenate([pmf, [max(1e-9, 1 - pmf.sum())]]).astype(np.float32))
        cdf[r, :q.size] = q
        cdf_len[r] = n + 2
        offset[r] = -(n // 2)
    return cdf, cdf_len, offset


def test_pmf_to_quantized_cdf_properties():
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 300):
        pmf = rng.random(n).astype(np.float32) ** 8
        pmf[rng.random(n) < 0.5] = 0.0  # empty bins must be made codable
        pmf[0] = max(pmf[0], 0.3)
        cdf = coding.pmf_to_quantized_cdf(pmf / pmf.sum())
        assert cdf[0] == 0 and cdf[-1] == 65536 and cdf.size == n + 1
        assert (np.diff(cdf) >= 1).all()
    with pytest.raises(ValueError):
        coding.pmf_to_quantized_cdf(np.zeros(4, np.float32))
    # a delta distribution over many bins: every other bin steals exactly one count
    pmf = np.zeros(100, np.float32)
    pmf[50] = 1.0
    cdf = coding.pmf_to_quantized_cdf(pmf)
    f = np.diff(cdf)
    assert f[50] == 65536 - 99 and (np.delete(f, 50) == 1).all()


@pytest.mark.parametrize("n,stream_len", [(1, 8), (8, 8), (9, 8), (1000, 64), (257, 4096)])
def test_oracle_round_trip_with_escapes(n, stream_len):
    rng = np.random.default_rng(n)
    cdf, cdf_len, offset = _tables(rng)
    idx = rng.integers(0, cdf.shape[0], n).astype(np.int32)
    sym = np.array([int(rng.integers(offset[r] - 3, offset[r] + cdf_len[r] + 1)) for r in idx], dtype=np.int32)
    sym[::17] += 100000          # far escapes (both zig-zag branches)
    sym[5::29] -= 70000
    data = o_rans.encode(sym, idx, cdf, cdf_len, offset, stream_len)
    assert data[:4] == b"b2r1"
    got = o_rans.decode(data, idx, cdf, cdf_len, offset)
    assert (got == sym).all()
    nn, sl, sizes, payload = coding.parse_container(data)
    assert nn == n and sl == stream_len and sizes.size == -(-n // stream_len) and payload.size == sizes.sum()


def test_empty_and_truncated_streams():
    assert coding.parse_container(o_rans.encode([], [], None, None, None, 16))[0] == 0
    rng = np.random.default_rng(3)
    cdf, cdf_len, offset = _tables(rng)
    data = o_rans.encode(np.zeros(40, np.int32), np.zeros(40, np.int32), cdf, cdf_len, offset, 16)
    with pytest.raises(ValueError, match="truncated"):
        coding.parse_container(data[:-2])
    with pytest.raises(ValueError, match="b2r1"):
        coding.parse_container(b"xxxx" + data[4:])


def test_container_header_is_validated_before_anything_reaches_the_gpu():
    """An untrusted .bin: n / stream_len / S / sizes must be mutually consistent (the GPU decoder indexes with them)."""
    import struct
    rng = np.random.default_rng(4)
    cdf, cdf_len, offset = _tables(rng)
    data = o_rans.encode(np.zeros(40, np.int32), np.zeros(40, np.int32), cdf, cdf_len, offset, 16)   # 3 streams
    n, sl, S = struct.unpack_from("<III", data, 4)
    assert (n, sl, S) == (40, 16, 3)
    hdr = lambda n, sl, S: data[:4] + struct.pack("<III", n, sl, S) + data[16:]
    with pytest.raises(ValueError, match="header mismatch"):
        coding.parse_container(hdr(40, 16, 2))        # S too small: the kernel would index offsets[] past its end
    with pytest.raises(ValueError, match="header mismatch"):
        coding.parse_container(hdr(400, 16, 3))       # more symbols than the streams can hold
    with pytest.raises(ValueError, match="header mismatch"):
        coding.parse_container(hdr(40, 0, 3))         # stream_len 0
    with pytest.raises(ValueError, match="truncated"):
        coding.parse_container(data[:20])
    bad = bytearray(data)
    struct.pack_into("<I", bad, 16, 1)                # a stream shorter than its two state words
    with pytest.raises(ValueError, match="stream size"):
        coding.parse_container(bytes(bad))
    with pytest.raises(ValueError, match="b2r1"):
        coding.parse_container(b"b2r1")


def test_cost_tracks_the_model():
    """Coding i.i.d. symbols from a row's own pmf costs ~ its entropy (+ framing)."""
    rng = np.random.default_rng(7)
    pmf = np.array([0.6, 0.2, 0.1, 0.05, 0.03, 0.02], np.float32)
    q = coding.pmf_to_quantized_cdf(np.concatenate([pmf * 0.999, [0.001]]).astype(np.float32))
    cdf = q[None, :].astype(np.int32)
    n = 4000
    sym = rng.choice(6, size=n, p=pmf / pmf.sum()).astype(np.int32)
    data = o_rans.encode(sym, np.zeros(n, np.int32), cdf, np.array([8], np.int32), np.array([0], np.int32), 4096)
    ent = -(pmf / pmf.sum() * np.log2(pmf / pmf.sum())).sum() * n
    bits = 8 * (len(data) - 20)
    assert ent < bits < ent * 1.03 + 64


def test_bin_container_round_trip(tmp_path):
    mv = {"strings": [[b"abc"], [b"de"]], "shape": (5, 8)}
    res = {"strings": [[b"fghij"], [b"klmnopq"]], "shape": (17, 30)}
    p = tmp_path / "bits_B.bin"
    coding.write_bin(str(p), 1626, mv, res)
    raw = p.read_bytes()
    assert len(raw) == 4 + 4 + 4 + 4 + 4 + 4 + 3 + 2 + 5 + 7   # LHBDC/encode_B.py:114-126 layout
    lam, s_mv, sh_mv, s_res, sh_res = coding.read_bin(str(p))
    assert lam == 1626 and tuple(sh_mv) == (5, 8) and tuple(sh_res) == (17, 30)
    assert s_mv == [[b"abc"], [b"de"]] and s_res == [[b"fghij"], [b"klmnopq"]]
