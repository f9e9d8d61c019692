// K-RANS: range-ANS entropy coder on the GPU (replaces the CPU `compressai.ans`
// BufferedRansEncoder / RansDecoder the reference reaches through `.compress()` / `.decompress()`,
// LHBDC/model/layers.py:93-117, and the Python-list marshalling of ~1 M symbols per frame).
//
// CompressAI's coder is one sequential 64-bit rANS stream; bit-compatibility with it cannot be verified offline
// (compressai is not installable here), so the container is b200vc's own and parity is a self round trip
// (DESIGN.md 8).  Format "b2r1": the symbol sequence is cut into independent streams of `stream_len` symbols; each
// stream is a 32-bit-state / 16-bit-word rANS with 16-bit probabilities (CompressAI's precision), coded by one
// thread.  The model side is CompressAI's: per-symbol CDF row `indexes[i]`, value = symbol - offset[row]; values
// outside [0, max_value) use the tail bin followed by a zig-zag raw value pushed as two uniform 16-bit symbols.
//
// Integer-only, data-dependent and serial per stream: bound by per-thread latency, not by HBM.
#include "common.cuh"

namespace b200vc {

constexpr uint32_t kRansL = 1u << 16;  // state lower bound; probabilities and renormalisation words are 16 bit

struct RansTables {
  const int32_t* cdf;      // [rows, stride]
  const int32_t* cdf_len;  // [rows]  (= pmf_length + 2)
  const int32_t* offset;   // [rows]
  int stride;
};

// ---- encoder: symbols are pushed in reverse, words are written backwards into the thread's scratch row -------
struct RansEnc {
  uint32_t x;
  uint16_t* base;
  int pos;  // next free slot is base[pos - 1]
  __device__ __forceinline__ void put(uint32_t start, uint32_t freq) {
    if ((uint64_t)x >= ((uint64_t)freq << 16)) {  // x_max = ((L >> 16) << 16) * freq  (freq may be 65536)
      base[--pos] = (uint16_t)(x & 0xFFFFu);
      x >>= 16;
    }
    x = ((x / freq) << 16) + (x % freq) + start;
  }
};

__global__ void __launch_bounds__(128)
rans_encode_kernel(const int32_t* __restrict__ symbols, const int32_t* __restrict__ indexes, RansTables t,
                   int64_t n_symbols, int stream_len, uint16_t* __restrict__ scratch, int scratch_stride,
                   int32_t* __restrict__ sizes, int n_streams) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_streams) return;
  const int64_t lo = (int64_t)s * stream_len;
  const int64_t hi = lo + stream_len < n_symbols ? lo + stream_len : n_symbols;
  RansEnc e;
  e.x = kRansL;
  e.base = scratch + (int64_t)s * scratch_stride;
  e.pos = scratch_stride;
  for (int64_t i = hi - 1; i >= lo; --i) {
    const int row = __ldg(indexes + i);
    const int32_t* cdf = t.cdf + (int64_t)row * t.stride;
    const int max_value = __ldg(t.cdf_len + row) - 2;
    int value = __ldg(symbols + i) - __ldg(t.offset + row);
    if (value < 0 || value >= max_value) {
      const uint32_t raw = value < 0 ? (uint32_t)(-2 * value - 1) : (uint32_t)(2 * (value - max_value));
      e.put(raw >> 16, 1u);      // decoder reads: tail symbol, low 16 bits, high 16 bits
      e.put(raw & 0xFFFFu, 1u);
      value = max_value;
    }
    const uint32_t start = (uint32_t)__ldg(cdf + value);
    const uint32_t freq = (uint32_t)__ldg(cdf + value + 1) - start;
    e.put(start, freq);
  }
  e.base[--e.pos] = (uint16_t)(e.x & 0xFFFFu);
  e.base[--e.pos] = (uint16_t)(e.x >> 16);
  sizes[s] = scratch_stride - e.pos;  // in 16-bit words; the stream is scratch[s][pos .. stride)
}

// one warp per stream: copy its words to the compacted payload
__global__ void rans_compact_kernel(const uint16_t* __restrict__ scratch, int scratch_stride,
                                    const int32_t* __restrict__ sizes, const int64_t* __restrict__ offsets,
                                    int n_streams, uint16_t* __restrict__ out) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= n_streams) return;
  const int n = sizes[s];
  const uint16_t* src = scratch + (int64_t)s * scratch_stride + (scratch_stride - n);
  uint16_t* dst = out + offsets[s];
  for (int k = lane; k < n; k += 32) dst[k] = src[k];
}

// A stream never reads past its own words: a truncated / corrupt container yields zeros instead of an
// out-of-bounds access, and is reported through `status` (a well-formed stream ends exactly at its last word with
// the state back at the encoder's initial value).
__global__ void __launch_bounds__(128)
rans_decode_kernel(const uint16_t* __restrict__ payload, const int64_t* __restrict__ offsets,
                   const int32_t* __restrict__ sizes, const int32_t* __restrict__ indexes, RansTables t,
                   int64_t n_symbols, int stream_len, int32_t* __restrict__ symbols, int n_streams,
                   int32_t* __restrict__ status) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_streams) return;
  const int64_t lo = (int64_t)s * stream_len;
  const int64_t hi = lo + stream_len < n_symbols ? lo + stream_len : n_symbols;
  const uint16_t* p = payload + offsets[s];
  const uint16_t* const end = p + sizes[s];
  bool overrun = false;
  auto next = [&]() -> uint32_t {
    if (p < end) return *p++;
    overrun = true;
    return 0u;
  };
  uint32_t x = next() << 16;
  x |= next();
  auto advance = [&](uint32_t start, uint32_t freq) {
    x = freq * (x >> 16) + (x & 0xFFFFu) - start;
    if (x < kRansL) x = (x << 16) | next();
  };
  for (int64_t i = lo; i < hi; ++i) {
    const int row = __ldg(indexes + i);
    const int32_t* cdf = t.cdf + (int64_t)row * t.stride;
    const int max_value = __ldg(t.cdf_len + row) - 2;
    const uint32_t slot = x & 0xFFFFu;
    // largest v in [0, max_value] with cdf[v] <= slot   (cdf[max_value + 1] = 65536 > slot)
    int a = 0, b = max_value + 1;
    while (b - a > 1) {
      const int m = (a + b) >> 1;
      if ((uint32_t)__ldg(cdf + m) <= slot) a = m; else b = m;
    }
    const uint32_t start = (uint32_t)__ldg(cdf + a);
    advance(start, (uint32_t)__ldg(cdf + a + 1) - start);
    int value = a;
    if (a == max_value) {
      const uint32_t lo16 = x & 0xFFFFu;
      advance(lo16, 1u);
      const uint32_t hi16 = x & 0xFFFFu;
      advance(hi16, 1u);
      const uint32_t raw = (hi16 << 16) | lo16;
      value = (raw & 1u) ? -(int)((raw + 1u) >> 1) : max_value + (int)(raw >> 1);
    }
    symbols[i] = value + __ldg(t.offset + row);
  }
  if (status != nullptr && (overrun || p != end || x != kRansL)) atomicOr(status, 1);
}

}  // namespace b200vc

using namespace b200vc;

extern "C" int b200vc_rans_scratch_words(int stream_len) {
  // worst case per symbol: one renormalisation word for the symbol + two for an escape; + 2 state words
  return stream_len > 0 ? 3 * stream_len + 2 : 0;
}

extern "C" int b200vc_rans_encode(const int32_t* symbols, const int32_t* indexes, const int32_t* cdf,
                                  const int32_t* cdf_len, const int32_t* offset, int cdf_stride, int64_t n_symbols,
                                  int stream_len, uint16_t* scratch, int32_t* sizes_words, void* stream) {
  B200VC_REQUIRE(symbols && indexes && cdf && cdf_len && offset && scratch && sizes_words, "rans_encode: null pointer");
  B200VC_REQUIRE(n_symbols > 0 && stream_len > 0 && cdf_stride > 2, "rans_encode: bad size");
  const int64_t n_streams = (n_symbols + stream_len - 1) / stream_len;
  B200VC_REQUIRE(n_streams < (1ll << 31), "rans_encode: too many streams");
  RansTables t{cdf, cdf_len, offset, cdf_stride};
  rans_encode_kernel<<<(unsigned)((n_streams + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      symbols, indexes, t, n_symbols, stream_len, scratch, b200vc_rans_scratch_words(stream_len), sizes_words,
      (int)n_streams);
  return check_launch("rans_encode");
}

extern "C" int b200vc_rans_compact(const uint16_t* scratch, int stream_len, const int32_t* sizes_words,
                                   const int64_t* offsets_words, int n_streams, uint16_t* out, void* stream) {
  B200VC_REQUIRE(scratch && sizes_words && offsets_words && out && n_streams > 0 && stream_len > 0,
                 "rans_compact: bad argument");
  const int64_t threads = (int64_t)n_streams * 32;
  rans_compact_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      scratch, b200vc_rans_scratch_words(stream_len), sizes_words, offsets_words, n_streams, out);
  return check_launch("rans_compact");
}

extern "C" int b200vc_rans_decode(const uint16_t* payload, const int64_t* offsets_words, const int32_t* sizes_words,
                                  const int32_t* indexes, const int32_t* cdf, const int32_t* cdf_len,
                                  const int32_t* offset, int cdf_stride, int64_t n_symbols, int stream_len,
                                  int32_t* symbols_out, int32_t* status, void* stream) {
  B200VC_REQUIRE(payload && offsets_words && sizes_words && indexes && cdf && cdf_len && offset && symbols_out,
                 "rans_decode: null pointer");
  B200VC_REQUIRE(n_symbols > 0 && stream_len > 0 && cdf_stride > 2, "rans_decode: bad size");
  const int64_t n_streams = (n_symbols + stream_len - 1) / stream_len;
  B200VC_REQUIRE(n_streams < (1ll << 31), "rans_decode: too many streams");
  RansTables t{cdf, cdf_len, offset, cdf_stride};
  rans_decode_kernel<<<(unsigned)((n_streams + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      payload, offsets_words, sizes_words, indexes, t, n_symbols, stream_len, symbols_out, (int)n_streams, status);
  return check_launch("rans_decode");
}
