// K-DCN: modulated deformable convolution.
//
// Replaces torchvision.ops.deform_conv2d behind the reference's DeformConv2d modules:
//   ICIP2023/src/model/m.py:29-34        DeformConv2d(96|64|32, same, kernel_size=3, padding=1, groups=8)
//   ICIP2024/src/model/helpers.py:40,57  OffsetDiversity.fusion = DeformConv2d(2C, C, 3, padding=1, groups=16), C = 64/96/128
// torchvision materialises the deformable im2col matrix ([N, Cin*9, Ho*Wo] fp32: 9x the input, written and re-read)
// and runs one small GEMM per weight group.  The reference's layers are *grouped* with 4-16 channels per group, so the
// GEMMs are tiny (K = 36..144, 4..12 output rows) and the operator is a gather, not a matrix product: here the
// samples are consumed as they are produced -- one thread owns one output position of one weight group, keeps the
// group's CO output channels in registers and walks (offset group, kernel point, input channel).  Offsets, masks and
// outputs are coalesced over positions; taps are L1 gathers around the position, like K-WARP.
// HBM traffic = input + offsets + mask + output once (no column matrix).
//
// Sampling arithmetic = torchvision/csrc/ops/cuda/deform_conv2d_kernel.cu (bilinear_interpolate + mask); the sum over
// (channel, kernel point) is sequential per thread where torchvision uses a cuBLAS GEMM, so results agree to fp32
// summation-order noise (tests hold 2e-5 of the output magnitude), not bit for bit.
#include "common.cuh"

namespace b200vc {

struct DcnParams {
  int N, Cin, H, W, Cout, Ho, Wo;
  int kh, kw, sh, sw, ph, pw, dh, dw;
  int groups, og;        // weight groups, offset groups
  int cin_g, cout_g;     // channels per weight group
  int ch_og;             // input channels per offset group
  int chunks;            // cout_g / CO
};

constexpr int kDcnThreads = 128;

template <int CO>
__global__ void __launch_bounds__(kDcnThreads)
deform_conv2d_kernel(const float* __restrict__ in, const float* __restrict__ offset, const float* __restrict__ mask,
                     const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ out,
                     DcnParams p) {
  extern __shared__ __align__(16) float s_w[];  // [cin_g][K][CO] of this (weight group, output chunk): one vector load per sample
  const int K = p.kh * p.kw;
  const int wg = blockIdx.y / p.chunks, chunk = blockIdx.y % p.chunks;
  const int n = blockIdx.z;
  const int co0 = wg * p.cout_g + chunk * CO;  // first output channel of this thread block
  const int nw = CO * p.cin_g * K;
  for (int i = threadIdx.x; i < nw; i += kDcnThreads) {
    const int o = i / (p.cin_g * K), ck = i - o * (p.cin_g * K);  // global order [o][c][k]
    s_w[ck * CO + o] = __ldg(weight + (int64_t)co0 * p.cin_g * K + i);
  }
  __syncthreads();

  const int P = p.Ho * p.Wo;
  const int pos = blockIdx.x * kDcnThreads + threadIdx.x;
  if (pos >= P) return;
  const int yo = pos / p.Wo, xo = pos % p.Wo;
  const int HW = p.H * p.W;
  float acc[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) acc[o] = bias ? __ldg(bias + co0 + o) : 0.f;

  const int c_begin = wg * p.cin_g, c_end = c_begin + p.cin_g;
  const float* in_n = in + (int64_t)n * p.Cin * HW;
  const int ybase = yo * p.sh - p.ph, xbase = xo * p.sw - p.pw;
  // offset groups touched by this weight group (one, when groups is a multiple of the offset groups)
  for (int g = c_begin / p.ch_og; g <= (c_end - 1) / p.ch_og; ++g) {
    const int c_lo = max(c_begin, g * p.ch_og), c_hi = min(c_end, (g + 1) * p.ch_og);
    const float* off_g = offset + ((int64_t)n * p.og + g) * 2 * K * P + pos;
    const float* msk_g = mask ? mask + ((int64_t)n * p.og + g) * K * P + pos : nullptr;
    for (int k = 0; k < K; ++k) {
      const int i = k / p.kw, j = k - i * p.kw;
      const float oy = __ldg(off_g + (int64_t)(2 * k) * P), ox = __ldg(off_g + (int64_t)(2 * k + 1) * P);
      const float m = msk_g ? __ldg(msk_g + (int64_t)k * P) : 1.f;
      const float y = __fadd_rn((float)(ybase + i * p.dh), oy), x = __fadd_rn((float)(xbase + j * p.dw), ox);
      // bilinear_interpolate(): zero outside (-1, H) x (-1, W); corners outside the image contribute zero
      const bool inside = !(y <= -1.f || (float)p.H <= y || x <= -1.f || (float)p.W <= x);
      const float fy = floorf(y), fx = floorf(x);
      const int y0 = (int)fy, x0 = (int)fx, y1 = y0 + 1, x1 = x0 + 1;
      const float lh = __fsub_rn(y, fy), lw = __fsub_rn(x, fx);
      const float hh = __fsub_rn(1.f, lh), hw = __fsub_rn(1.f, lw);
      const bool vy0 = inside && y0 >= 0, vy1 = inside && y1 <= p.H - 1;
      const bool vx0 = x0 >= 0, vx1 = x1 <= p.W - 1;
      // invalid corners: weight 0 and a clamped (always readable) address
      const float w1 = (vy0 && vx0) ? __fmul_rn(hh, hw) : 0.f, w2 = (vy0 && vx1) ? __fmul_rn(hh, lw) : 0.f;
      const float w3 = (vy1 && vx0) ? __fmul_rn(lh, hw) : 0.f, w4 = (vy1 && vx1) ? __fmul_rn(lh, lw) : 0.f;
      const int yc0 = min(max(y0, 0), p.H - 1), yc1 = min(max(y1, 0), p.H - 1);
      const int xc0 = min(max(x0, 0), p.W - 1), xc1 = min(max(x1, 0), p.W - 1);
      const int o1 = yc0 * p.W + xc0, o2 = yc0 * p.W + xc1, o3 = yc1 * p.W + xc0, o4 = yc1 * p.W + xc1;
      const float* ws = s_w + ((c_lo - c_begin) * K + k) * CO;
#pragma unroll 4
      for (int c = c_lo; c < c_hi; ++c) {
        const float* pl = in_n + (int64_t)c * HW;
        const float v1 = __ldg(pl + o1), v2 = __ldg(pl + o2), v3 = __ldg(pl + o3), v4 = __ldg(pl + o4);
        float val = __fmul_rn(w1, v1);
        val = __fmaf_rn(w2, v2, val);
        val = __fmaf_rn(w3, v3, val);
        val = __fmaf_rn(w4, v4, val);
        val = __fmul_rn(m, val);
        float wv[CO];
        if (CO % 4 == 0) {
#pragma unroll
          for (int q = 0; q < CO / 4; ++q) {
            const float4 t = reinterpret_cast<const float4*>(ws)[q];
            wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
          }
        } else if (CO % 2 == 0) {
#pragma unroll
          for (int q = 0; q < CO / 2; ++q) {
            const float2 t = reinterpret_cast<const float2*>(ws)[q];
            wv[2 * q] = t.x; wv[2 * q + 1] = t.y;
          }
        } else {
#pragma unroll
          for (int o = 0; o < CO; ++o) wv[o] = ws[o];
        }
#pragma unroll
        for (int o = 0; o < CO; ++o) acc[o] = __fmaf_rn(wv[o], val, acc[o]);
        ws += K * CO;
      }
    }
  }
  float* op = out + ((int64_t)n * p.Cout + co0) * P + pos;
#pragma unroll
  for (int o = 0; o < CO; ++o) op[(int64_t)o * P] = acc[o];
}

// ---------------------------------------------------------------------------------------------------
// Group-channels-last path.  ncu on the kernel above (ICIP2024 fusion, 544 x 960): l1tex data-pipe wavefronts 99 % of
// peak -- 288 scalar tap loads per position and group, 3.5 wavefronts each.  With the CIN_G channels of a weight
// group stored contiguously per pixel ([N, groups, H*W, CIN_G], one extra streaming pass), a tap is CIN_G/4 128-bit
// loads that bring all the group's channels at once: 8x fewer load instructions, ~3.5x fewer wavefronts.
template <int CIN_G>
__global__ void __launch_bounds__(256)
dcn_to_group_last_kernel(const float* __restrict__ in, float* __restrict__ ws, int HW, int groups) {
  const int pos = blockIdx.x * 256 + threadIdx.x;
  const int g = blockIdx.y, n = blockIdx.z;
  if (pos >= HW) return;
  const float* src = in + ((int64_t)n * groups + g) * CIN_G * HW + pos;
  float4* dst = reinterpret_cast<float4*>(ws + (((int64_t)n * groups + g) * HW + pos) * CIN_G);
#pragma unroll
  for (int q = 0; q < CIN_G / 4; ++q)
    dst[q] = make_float4(__ldg(src + (int64_t)(4 * q) * HW), __ldg(src + (int64_t)(4 * q + 1) * HW),
                         __ldg(src + (int64_t)(4 * q + 2) * HW), __ldg(src + (int64_t)(4 * q + 3) * HW));
}

// T threads share one output position, each taking CIN_G / T channels; partial sums meet in a butterfly shuffle and
// each of the T lanes stores CO / T of the outputs.  T = 1 is what ships (see launch_dcn_gl).
template <int CO, int CIN_G, int T>
__global__ void __launch_bounds__(kDcnThreads)
deform_conv2d_gl_kernel(const float* __restrict__ ws_in, const float* __restrict__ offset,
                        const float* __restrict__ mask, const float* __restrict__ weight,
                        const float* __restrict__ bias, float* __restrict__ out, DcnParams p) {
  constexpr int CPT = CIN_G / T;  // channels per thread
  static_assert(CPT % 4 == 0 && (T == 1 || T == 2 || T == 4) && CO % T == 0, "bad split");
  extern __shared__ __align__(16) float s_w[];  // [CIN_G][K][CO]
  const int K = p.kh * p.kw;
  const int wg = blockIdx.y, n = blockIdx.z;
  const int co0 = wg * CO;
  for (int i = threadIdx.x; i < CO * CIN_G * K; i += kDcnThreads) {
    const int o = i / (CIN_G * K), ck = i - o * (CIN_G * K);
    s_w[ck * CO + o] = __ldg(weight + (int64_t)co0 * CIN_G * K + i);
  }
  __syncthreads();
  const int P = p.Ho * p.Wo;
  const int gpos = (blockIdx.x * kDcnThreads + threadIdx.x) / T;
  const int sub = threadIdx.x % T;
  const bool live = gpos < P;
  const int pos = live ? gpos : P - 1;  // dead lanes still take part in the shuffles
  const int yo = pos / p.Wo, xo = pos % p.Wo;
  float acc[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) acc[o] = 0.f;
  const int g = (wg * CIN_G) / p.ch_og;  // the launcher guarantees the weight group lies inside one offset group
  const float4* in_g = reinterpret_cast<const float4*>(ws_in + ((int64_t)n * p.groups + wg) * (int64_t)p.H * p.W * CIN_G) +
                       sub * (CPT / 4);
  const float* off_g = offset + ((int64_t)n * p.og + g) * 2 * K * P + pos;
  const float* msk_g = mask ? mask + ((int64_t)n * p.og + g) * K * P + pos : nullptr;
  const int ybase = yo * p.sh - p.ph, xbase = xo * p.sw - p.pw;
  // All offsets and masks of the position are requested before the first one is used: one DRAM round trip per
  // thread instead of one per kernel point.
  constexpr int KP = 9;  // prefetch depth: the whole 3x3 kernel of the reference's layers
  float oyv[KP], oxv[KP], mv[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    const bool on = k < K;
    oyv[k] = on ? __ldg(off_g + (int64_t)(2 * k) * P) : 0.f;
    oxv[k] = on ? __ldg(off_g + (int64_t)(2 * k + 1) * P) : 0.f;
    mv[k] = (on && msk_g) ? __ldg(msk_g + (int64_t)k * P) : 1.f;
  }
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    if (k >= K) break;
    const int i = k / p.kw, j = k - i * p.kw;
    const float oy = oyv[k], ox = oxv[k], m = mv[k];
    const float y = __fadd_rn((float)(ybase + i * p.dh), oy), x = __fadd_rn((float)(xbase + j * p.dw), ox);
    const bool inside = !(y <= -1.f || (float)p.H <= y || x <= -1.f || (float)p.W <= x);
    const float fy = floorf(y), fx = floorf(x);
    const int y0 = (int)fy, x0 = (int)fx, y1 = y0 + 1, x1 = x0 + 1;
    const float lh = __fsub_rn(y, fy), lw = __fsub_rn(x, fx);
    const float hh = __fsub_rn(1.f, lh), hw = __fsub_rn(1.f, lw);
    const bool vy0 = inside && y0 >= 0, vy1 = inside && y1 <= p.H - 1;
    const bool vx0 = x0 >= 0, vx1 = x1 <= p.W - 1;
    const float w1 = (vy0 && vx0) ? __fmul_rn(hh, hw) : 0.f, w2 = (vy0 && vx1) ? __fmul_rn(hh, lw) : 0.f;
    const float w3 = (vy1 && vx0) ? __fmul_rn(lh, hw) : 0.f, w4 = (vy1 && vx1) ? __fmul_rn(lh, lw) : 0.f;
    const int yc0 = min(max(y0, 0), p.H - 1), yc1 = min(max(y1, 0), p.H - 1);
    const int xc0 = min(max(x0, 0), p.W - 1), xc1 = min(max(x1, 0), p.W - 1);
    const float4* t1 = in_g + (int64_t)(yc0 * p.W + xc0) * (CIN_G / 4);
    const float4* t2 = in_g + (int64_t)(yc0 * p.W + xc1) * (CIN_G / 4);
    const float4* t3 = in_g + (int64_t)(yc1 * p.W + xc0) * (CIN_G / 4);
    const float4* t4 = in_g + (int64_t)(yc1 * p.W + xc1) * (CIN_G / 4);
    const float* ws = s_w + ((sub * CPT) * K + k) * CO;
#pragma unroll
    for (int q = 0; q < CPT / 4; ++q) {
      const float4 a = __ldg(t1 + q), b = __ldg(t2 + q), c = __ldg(t3 + q), d = __ldg(t4 + q);
      const float va[4] = {a.x, a.y, a.z, a.w}, vb[4] = {b.x, b.y, b.z, b.w};
      const float vc[4] = {c.x, c.y, c.z, c.w}, vd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float val = __fmul_rn(w1, va[e]);
        val = __fmaf_rn(w2, vb[e], val);
        val = __fmaf_rn(w3, vc[e], val);
        val = __fmaf_rn(w4, vd[e], val);
        val = __fmul_rn(m, val);
        const float* wc = ws + (4 * q + e) * K * CO;
        float wv[CO];
        if (CO % 4 == 0) {
#pragma unroll
          for (int r = 0; r < CO / 4; ++r) {
            const float4 t = reinterpret_cast<const float4*>(wc)[r];
            wv[4 * r] = t.x; wv[4 * r + 1] = t.y; wv[4 * r + 2] = t.z; wv[4 * r + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int r = 0; r < CO / 2; ++r) {
            const float2 t = reinterpret_cast<const float2*>(wc)[r];
            wv[2 * r] = t.x; wv[2 * r + 1] = t.y;
          }
        }
#pragma unroll
        for (int o = 0; o < CO; ++o) acc[o] = __fmaf_rn(wv[o], val, acc[o]);
      }
    }
  }
  if (T > 1) {
#pragma unroll
    for (int o = 0; o < CO; ++o) {
      acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 1);
      if (T > 2) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 2);
    }
  }
  if (live) {
    float* op = out + ((int64_t)n * p.Cout + co0) * P + pos;
#pragma unroll
    for (int o = 0; o < CO; ++o) {
      if (o % T == sub) op[(int64_t)o * P] = __fadd_rn(acc[o], bias ? __ldg(bias + co0 + o) : 0.f);
    }
  }
}

template <int CO, int CIN_G>
static int launch_dcn_gl(const float* in, float* workspace, const float* offset, const float* mask, const float* weight,
                         const float* bias, float* out, const DcnParams& p, cudaStream_t st) {
  // T > 1 (one 16-byte channel quad per thread) was measured on B200 (tools/dcn_time.py): 1 406 -> 1 133 us on
  // white-noise offsets at 128->64 / 544x960, no change on smooth offsets, 275 -> 396 us at 256->128 / 136x240
  // (coordinate work is duplicated T times and the kernel is issue- as much as wavefront-limited) => not used.
  constexpr int T = 1;
  const int HW = p.H * p.W, P = p.Ho * p.Wo;
  dcn_to_group_last_kernel<CIN_G><<<dim3((HW + 255) / 256, p.groups, p.N), 256, 0, st>>>(in, workspace, HW, p.groups);
  const size_t smem = (size_t)CO * CIN_G * p.kh * p.kw * sizeof(float);
  const int64_t threads = (int64_t)P * T;
  deform_conv2d_gl_kernel<CO, CIN_G, T><<<dim3((unsigned)((threads + kDcnThreads - 1) / kDcnThreads), p.groups, p.N),
                                         kDcnThreads, smem, st>>>(workspace, offset, mask, weight, bias, out, p);
  return check_launch("deform_conv2d_f32(group-last)");
}

template <int CIN_G>
static int dispatch_dcn_gl(const float* in, float* workspace, const float* offset, const float* mask,
                           const float* weight, const float* bias, float* out, const DcnParams& p, cudaStream_t st) {
  switch (p.cout_g) {
    case 4: return launch_dcn_gl<4, CIN_G>(in, workspace, offset, mask, weight, bias, out, p, st);
    case 6: return launch_dcn_gl<6, CIN_G>(in, workspace, offset, mask, weight, bias, out, p, st);
    case 8: return launch_dcn_gl<8, CIN_G>(in, workspace, offset, mask, weight, bias, out, p, st);
    case 12: return launch_dcn_gl<12, CIN_G>(in, workspace, offset, mask, weight, bias, out, p, st);
    case 16: return launch_dcn_gl<16, CIN_G>(in, workspace, offset, mask, weight, bias, out, p, st);
    default: return B200VC_EUNSUPPORTED;
  }
}

template <int CO>
static int launch_dcn(const float* in, const float* offset, const float* mask, const float* weight, const float* bias,
                      float* out, DcnParams p, cudaStream_t st) {
  p.chunks = p.cout_g / CO;
  const int P = p.Ho * p.Wo;
  const size_t smem = (size_t)CO * p.cin_g * p.kh * p.kw * sizeof(float);
  if (smem > 48 * 1024) {
    set_error("deform_conv2d_f32: %zu B of weights per block exceed 48 KB", smem);
    return B200VC_EUNSUPPORTED;
  }
  dim3 grid((P + kDcnThreads - 1) / kDcnThreads, p.groups * p.chunks, p.N);
  deform_conv2d_kernel<CO><<<grid, kDcnThreads, smem, st>>>(in, offset, mask, weight, bias, out, p);
  return check_launch("deform_conv2d_f32");
}

}  // namespace b200vc

using namespace b200vc;

extern "C" int b200vc_deform_conv2d_f32(const float* input, const float* offset, const float* mask,
                                        const float* weight, const float* bias, float* out, float* workspace, int N,
                                        int Cin, int H,
                                        int W, int Cout, int kh, int kw, int stride_h, int stride_w, int pad_h,
                                        int pad_w, int dil_h, int dil_w, int groups, int offset_groups, void* stream) {
  B200VC_REQUIRE(input && offset && weight && out, "deform_conv2d_f32: null pointer");
  B200VC_REQUIRE(N > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "deform_conv2d_f32: bad shape");
  B200VC_REQUIRE(kh > 0 && kw > 0 && stride_h > 0 && stride_w > 0 && dil_h > 0 && dil_w > 0 && pad_h >= 0 && pad_w >= 0,
                 "deform_conv2d_f32: bad kernel geometry");
  B200VC_REQUIRE(groups > 0 && Cin % groups == 0 && Cout % groups == 0, "deform_conv2d_f32: groups=%d does not divide "
                 "Cin=%d / Cout=%d", groups, Cin, Cout);
  B200VC_REQUIRE(offset_groups > 0 && Cin % offset_groups == 0, "deform_conv2d_f32: offset_groups=%d does not divide "
                 "Cin=%d", offset_groups, Cin);
  DcnParams p;
  p.N = N; p.Cin = Cin; p.H = H; p.W = W; p.Cout = Cout;
  p.kh = kh; p.kw = kw; p.sh = stride_h; p.sw = stride_w; p.ph = pad_h; p.pw = pad_w; p.dh = dil_h; p.dw = dil_w;
  p.Ho = (H + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
  p.Wo = (W + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
  B200VC_REQUIRE(p.Ho > 0 && p.Wo > 0, "deform_conv2d_f32: empty output");
  B200VC_REQUIRE((int64_t)H * W < (1ll << 31) && (int64_t)p.Ho * p.Wo < (1ll << 31), "deform_conv2d_f32: plane too large");
  B200VC_REQUIRE(N <= 65535 && (int64_t)groups * (Cout / groups) <= 65535, "deform_conv2d_f32: grid too large");
  p.groups = groups; p.og = offset_groups;
  p.cin_g = Cin / groups; p.cout_g = Cout / groups; p.ch_og = Cin / offset_groups;
  p.chunks = 1;
  cudaStream_t st = (cudaStream_t)stream;
  // all of a group's output channels in one thread when their count is one of the instantiated sizes (the
  // reference's layers: 4, 6, 8, 12), otherwise in slices (the samples are then recomputed per slice)
  // group-channels-last path: needs the scratch copy, a weight group inside one offset group, 4 | cin_g <= 16,
  // 16-byte aligned scratch and a 3x3-sized (<= 48 KB) weight slice
  if (workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15u) == 0 &&
      p.ch_og % p.cin_g == 0 && kh * kw <= 9 &&
      (size_t)p.cout_g * p.cin_g * kh * kw * sizeof(float) <= 48 * 1024 && p.groups <= 65535) {
    int rc = B200VC_EUNSUPPORTED;
    if (p.cin_g == 4) rc = dispatch_dcn_gl<4>(input, workspace, offset, mask, weight, bias, out, p, st);
    else if (p.cin_g == 8) rc = dispatch_dcn_gl<8>(input, workspace, offset, mask, weight, bias, out, p, st);
    else if (p.cin_g == 12) rc = dispatch_dcn_gl<12>(input, workspace, offset, mask, weight, bias, out, p, st);
    else if (p.cin_g == 16) rc = dispatch_dcn_gl<16>(input, workspace, offset, mask, weight, bias, out, p, st);
    if (rc != B200VC_EUNSUPPORTED) return rc;
  }
  const int cg = p.cout_g;
  if (cg == 16) return launch_dcn<16>(input, offset, mask, weight, bias, out, p, st);
  if (cg == 12) return launch_dcn<12>(input, offset, mask, weight, bias, out, p, st);
  if (cg == 8) return launch_dcn<8>(input, offset, mask, weight, bias, out, p, st);
  if (cg == 6) return launch_dcn<6>(input, offset, mask, weight, bias, out, p, st);
  if (cg % 4 == 0) return launch_dcn<4>(input, offset, mask, weight, bias, out, p, st);
  if (cg % 3 == 0) return launch_dcn<3>(input, offset, mask, weight, bias, out, p, st);
  if (cg % 2 == 0) return launch_dcn<2>(input, offset, mask, weight, bias, out, p, st);
  return launch_dcn<1>(input, offset, mask, weight, bias, out, p, st);
}
