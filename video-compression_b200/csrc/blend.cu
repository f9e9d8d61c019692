// K-BLEND: mask blend + residual (+ optional SSE partials), and K-SSE: uint8-domain squared error.
// Replaces LHBDC/model/m.py:63-67, Flex-Rate.../b_model/b_model.py:68-73, ICIP2024/src/opt_helpers.py:35-45
// and the D2H + numpy PSNR of LHBDC/test/testing.py:176-182.
//
// Pure streaming (HBM-bound): 128-bit loads/stores, grid-stride over float4 units of one sample.
// Each arithmetic step is rounded where torch's separate elementwise kernels round (no FMA contraction).
#include "common.cuh"

namespace b200vc {

constexpr int kBlendThreads = 256;

__device__ __forceinline__ float blend_one(int mode, float m0, float m1, float a, float b) {
  if (mode == B200VC_BLEND_MASK) {
    // mask*fw + (1.0 - mask)*bw
    return __fadd_rn(__fmul_rn(m0, a), __fmul_rn(__fsub_rn(1.f, m0), b));
  } else if (mode == B200VC_BLEND_NORMW) {
    // w = 0.5*sigmoid(logit); (w1*xb + w2*xa)/(w1 + w2 + 1e-8)
    const float w1 = __fmul_rn(0.5f, sigmoid_f(m0)), w2 = __fmul_rn(0.5f, sigmoid_f(m1));
    const float num = __fadd_rn(__fmul_rn(w1, a), __fmul_rn(w2, b));
    const float den = __fadd_rn(__fadd_rn(w1, w2), 1e-8f);
    return __fdiv_rn(num, den);
  }
  // 0.5*w1 + (1-0.5)*w2
  return __fadd_rn(__fmul_rn(0.5f, a), __fmul_rn(0.5f, b));
}

// One sample per blockIdx.y.  P = H*W (plane), planes4 = P/4 when VEC == 4.
template <int VEC>
__global__ void __launch_bounds__(kBlendThreads)
blend_kernel(int mode, const float* __restrict__ mask, const float* __restrict__ a, int64_t a_bs,
             const float* __restrict__ b, int64_t b_bs, const float* __restrict__ x,
             float* __restrict__ pred, float* __restrict__ res, double* __restrict__ sse_partials, int64_t P,
             Finish fin) {
  const int n = blockIdx.y;
  const int mask_ch = (mode == B200VC_BLEND_NORMW) ? 2 : 1;
  const float* mp = mask ? mask + (int64_t)n * mask_ch * P : nullptr;
  const float* ap = a + (int64_t)n * a_bs;
  const float* bp = b + (int64_t)n * b_bs;
  const float* xp = x + (int64_t)n * 3 * P;
  float* pp = pred ? pred + (int64_t)n * 3 * P : nullptr;
  float* rp = res ? res + (int64_t)n * 3 * P : nullptr;
  float sse = 0.f;
  const int64_t units = P / VEC;
  for (int64_t i = (int64_t)blockIdx.x * kBlendThreads + threadIdx.x; i < units;
       i += (int64_t)gridDim.x * kBlendThreads) {
    const int64_t o = i * VEC;
    float m0[VEC], m1[VEC];
    if constexpr (VEC == 4) {
      float4 t = mp ? ld_stream4(mp + o) : make_float4(0, 0, 0, 0);
      m0[0] = t.x; m0[1] = t.y; m0[2] = t.z; m0[3] = t.w;
      if (mode == B200VC_BLEND_NORMW) t = ld_stream4(mp + P + o);
      m1[0] = t.x; m1[1] = t.y; m1[2] = t.z; m1[3] = t.w;
    } else {
      m0[0] = mp ? mp[o] : 0.f;
      m1[0] = (mode == B200VC_BLEND_NORMW) ? mp[P + o] : 0.f;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float av[VEC], bv[VEC], xv[VEC], pv[VEC], rv[VEC];
      if constexpr (VEC == 4) {
        const float4 ta = ld_stream4(ap + c * P + o), tb = ld_stream4(bp + c * P + o),
                     tx = ld_stream4(xp + c * P + o);
        av[0] = ta.x; av[1] = ta.y; av[2] = ta.z; av[3] = ta.w;
        bv[0] = tb.x; bv[1] = tb.y; bv[2] = tb.z; bv[3] = tb.w;
        xv[0] = tx.x; xv[1] = tx.y; xv[2] = tx.z; xv[3] = tx.w;
      } else {
        av[0] = ap[c * P + o]; bv[0] = bp[c * P + o]; xv[0] = xp[c * P + o];
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        pv[k] = blend_one(mode, m0[k], m1[k], av[k], bv[k]);
        rv[k] = __fsub_rn(xv[k], pv[k]);
        if (sse_partials) {
          const float d = __fsub_rn(fminf(fmaxf(pv[k], 0.f), 1.f), xv[k]);
          sse = __fmaf_rn(d, d, sse);
        }
      }
      if constexpr (VEC == 4) {
        if (pp) st_stream4(pp + c * P + o, make_float4(pv[0], pv[1], pv[2], pv[3]));
        if (rp) st_stream4(rp + c * P + o, make_float4(rv[0], rv[1], rv[2], rv[3]));
      } else {
        if (pp) pp[c * P + o] = pv[0];
        if (rp) rp[c * P + o] = rv[0];
      }
    }
  }
  if (sse_partials) {
    const double tot = block_sum_to_f64<kBlendThreads>(sse);
    publish_partial<kBlendThreads>(tot, sse_partials + (int64_t)n * gridDim.x, blockIdx.x, gridDim.x, n, fin);
  }
}

// uint8-domain SSE over the crop [:h,:w] of every plane: float_to_uint8 = round(clip(x,0,1)*255) (half-even).
// One CTA walks whole rows (no per-element div/mod), 128-bit loads when the row start is 16-byte aligned.
__device__ __forceinline__ float sq_u8_diff(float a, float b) {
  const float qa = rintf(__fmul_rn(fminf(fmaxf(a, 0.f), 1.f), 255.f));
  const float qb = rintf(__fmul_rn(fminf(fmaxf(b, 0.f), 1.f), 255.f));
  const float d = qa - qb;
  return d * d;  // exact integer <= 65025
}

__global__ void __launch_bounds__(kBlendThreads)
sse_u8_kernel(const float* __restrict__ a0, const float* __restrict__ b0, double* __restrict__ partials,
              int planes, int H, int W, int h, int w, Finish fin) {
  // blockIdx.y = sample: per-sample sums, so one launch scores all frames of a hierarchy level
  const float* a = a0 + (int64_t)blockIdx.y * planes * H * W;
  const float* b = b0 + (int64_t)blockIdx.y * planes * H * W;
  const int rows = planes * h;
  const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) == 0;
  const int w4 = vec ? (w / 4) : 0;
  double dacc = 0.0;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int pl = row / h, yy = row - pl * h;
    const int64_t o = ((int64_t)pl * H + yy) * W;
    float acc = 0.f;  // <= (4 + 1) elements per thread per row at 1080p: far below 2^24 / 65025 = 258 terms
    int cnt = 0;
    for (int x4 = threadIdx.x; x4 < w4; x4 += kBlendThreads) {
      const float4 va = ld_stream4(a + o + 4 * x4), vb = ld_stream4(b + o + 4 * x4);
      acc += sq_u8_diff(va.x, vb.x) + sq_u8_diff(va.y, vb.y) + sq_u8_diff(va.z, vb.z) + sq_u8_diff(va.w, vb.w);
      if ((cnt += 4) >= 252) { dacc += (double)acc; acc = 0.f; cnt = 0; }
    }
    for (int x = 4 * w4 + threadIdx.x; x < w; x += kBlendThreads) {
      acc += sq_u8_diff(__ldg(a + o + x), __ldg(b + o + x));
      if (++cnt >= 252) { dacc += (double)acc; acc = 0.f; cnt = 0; }
    }
    dacc += (double)acc;
  }
  __shared__ double s_part[kBlendThreads / 32];
  double d = warp_sum(dacc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = d;
  __syncthreads();
  double tot = 0.0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kBlendThreads / 32; ++i) tot += s_part[i];
  }
  publish_partial<kBlendThreads>(tot, partials + (int64_t)blockIdx.y * gridDim.x, blockIdx.x, gridDim.x, blockIdx.y, fin);
}

__global__ void sum_partials_kernel(const double* __restrict__ partials, int n_per, double* __restrict__ out) {
  // one CTA of 256 threads per output; fixed strided order then fixed tree => deterministic
  const double* p = partials + (int64_t)blockIdx.x * n_per;
  double v = 0.0;
  for (int k = threadIdx.x; k < n_per; k += 256) v += p[k];
  __shared__ double s_part[8];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += s_part[i];
    out[blockIdx.x] = tot;
  }
}

static bool aligned16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace b200vc

using namespace b200vc;

extern "C" int b200vc_blend_residual_f32(int mode, const float* mask, const float* a, int64_t a_bs,
                                         const float* b, int64_t b_bs, const float* x_cur, float* pred,
                                         float* res, double* sse_partials, int n_blocks, double* sse_totals,
                                         int32_t* counters, int N, int H, int W, void* stream) {
  B200VC_REQUIRE(!sse_totals || (sse_partials && counters), "blend_residual_f32: totals need partials and counters");
  B200VC_REQUIRE(a && b && x_cur, "blend_residual_f32: null pointer");
  B200VC_REQUIRE(mode >= 0 && mode <= 2, "blend_residual_f32: unknown mode %d", mode);
  B200VC_REQUIRE(mode == B200VC_BLEND_HALF || mask, "blend_residual_f32: mask required for mode %d", mode);
  B200VC_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && n_blocks > 0, "blend_residual_f32: bad shape");
  const int64_t P = (int64_t)H * W;
  const bool vec = (P % 4 == 0) && (a_bs % 4 == 0) && (b_bs % 4 == 0) && aligned16(mask) && aligned16(a) &&
                   aligned16(b) && aligned16(x_cur) && aligned16(pred) && aligned16(res);
  dim3 grid(n_blocks, N);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec)
    blend_kernel<4><<<grid, kBlendThreads, 0, st>>>(mode, mask, a, a_bs, b, b_bs, x_cur, pred, res,
                                                    sse_partials, P, Finish{sse_totals, counters});
  else
    blend_kernel<1><<<grid, kBlendThreads, 0, st>>>(mode, mask, a, a_bs, b, b_bs, x_cur, pred, res,
                                                    sse_partials, P, Finish{sse_totals, counters});
  return check_launch("blend_residual_f32");
}

extern "C" int b200vc_sse_u8_f32(const float* a, const float* b, double* partials, int n_blocks, double* totals,
                                 int32_t* counters, int N, int C, int H, int W, int h, int w, void* stream) {
  B200VC_REQUIRE(!totals || counters, "sse_u8_f32: totals need counters");
  B200VC_REQUIRE(a && b && partials, "sse_u8_f32: null pointer");
  B200VC_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && h > 0 && w > 0 && h <= H && w <= W && n_blocks > 0,
                 "sse_u8_f32: bad shape");
  B200VC_REQUIRE(N <= 65535, "sse_u8_f32: N too large");
  sse_u8_kernel<<<dim3(n_blocks, N), kBlendThreads, 0, (cudaStream_t)stream>>>(a, b, partials, C, H, W, h, w,
                                                                               Finish{totals, counters});
  return check_launch("sse_u8_f32");
}

extern "C" int b200vc_sum_partials_f64(const double* partials, int n_per, int n_out, double* out,
                                       void* stream) {
  B200VC_REQUIRE(partials && out && n_per > 0 && n_out > 0, "sum_partials_f64: bad argument");
  sum_partials_kernel<<<n_out, 256, 0, (cudaStream_t)stream>>>(partials, n_per, out);
  return check_launch("sum_partials_f64");
}
