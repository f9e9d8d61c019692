// K-GDN / K-IGDN: generalised divisive normalisation.
//
// Replaces compressai.layers.GDN.forward (x**2 -> 1x1 conv with gamma, beta -> rsqrt|sqrt -> x*norm; four
// read+write sweeps in the reference) with ONE pass:   out_i = x_i * f(beta_i + sum_j gamma_ij x_j^2) [+ addend].
// Instances: LHBDC/model/layers.py:49-53,84-88,124-128,159-163 (inside ResidualBlockWithStride/Upsample).
//
// Layout: x [N, C, HW] (NCHW, positions contiguous).  Algorithmic bytes 2*C*4 per position (+C*4 with addend),
// 2*C*C flop per position.
//
// This file holds
//   * gdn_prepare: NonNegativeParametrizer applied once per weight version (+ transposed / tf32-split images)
//   * the CUDA-core fp32 kernel (exact fp32 FFMA; any C in {64,128,192}); the tcgen05 kernel is in gdn_tc.cu.
#include <atomic>

#include "common.cuh"

namespace b200vc {

int launch_gdn_tc(const float* x, const float* params, const float* addend, float* out, int N, int C, int64_t HW,
                  int inverse, cudaStream_t st);
int launch_gdn_tc192(const float* x, const float* params, const float* addend, float* out, int N, int64_t HW,
                     int inverse, cudaStream_t st);  // gdn_tc192.cu  // gdn_tc.cu

// params: [0,C) beta | [C, C+C^2) gamma[i][j] | [C+C^2, C+2C^2) gammaT[j][i] | [C+2C^2, C+4C^2) tf32 hi[i][j], lo[i][j]
__host__ __device__ inline int64_t gdn_off_gamma(int C) { return C; }
__host__ __device__ inline int64_t gdn_off_gammaT(int C) { return (int64_t)C + (int64_t)C * C; }
__host__ __device__ inline int64_t gdn_off_tc(int C) { return (int64_t)C + 2 * (int64_t)C * C; }

__global__ void gdn_prepare_kernel(const float* __restrict__ beta, const float* __restrict__ gamma,
                                   float beta_bound, float gamma_bound, float pedestal,
                                   float* __restrict__ params, int C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < C) {
    const float b = fmaxf(beta[idx], beta_bound);
    params[idx] = __fsub_rn(__fmul_rn(b, b), pedestal);
  }
  if (idx < C * C) {
    const int i = idx / C, j = idx % C;
    const float g0 = fmaxf(gamma[idx], gamma_bound);
    const float g = __fsub_rn(__fmul_rn(g0, g0), pedestal);
    params[gdn_off_gamma(C) + idx] = g;
    params[gdn_off_gammaT(C) + (int64_t)j * C + i] = g;
    // tf32 split: hi keeps the top 10 mantissa bits (round-to-nearest-even on bit 13), lo = g - hi (exact)
    uint32_t u = __float_as_uint(g);
    uint32_t r = u + 0xFFFu + ((u >> 13) & 1u);
    r &= 0xFFFFE000u;
    const float hi = __uint_as_float(r);
    const float lo = __fsub_rn(g, hi);
    float* tc = params + gdn_off_tc(C);  // row-major hi[i][j] then lo[i][j]: the tcgen05 kernel's TMEM image
    tc[idx] = hi;
    tc[(int64_t)C * C + idx] = lo;
  }
}

// ------------------------------------------------------------------------------ CUDA-core fp32 kernel
// CTA = 256 threads = 16 (channel groups) x 16 (position groups); tile = C channels x TP positions.
// Thread (ti, tp) owns channels {ti*4 + 64*r + 0..3} and positions {tp*4 + 64*g + 0..3}.
// smem: gT[C][C] (gamma transposed, loaded once per persistent CTA) + xs[C][TP] (current x tile).
template <int C, int TP>
__global__ void __launch_bounds__(256, 1)
gdn_fp32_kernel(const float* __restrict__ x, const float* __restrict__ params, const float* addend,
                float* out, int64_t HW, int tiles_per_sample, int total_tiles, int inverse) {
  constexpr int RG = C / 64;   // channel groups of 4 per thread
  constexpr int PG = TP / 64;  // position groups of 4 per thread
  extern __shared__ __align__(16) float smem[];
  float* gT = smem;            // [C][C]
  float* xs = smem + C * C;    // [C][TP]
  const int tid = threadIdx.x;
  const int tp = tid & 15, ti = tid >> 4;

  const float4* gsrc = reinterpret_cast<const float4*>(params + gdn_off_gammaT(C));
  for (int k = tid; k < C * C / 4; k += 256) reinterpret_cast<float4*>(gT)[k] = __ldg(gsrc + k);

  float beta[RG * 4];
#pragma unroll
  for (int r = 0; r < RG; ++r)
#pragma unroll
    for (int e = 0; e < 4; ++e) beta[r * 4 + e] = __ldg(params + ti * 4 + 64 * r + e);

  const bool vec_ok = (HW % 4 == 0);
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int n = tile / tiles_per_sample;
    const int64_t p0 = (int64_t)(tile % tiles_per_sample) * TP;
    const float* xn = x + (int64_t)n * C * HW;
    __syncthreads();  // previous tile fully consumed (also orders the gT fill on the first iteration)
    if (vec_ok) {
      for (int k = tid; k < C * TP / 4; k += 256) {
        const int j = k / (TP / 4), q = k % (TP / 4);
        const int64_t p = p0 + q * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < HW) v = ld_stream4(xn + (int64_t)j * HW + p);
        reinterpret_cast<float4*>(xs)[k] = v;
      }
    } else {
      for (int k = tid; k < C * TP; k += 256) {
        const int j = k / TP, q = k % TP;
        const int64_t p = p0 + q;
        xs[k] = (p < HW) ? __ldg(xn + (int64_t)j * HW + p) : 0.f;
      }
    }
    __syncthreads();

    float acc[RG * 4][PG * 4];
#pragma unroll
    for (int r = 0; r < RG * 4; ++r)
#pragma unroll
      for (int q = 0; q < PG * 4; ++q) acc[r][q] = 0.f;

#pragma unroll 4
    for (int j = 0; j < C; ++j) {
      float xv[PG * 4], gv[RG * 4];
#pragma unroll
      for (int g = 0; g < PG; ++g) {
        const float4 t = *reinterpret_cast<const float4*>(xs + j * TP + tp * 4 + 64 * g);
        xv[g * 4 + 0] = __fmul_rn(t.x, t.x);
        xv[g * 4 + 1] = __fmul_rn(t.y, t.y);
        xv[g * 4 + 2] = __fmul_rn(t.z, t.z);
        xv[g * 4 + 3] = __fmul_rn(t.w, t.w);
      }
#pragma unroll
      for (int r = 0; r < RG; ++r) {
        const float4 t = *reinterpret_cast<const float4*>(gT + j * C + ti * 4 + 64 * r);
        gv[r * 4 + 0] = t.x; gv[r * 4 + 1] = t.y; gv[r * 4 + 2] = t.z; gv[r * 4 + 3] = t.w;
      }
#pragma unroll
      for (int r = 0; r < RG * 4; ++r)
#pragma unroll
        for (int q = 0; q < PG * 4; ++q) acc[r][q] = fmaf(gv[r], xv[q], acc[r][q]);
    }

    // epilogue: norm = acc + beta; rsqrt | sqrt; * x; (+ addend); store
    float* on = out + (int64_t)n * C * HW;
    const float* an = addend ? addend + (int64_t)n * C * HW : nullptr;
#pragma unroll
    for (int r = 0; r < RG * 4; ++r) {
      const int i = ti * 4 + 64 * (r / 4) + (r % 4);
#pragma unroll
      for (int g = 0; g < PG; ++g) {
        const int pl = tp * 4 + 64 * g;
        const int64_t p = p0 + pl;
        const float4 xv = *reinterpret_cast<const float4*>(xs + i * TP + pl);
        float o[4];
        const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float nrm = __fadd_rn(acc[r][g * 4 + e], beta[r]);
          const float f = inverse ? sqrtf(nrm) : rsqrtf(nrm);
          o[e] = inverse == 2 ? nrm : __fmul_rn(xe[e], f);
        }
        if (vec_ok) {
          if (p < HW) {
            if (an) {
              const float4 a4 = *reinterpret_cast<const float4*>(an + (int64_t)i * HW + p);  // may alias `out`
              o[0] = __fadd_rn(o[0], a4.x); o[1] = __fadd_rn(o[1], a4.y);
              o[2] = __fadd_rn(o[2], a4.z); o[3] = __fadd_rn(o[3], a4.w);
            }
            st_stream4(on + (int64_t)i * HW + p, make_float4(o[0], o[1], o[2], o[3]));
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (p + e < HW) {
              float v = o[e];
              if (an) v = __fadd_rn(v, an[(int64_t)i * HW + p + e]);
              on[(int64_t)i * HW + p + e] = v;
            }
        }
      }
    }
  }
}

template <int C, int TP>
static int launch_gdn_fp32(const float* x, const float* params, const float* addend, float* out, int N, int64_t HW,
                           int inverse, cudaStream_t st) {
  const size_t smem = (size_t)(C * C + C * TP) * sizeof(float);
  static std::atomic<bool> configured[64];  // zero-initialised; idempotent set-up, safe under concurrent hosts
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gdn_fp32_kernel<C, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) {
      set_error("gdn_f32: cudaFuncSetAttribute(%zu B smem) failed: %s", smem, cudaGetErrorString(e));
      return B200VC_ECUDA;
    }
    configured[dev] = true;
  }
  const int64_t tps = (HW + TP - 1) / TP;
  const int64_t total = tps * N;
  if (total >= (1ll << 31)) {
    set_error("gdn_f32: too many tiles");
    return B200VC_EINVAL;
  }
  const int grid = (int)(total < sm_count() ? total : sm_count());
  gdn_fp32_kernel<C, TP><<<grid, 256, smem, st>>>(x, params, addend, out, HW, (int)tps, (int)total, inverse);
  return check_launch("gdn_f32");
}

}  // namespace b200vc

using namespace b200vc;

extern "C" int64_t b200vc_gdn_params_floats(int C) { return C > 0 ? (int64_t)C + 4 * (int64_t)C * C : 0; }

extern "C" int b200vc_gdn_prepare_f32(const float* beta, const float* gamma, float beta_bound, float gamma_bound,
                                      float pedestal, float* params_out, int C, void* stream) {
  B200VC_REQUIRE(beta && gamma && params_out, "gdn_prepare_f32: null pointer");
  B200VC_REQUIRE(C > 0 && C % 32 == 0 && C <= 1024, "gdn_prepare_f32: C=%d must be a multiple of 32", C);
  const int n = C * C;
  gdn_prepare_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(beta, gamma, beta_bound, gamma_bound,
                                                                       pedestal, params_out, C);
  return check_launch("gdn_prepare_f32");
}

extern "C" int b200vc_gdn_f32(const float* x, const float* params, const float* addend, float* out, int N, int C,
                              int64_t HW, int inverse, int impl, void* stream) {
  B200VC_REQUIRE(x && params && out, "gdn_f32: null pointer");
  B200VC_REQUIRE(N > 0 && HW > 0, "gdn_f32: bad shape N=%d HW=%lld", N, (long long)HW);
  B200VC_REQUIRE(impl >= 0 && impl <= 2, "gdn_f32: unknown impl %d", impl);
  B200VC_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0 &&
                     (reinterpret_cast<uintptr_t>(params) & 15u) == 0 &&
                     (addend == nullptr || (reinterpret_cast<uintptr_t>(addend) & 15u) == 0),
                 "gdn_f32: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == 2 || (impl == 0 && (C == 128 || C == 192) && HW % 4 == 0)) {
    const int rc = C == 192 ? launch_gdn_tc192(x, params, addend, out, N, HW, inverse, st)
                            : launch_gdn_tc(x, params, addend, out, N, C, HW, inverse, st);
    if (rc != B200VC_EUNSUPPORTED || impl == 2) return rc;
  }
  switch (C) {
    case 64: return launch_gdn_fp32<64, 128>(x, params, addend, out, N, HW, inverse, st);
    case 128: return launch_gdn_fp32<128, 128>(x, params, addend, out, N, HW, inverse, st);
    case 192: return launch_gdn_fp32<192, 64>(x, params, addend, out, N, HW, inverse, st);
    default:
      set_error("gdn_f32: unsupported channel count C=%d (supported: 64, 128, 192)", C);
      return B200VC_EUNSUPPORTED;
  }
}
