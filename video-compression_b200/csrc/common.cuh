// Shared helpers for the b200vc kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200vc.h"

namespace b200vc {

void set_error(const char* fmt, ...);

#define B200VC_REQUIRE(cond, ...)        \
  do {                                   \
    if (!(cond)) {                       \
      ::b200vc::set_error(__VA_ARGS__);  \
      return B200VC_EINVAL;              \
    }                                    \
  } while (0)

// Check the launch (never synchronises).
int check_launch(const char* what);

int sm_count();

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of per-thread fp32 partials, finished in fp64, result valid in thread 0.
// Fixed shuffle tree + fixed warp order => deterministic for a given launch shape.
template <int kThreads>
__device__ __forceinline__ double block_sum_to_f64(float v) {
  __shared__ double s_part[kThreads / 32];
  double d = warp_sum((double)v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s_part[wid] = d;
  __syncthreads();
  double tot = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) tot += s_part[i];
  }
  return tot;
}

// Optional fused finish of the per-CTA partials ("last CTA done"): every CTA publishes its fp64 partial, takes a
// ticket on the sample's counter, and the CTA that draws the last ticket re-reads ALL partials of the sample in the
// same strided order + tree as sum_partials_kernel (blend.cu) -- so the total is bit-identical to the two-launch
// form, independent of which CTA finishes last -- writes totals[sample] and re-arms the counter (0).  Counters must
// be zero on entry; the kernels leave them zero.  totals == nullptr keeps the two-launch form.
struct Finish {
  double* totals;
  int32_t* counters;
};

// The finishing CTA of a sample: re-reads ALL its partials in the strided order + tree of sum_partials_kernel
// (blend.cu), writes totals[sample] and re-arms the counter.  Called by every thread of the CTA.
template <int kThreads>
__device__ __forceinline__ void finish_sample(const double* __restrict__ sample_partials, int n_per, int sample,
                                              Finish f) {
  static_assert(kThreads == 256, "must mirror sum_partials_kernel's 256-thread reduction order");
  __threadfence();
  double v = 0.0;
  for (int k = threadIdx.x; k < n_per; k += kThreads) v += __ldcg(sample_partials + k);
  __shared__ double s_fin[kThreads / 32];
  v = warp_sum(v);
  __syncthreads();  // s_fin may still be read by a previous call
  if ((threadIdx.x & 31) == 0) s_fin[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) t += s_fin[i];
    f.totals[sample] = t;
    f.counters[sample] = 0;
  }
}

// Must be called by every thread of the CTA; `tot` is read from thread 0 only.
template <int kThreads>
__device__ __forceinline__ void publish_partial(double tot, double* __restrict__ sample_partials, int idx,
                                                int n_per, int sample, Finish f) {
  static_assert(kThreads == 256, "must mirror sum_partials_kernel's 256-thread reduction order");
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    sample_partials[idx] = tot;
    int last = 0;
    if (f.totals != nullptr) {
      __threadfence();  // the partial is visible device-wide before the ticket is
      last = atomicAdd(f.counters + sample, 1) == n_per - 1;
    }
    s_last = last;
  }
  if (f.totals == nullptr) return;  // uniform
  __syncthreads();
  if (!s_last) return;
  finish_sample<kThreads>(sample_partials, n_per, sample, f);
}

// ATen CUDA sigmoid for float: 1 / (1 + exp(-x)).
__device__ __forceinline__ float sigmoid_f(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

// Streaming 128-bit global accessors (data touched once: do not pollute L1).
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream4(int32_t* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// Ampere-style asynchronous copies global -> shared (LDGSTS), L2-only caching for 16-byte copies.
template <int kBytes>
__device__ __forceinline__ void cp_async(uint32_t smem_dst, const void* gmem_src) {
  static_assert(kBytes == 4 || kBytes == 16, "cp.async size");
  if constexpr (kBytes == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

}  // namespace b200vc
