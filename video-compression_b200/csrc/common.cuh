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

}  // namespace b200vc
