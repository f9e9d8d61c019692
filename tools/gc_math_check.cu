// Exhaustive / randomised bit-identity check of the packed-fp32 likelihood arithmetic (csrc/gc_math.cuh) against the
// scalar routines it restates, plus an FFMA2-vs-FFMA issue-rate probe.  Test infrastructure, not product code.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/gc_math_check tools/gc_math_check.cu
//   tools/_bin/gc_math_check            (prints one JSON line; exit code 1 on any mismatch)
#include <cstdio>
#include <cstdlib>

#include "../video-compression_b200/csrc/gc_math.cuh"

using namespace b200vc;

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// every one of the 2^32 arguments: thread i checks 2i and 2i+1 (one packed evaluation), in both template forms
__global__ void erfc_all(unsigned long long* bad, unsigned long long* bad_nan, uint32_t* first_bad) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // 2^31 threads
  const uint32_t u0 = (uint32_t)(2 * i), u1 = u0 + 1;
  const float x0 = __uint_as_float(u0), x1 = __uint_as_float(u1);
  const float2 got = erfc2<true>(make_float2(x0, x1));
  const float w0 = erfcf(x0), w1 = erfcf(x1);
  const bool n0 = x0 != x0, n1 = x1 != x1;
  int b = 0, bn = 0;
  if (__float_as_uint(got.x) != __float_as_uint(w0)) { if (n0) ++bn; else { ++b; atomicMin(first_bad, u0); } }
  if (__float_as_uint(got.y) != __float_as_uint(w1)) { if (n1) ++bn; else { ++b; atomicMin(first_bad, u1); } }
  if (!(u0 >> 31)) {  // non-negative arguments: the form without the reflection
    const float2 g2 = erfc2<false>(make_float2(x0, x1));
    if (!n0 && __float_as_uint(g2.x) != __float_as_uint(w0)) { ++b; atomicMin(first_bad, u0); }
    if (!n1 && __float_as_uint(g2.y) != __float_as_uint(w1)) { ++b; atomicMin(first_bad, u1); }
  }
  if (fabsf(x0) <= 9.25f && fabsf(x1) <= 9.25f) {  // the forms without the far-tail guards
    const float2 g3 = erfc2<true, false>(make_float2(x0, x1));
    if (__float_as_uint(g3.x) != __float_as_uint(w0)) { ++b; atomicMin(first_bad, u0); }
    if (__float_as_uint(g3.y) != __float_as_uint(w1)) { ++b; atomicMin(first_bad, u1); }
    if (!(u0 >> 31)) {
      const float2 g4 = erfc2<false, false>(make_float2(x0, x1));
      if (__float_as_uint(g4.x) != __float_as_uint(w0)) { ++b; atomicMin(first_bad, u0); }
      if (__float_as_uint(g4.y) != __float_as_uint(w1)) { ++b; atomicMin(first_bad, u1); }
    }
  }
  if (b) atomicAdd(bad, (unsigned long long)b);
  if (bn) atomicAdd(bad_nan, (unsigned long long)bn);
}

// random in-range operands: a in {0} u +-[2^-60, 2^60], b in [2^-60, 2^60]
__device__ __forceinline__ float rnd_mag(uint32_t h) {
  const uint32_t e = 127 - 60 + (h >> 23) % 121;  // biased exponent 67..187
  return __uint_as_float((e << 23) | (h & 0x7fffffu));
}
__global__ void div_rand(unsigned long long* bad, uint32_t seed, int rounds) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  int b = 0;
  for (int k = 0; k < rounds; ++k) {
    const uint32_t h0 = mix(tid * 4u + 0x9e3779b9u * (k + 1) + seed), h1 = mix(h0 ^ 0x85ebca6bu),
                   h2 = mix(h1 + 0xc2b2ae35u), h3 = mix(h2 ^ tid), h4 = mix(h3 + k), h5 = mix(h4 ^ seed);
    float2 num_a = make_float2(rnd_mag(h0), rnd_mag(h1)), num_b = make_float2(rnd_mag(h2), rnd_mag(h3));
    if (h4 & 1) num_a.x = -num_a.x;
    if (h4 & 2) num_a.y = -num_a.y;
    if (h4 & 4) num_b.x = -num_b.x;
    if (h4 & 8) num_b.y = -num_b.y;
    if ((h4 & 0xff0) == 0) num_a.x = 0.f;
    // half of the rounds: numerators close to the denominator's magnitude (quotients near 1, the K-GC case)
    float2 den = make_float2(rnd_mag(h5), rnd_mag(mix(h5)));
    if (k & 1) {
      num_a.x = __uint_as_float((__float_as_uint(den.x) & 0x7f800000u) | (h0 & 0x7fffffu));
      num_b.y = -__uint_as_float((__float_as_uint(den.y) & 0x7f800000u) | (h3 & 0x7fffffu));
    }
    float2 nb;
    const float2 r = div2_recip(den, &nb);
    const float2 qa = div2_apply(num_a, r, nb), qb = div2_apply(num_b, r, nb);
    b += __float_as_uint(qa.x) != __float_as_uint(__fdiv_rn(num_a.x, den.x));
    b += __float_as_uint(qa.y) != __float_as_uint(__fdiv_rn(num_a.y, den.y));
    b += __float_as_uint(qb.x) != __float_as_uint(__fdiv_rn(num_b.x, den.x));
    b += __float_as_uint(qb.y) != __float_as_uint(__fdiv_rn(num_b.y, den.y));
  }
  if (b) atomicAdd(bad, (unsigned long long)b);
}

// the K-GC operand shape: sigma in [0.11, 256], v = |integer + tiny|, numerators +-0.5 - v
__global__ void div_gc(unsigned long long* bad, uint32_t seed, int rounds) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  int b = 0;
  for (int k = 0; k < rounds; ++k) {
    const uint32_t h0 = mix(tid * 2u + 0x9e3779b9u * (k + 1) + seed), h1 = mix(h0 ^ 0x85ebca6bu), h2 = mix(h1 + k),
                   h3 = mix(h2 ^ seed);
    const float s0 = 0.11f + (h0 >> 8) * (1.f / 16777216.f) * ((h0 & 1) ? 255.89f : 3.f);
    const float s1 = 0.11f + (h1 >> 8) * (1.f / 16777216.f) * ((h1 & 1) ? 30.f : 0.7f);
    const float v0 = fabsf((float)((int)(h2 % 41) - 20) + ((h2 >> 8) & 3) * 5.9604645e-8f * (h2 >> 12 & 15));
    const float v1 = (h3 & 0x300) ? (float)(h3 % 7) : (h3 >> 10) * (1.f / 4194304.f);
    const float2 den = make_float2(s0, s1);
    const float2 nu = make_float2(0.5f - v0, 0.5f - v1), nl = make_float2(-0.5f - v0, -0.5f - v1);
    float2 nb;
    const float2 r = div2_recip(den, &nb);
    const float2 qa = div2_apply(nu, r, nb), qb = div2_apply(nl, r, nb);
    b += __float_as_uint(qa.x) != __float_as_uint(__fdiv_rn(nu.x, den.x));
    b += __float_as_uint(qa.y) != __float_as_uint(__fdiv_rn(nu.y, den.y));
    b += __float_as_uint(qb.x) != __float_as_uint(__fdiv_rn(nl.x, den.x));
    b += __float_as_uint(qb.y) != __float_as_uint(__fdiv_rn(nl.y, den.y));
  }
  if (b) atomicAdd(bad, (unsigned long long)b);
}

// issue-rate probe: 8 independent chains per thread, scalar FFMA vs packed FFMA2
template <bool kPacked>
__global__ void __launch_bounds__(256) fma_rate(float* out, int iters, float c0, float c1) {
  float2 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  const float2 a = f2(c0), b = f2(c1);
  for (int k = 0; k < iters; ++k) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (kPacked) {
        v[i] = __ffma2_rn(v[i], a, b);
      } else {
        v[i].x = __fmaf_rn(v[i].x, c0, c1);
        v[i].y = __fmaf_rn(v[i].y, c0, c1);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

int main() {
  unsigned long long *d_cnt, h_cnt[4] = {0, 0, 0, 0};
  uint32_t *d_first, h_first = 0xffffffffu;
  CK(cudaMalloc(&d_cnt, sizeof(h_cnt)));
  CK(cudaMalloc(&d_first, 4));
  CK(cudaMemcpy(d_cnt, h_cnt, sizeof(h_cnt), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_first, &h_first, 4, cudaMemcpyHostToDevice));
  erfc_all<<<(1u << 31) / 256, 256>>>(d_cnt + 0, d_cnt + 1, d_first);
  CK(cudaGetLastError());
  div_rand<<<148 * 64, 256>>>(d_cnt + 2, 12345u, 1024);  // 148*64*256*1024*4 = 9.9e9 quotients
  CK(cudaGetLastError());
  div_gc<<<148 * 64, 256>>>(d_cnt + 3, 777u, 1024);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&h_first, d_first, 4, cudaMemcpyDeviceToHost));

  // FFMA vs FFMA2 rate
  float* d_out;
  const int blocks = 148 * 8, iters = 4096;
  CK(cudaMalloc(&d_out, blocks * 256 * sizeof(float)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms[2] = {0, 0};
  for (int p = 0; p < 2; ++p) {
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      if (p) fma_rate<true><<<blocks, 256>>>(d_out, iters, 0.999f, 1e-3f);
      else fma_rate<false><<<blocks, 256>>>(d_out, iters, 0.999f, 1e-3f);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms[p], e0, e1));
    }
  }
  const double lane_fma = (double)blocks * 256 * iters * 16;  // fp32 FMAs per launch
  printf("{\"erfc2_mismatch\": %llu, \"erfc2_mismatch_nan_inputs\": %llu, \"erfc2_first_bad\": \"0x%08x\", "
         "\"div2_mismatch_random\": %llu, \"div2_mismatch_gc_shaped\": %llu, \"div2_quotients\": %.3g, "
         "\"ffma_tflops\": %.2f, \"ffma2_tflops\": %.2f}\n",
         h_cnt[0], h_cnt[1], h_first, h_cnt[2], h_cnt[3], 2.0 * 148 * 64 * 256 * 1024 * 4,
         2 * lane_fma / (ms[0] * 1e-3) / 1e12, 2 * lane_fma / (ms[1] * 1e-3) / 1e12);
  return (h_cnt[0] || h_cnt[2] || h_cnt[3]) ? 1 : 0;
}
