// Dispatch / pipe-sharing probe for sm_100a: how many cycles per SM sub-partition do mixes of FFMA2, FFMA, ALU and
// MUFU instructions take?  (Question behind it: does a packed FFMA2 cost one issue slot or two?)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float mufu_rcp(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// MODE: 0 FFMA2 x8 | 1 FFMA x8 | 2 FFMA2 x8 + LOP3 x8 | 3 FFMA x8 + LOP3 x8 | 4 FFMA2 x8 + FFMA x8
//       5 FFMA2 x8 + MUFU x2 | 6 FFMA x8 + MUFU x2 | 7 MUFU x2 | 8 LOP3 x8 | 9 FFMA2 x8 + LOP3 x8 + MUFU x2
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, long long* cyc, int iters, float c0, float c1, unsigned m) {
  float2 v[8]; float s[8]; unsigned u[8]; float w[2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); s[i] = i + threadIdx.x; u[i] = threadIdx.x * 7 + i; }
  w[0] = 1.5f + threadIdx.x; w[1] = 2.5f + threadIdx.x;
  const float2 a = make_float2(c0, c0), b = make_float2(c1, c1);
  __syncthreads();
  const long long t0 = clock64();
  for (int k = 0; k < iters; ++k) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || MODE == 2 || MODE == 4 || MODE == 5 || MODE == 9) v[i] = __ffma2_rn(v[i], a, b);
      if (MODE == 1 || MODE == 3 || MODE == 4 || MODE == 6) s[i] = __fmaf_rn(s[i], c0, c1);
      if (MODE == 2 || MODE == 3 || MODE == 8 || MODE == 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(m), "r"(u[(i + 1) & 7]));
    }
    if (MODE == 5 || MODE == 6 || MODE == 7 || MODE == 9) { w[0] = mufu_rcp(w[0]); w[1] = mufu_rcp(w[1]); }
  }
  const long long t1 = clock64();
  float acc = w[0] + w[1];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += v[i].x + v[i].y + s[i] + __uint_as_float(u[i] & 0x3fffffffu);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, long long* cyc, int warps_per_smsp) {
  const int iters = 2048, blocks = 148 * warps_per_smsp / 2;  // 256 threads = 8 warps = 2 per SMSP
  probe<MODE><<<blocks, 256>>>(out, cyc, iters, 0.999f, 1e-3f, 0x5a5a5a5au);
  cudaDeviceSynchronize();
  long long h[148 * 8];
  cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < blocks; ++i) mean += h[i];
  mean /= blocks;
  printf("%-36s warps/SMSP %2d: %.2f cycles per loop body per SMSP\n", name, warps_per_smsp, mean / iters / warps_per_smsp);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&cyc, 148 * 8 * 8);
  for (int w : {4, 8}) {
    run<0>("FFMA2 x8", out, cyc, w);
    run<1>("FFMA x8", out, cyc, w);
    run<8>("LOP3 x8", out, cyc, w);
    run<7>("MUFU x2", out, cyc, w);
    run<2>("FFMA2 x8 + LOP3 x8", out, cyc, w);
    run<3>("FFMA x8 + LOP3 x8", out, cyc, w);
    run<4>("FFMA2 x8 + FFMA x8", out, cyc, w);
    run<5>("FFMA2 x8 + MUFU x2", out, cyc, w);
    run<6>("FFMA x8 + MUFU x2", out, cyc, w);
    run<9>("FFMA2 x8 + LOP3 x8 + MUFU x2", out, cyc, w);
  }
  return 0;
}
