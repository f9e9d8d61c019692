// Microbenchmark (diagnostics, not product code): cycles per tcgen05.mma kind::tf32 (M=128, K=8) on B200 for the
// operand configurations the GDN kernel could use.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu
// Numerical results are garbage (operands are zeros); only issue-to-completion time is measured.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// mode: 0 = TS (A in TMEM), B MN-major SW128_32B | 1 = TS, B K-major SW128 | 2 = SS, A K-major SW128, B K-major SW128
//       3 = SS, A K-major, B MN-major SW128_32B | 4 = TS kind::f16 (bf16), B K-major SW128 (K = 16 per instr)
template <int mode, int N, int nacc>
__global__ void __launch_bounds__(128, 1) bench(int tiles, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  uint32_t* z = reinterpret_cast<uint32_t*>(raw);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) z[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // warp-uniform operands (shfl result) + elect.sync: lets ptxas keep every UTCHMMA operand in uniform registers
  // instead of wrapping each MMA in an ELECT / R2UR.BROADCAST waterfall loop
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
  uint32_t elected = 0;
  if (threadIdx.x < 32)
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(elected));
  if (threadIdx.x < 32 && elected) {
    const bool ts = (mode == 0 || mode == 1 || mode == 4);
    const bool b_mn = (mode == 0 || mode == 3);
    const bool f16 = (mode == 4);
    const uint32_t idesc = (1u << 4) | ((f16 ? 1u : 2u) << 7) | ((f16 ? 1u : 2u) << 10) | ((b_mn ? 1u : 0u) << 16) |
                           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_smem = base;               // 64 KB
    const uint32_t b_smem = base + 64 * 1024;   // up to 128 KB
    const long long t0 = clock64();
#pragma unroll 1
    for (int t = 0; t < tiles; ++t) {
#pragma unroll
      for (int i = 0; i < 48; ++i) {
        const int g = i & 15;
        uint64_t bdesc;
        if (b_mn) bdesc = make_desc(b_smem + g * 1024, 16384, 512, 1);                 // 32-position atoms 16 KB apart
        else      bdesc = make_desc(b_smem + (g >> 2) * (N * 128) + (g & 3) * 32, 16, 1024, 2);
        const uint32_t d = tmem + 256 + (i % nacc) * N;   // nacc independent accumulators, round robin
        const uint32_t acc = (i >= nacc);
        if (ts) {
          const uint32_t a = tmem + 8 * g;
          if (f16)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
        } else {
          const uint64_t adesc = make_desc(a_smem + (g >> 2) * (128 * 128) + (g & 3) * 32, 16, 1024, 2);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int mode, int N, int nacc>
static void run(long long* d, const char* name) {
  if ((mode == 0 || mode == 3) && N > 128) return;   // needs N/32 atoms at a uniform stride within 128 KB
  if (nacc * N > 256) return;
  const int tiles = 64;
  long long h = 0;
  cudaFuncSetAttribute(bench<mode, N, nacc>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024 + 1024);
  bench<mode, N, nacc><<<148, 128, 201 * 1024 + 1024>>>(2, d);  // warm-up
  bench<mode, N, nacc><<<148, 128, 201 * 1024 + 1024>>>(tiles, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-42s N=%3d : CUDA error %s\n", name, N, cudaGetErrorString(e));
    exit(1);
  }
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / (48.0 * tiles);
  const double k = (mode == 4) ? 16 : 8;
  printf("%-42s N=%3d acc=%d : %7.1f clk / MMA  -> %6.0f MAC/clk/SM\n", name, N, nacc, per, 128.0 * N * k / per);
}
template <int mode>
static void run_mode(long long* d, const char* name) {
  run<mode, 32, 1>(d, name);  run<mode, 64, 1>(d, name);  run<mode, 64, 2>(d, name);
  run<mode, 128, 1>(d, name); run<mode, 128, 2>(d, name); run<mode, 256, 1>(d, name);
}
int main() {
  long long* d;
  cudaMalloc(&d, 8);
  run_mode<0>(d, "TS  B=MN-major SW128_32B (current)");
  run_mode<1>(d, "TS  B=K-major SW128");
  run_mode<2>(d, "SS  A,B K-major SW128");
  run_mode<3>(d, "SS  A K-major, B MN-major SW128_32B");
  run_mode<4>(d, "TS  kind::f16 bf16 B K-major (K=16/instr)");
  return 0;
}
