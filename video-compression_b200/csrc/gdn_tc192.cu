// K-GDN / K-IGDN on the 5th-gen tensor cores for C = 192 (the width of the I-frame codec's transforms:
// compressai mbt2018_mean at quality >= 5 has N = 192, LHBDC/test/testing.py:78-86,209; and of the joint-autoregressive
// base class the ICIP codecs derive from).  Same mathematics and 3xTF32 scheme as gdn_tc.cu (C = 128):
//
//   norm[i, p] = beta_i + sum_j gamma[i, j] * x[j, p]^2 ,   out[i, p] = x[i, p] * (rsqrt | sqrt)(norm[i, p]) [+ addend]
//
// What is different at C = 192 -- gamma no longer fits one SM: hi + lo images of a 192 x 192 fp32 matrix are 295 KB,
// more than TMEM (256 KB, of which the accumulators need a part) or shared memory.  So the OUTPUT CHANNELS are split:
// a CTA owns 96 of the 192 output channels (h = blockIdx.x & 1) and the whole contraction length K = 192:
//   * A operand = gamma rows [96h, 96h+96) x 192, zero-padded to UMMA M = 128 lanes, hi in TMEM columns [0,192),
//     lo in [192,384), resident for the whole kernel;
//   * D = 128 lanes x 64 columns fp32, ONE accumulator per tile (ghi*hi + ghi*lo + glo*hi, 72 tcgen05.mma of
//     K = 8), double buffered in the remaining 128 TMEM columns;
//   * B operand = the x^2 tile, 192 channels x 64 positions, MN-major SWIZZLE_128B_BASE32B, hi / lo images (48 KB each);
//   * the two CTAs of a pair (blockIdx 2m, 2m+1) walk the same tiles, so the second read of x is an L2 hit: DRAM traffic
//     stays at the algorithmic 2 * 192 * 4 B per position;
//   * raw ring of two 48-KB tiles: a slot is free again as soon as the split warps and the epilogue warps (which take
//     their x row into registers early) have read it; results leave through a separate 24-KB staging tile
//     (96 channels x 64 positions) and a TMA store / reduce-add, so the pipeline depth does not depend on the store.
// Shared memory: 2 x 48 (raw) + 2 x 48 (hi, lo) + 24 (out) = 216 KB.  Tensor time per item: 72 MMAs x 32 clk = 2.3 k clk
// against ~2.1 k clk of HBM time per (tile, half) -- one quarter of the tensor work is the M padding.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "gdn_tc_common.cuh"

namespace b200vc {
namespace tc192 {

using namespace tc;

constexpr int kC = 192;            // channels = contraction length K
constexpr int kRows = 96;          // output channels per CTA
constexpr int kM = 128;            // UMMA M (rows 96..127 of the A operand are zero)
constexpr int kTileP = 64;         // positions per tile (UMMA N)
constexpr int kHalfBytes = kC * 32 * 4;          // one [192 x 32] fp32 box = 24 KB
constexpr int kTileBytes = 2 * kHalfBytes;       // 48 KB
constexpr int kOutHalfBytes = kRows * 32 * 4;    // one [96 x 32] box = 12 KB
constexpr int kOutBytes = 2 * kOutHalfBytes;     // 24 KB
constexpr int kKSteps = kC / 8;                  // 24 k-steps of K = 8 (TF32)
constexpr int kXfWarps = 8, kEpiWarps = 8;
constexpr int kFirstXf = 2, kFirstEpi = kFirstXf + kXfWarps;
constexpr int kThreads = (kFirstEpi + kEpiWarps) * 32;   // 576
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColGhi = 0, kColGlo = kC, kColD = 2 * kC, kColDStage = kTileP;   // 192 + 192 + 2 x 64 = 512
constexpr int kRaw = 2;
constexpr int kRawOff = 0, kHiOff = kRaw * kTileBytes, kLoOff = kHiOff + kTileBytes, kOutOff = kLoOff + kTileBytes;
constexpr int kBarOff = kOutOff + kOutBytes;
constexpr int kNumBars = 2 * kRaw + 4 + 4 + 1;
constexpr int kSmemBytes = kBarOff + 8 * kNumBars + 16 + 1024 /*alignment slack*/;
constexpr uint32_t kIdesc = idesc_tf32(kM, kTileP);

template <int INV>
__global__ void __launch_bounds__(kThreads, 1)
gdn_tc192_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_out,
                 const float* __restrict__ params, const float* __restrict__ addend, int64_t HW, int tiles_per_sample,
                 int total_tiles, int accumulate) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  auto raw_addr = [&](int r) { return smem_base + kRawOff + r * kTileBytes; };
  const uint32_t bar_base = smem_base + kBarOff;
  auto raw_full = [&](int r) { return bar_base + 8 * r; };
  auto raw_empty = [&](int r) { return bar_base + 8 * (kRaw + r); };
  // operand barriers per channel half (c = 0: channels 0-95 = k-steps 0-11, c = 1: the rest): the tensor core starts
  // after half the split, and the split of the next tile starts after half the MMAs
  auto ab_full = [&](int c) { return bar_base + 8 * (2 * kRaw + c); };
  auto ab_empty = [&](int c) { return bar_base + 8 * (2 * kRaw + 2 + c); };
  auto d_full = [&](int d) { return bar_base + 8 * (2 * kRaw + 4 + d); };
  auto d_empty = [&](int d) { return bar_base + 8 * (2 * kRaw + 6 + d); };
  const uint32_t gamma_ready = bar_base + 8 * (2 * kRaw + 8);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + kBarOff + 8 * kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x & 1;                      // which 96 output channels
  const int first_tile = blockIdx.x >> 1, tile_step = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int r = 0; r < kRaw; ++r) {
      mbar_init(raw_full(r), 1);
      mbar_init(raw_empty(r), kXfWarps + kEpiWarps);   // split warps + epilogue warps have read the slot
    }
    for (int c = 0; c < 2; ++c) {
      mbar_init(ab_full(c), kXfWarps);
      mbar_init(ab_empty(c), 1);
    }
    for (int d = 0; d < 2; ++d) {
      mbar_init(d_full(d), 1);
      mbar_init(d_empty(d), kEpiWarps);
    }
    mbar_init(gamma_ready, kXfWarps + kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_out)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(const_cast<uint32_t*>(tmem_slot))), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // gamma hi / lo rows [96h, 96h+96) -> TMEM lanes 0..95 (lanes 96..127: zeros), by the 16 split + epilogue warps:
  // warp w owns lane quarter (w & 3); the four warps of a quarter take {hi, lo} x {columns 0-95, 96-191}.
  if (warp >= kFirstXf) {
    const int q = warp & 3, grp = (warp - kFirstXf) >> 2;
    const int which = grp & 1, half = grp >> 1;
    const int i = 32 * q + lane;                     // local row == TMEM lane
    const bool real = i < kRows;
    const float* src_row = params + kC + 2 * kC * kC + (int64_t)which * kC * kC + (int64_t)(kRows * h + (real ? i : 0)) * kC;
#pragma unroll 1
    for (int part = 3 * half; part < 3 * half + 3; ++part) {
      uint32_t v[32];
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        float4 t = __ldg(reinterpret_cast<const float4*>(src_row + part * 32 + e));
        if (!real) t = make_float4(0.f, 0.f, 0.f, 0.f);
        v[e] = __float_as_uint(t.x); v[e + 1] = __float_as_uint(t.y);
        v[e + 2] = __float_as_uint(t.z); v[e + 3] = __float_as_uint(t.w);
      }
      tmem_st32(tmem + ((uint32_t)(32 * q) << 16) + (which ? kColGlo : kColGhi) + part * 32, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(gamma_ready);
  }

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int k = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++k) {
        const int r = k % kRaw, ph = (k / kRaw) & 1;
        const int row0 = (tile / tiles_per_sample) * kC;
        const int p0 = (tile % tiles_per_sample) * kTileP;
        mbar_wait(raw_empty(r), ph ^ 1);
        mbar_arrive_expect_tx(raw_full(r), kTileBytes);
        tma_load_2d(raw_addr(r), &map_x, p0, row0, raw_full(r));
        tma_load_2d(raw_addr(r) + kHalfBytes, &map_x, p0 + 32, row0, raw_full(r));
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, one elected lane)
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    mbar_wait(gamma_ready, 0);
    tc_fence_after();
    const uint64_t hi_desc = make_b_desc_mn_tf32(smem_base + kHiOff, kHalfBytes);
    const uint64_t lo_desc = make_b_desc_mn_tf32(smem_base + kLoOff, kHalfBytes);
    int k = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++k) {
      const int pa = k & 1, d = k & 1, pd = (k >> 1) & 1;
      mbar_wait(d_empty(d), pd ^ 1);
      const uint32_t dacc = tmem_u + kColD + kColDStage * d;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        mbar_wait(ab_full(c), pa);
        tc_fence_after();
        if (elect_one()) {
          // the small cross terms go first, the main term last (the accumulator truncates: small before large)
#pragma unroll
          for (int g = 12 * c; g < 12 * c + 12; ++g)   // D += ghi * lo
            umma_tf32_ts_desc(dacc, tmem_u + kColGhi + 8 * g, lo_desc + (uint64_t)(g * (1024 >> 4)), kIdesc,
                              (c != 0 || g != 0) ? 1u : 0u);
#pragma unroll
          for (int g = 12 * c; g < 12 * c + 12; ++g)   // D += glo * hi
            umma_tf32_ts_desc(dacc, tmem_u + kColGlo + 8 * g, hi_desc + (uint64_t)(g * (1024 >> 4)), kIdesc, 1u);
#pragma unroll
          for (int g = 12 * c; g < 12 * c + 12; ++g)   // D += ghi * hi
            umma_tf32_ts_desc(dacc, tmem_u + kColGhi + 8 * g, hi_desc + (uint64_t)(g * (1024 >> 4)), kIdesc, 1u);
          umma_commit(ab_empty(c));
          if (c == 1) umma_commit(d_full(d));
        }
        __syncwarp();
      }
    }
  } else if (warp < kFirstEpi) {
    // ------------------------------------------------------------------ square + TF32 hi/lo split (8 warps)
    const int t = threadIdx.x - kFirstXf * 32;
    // one channel half of both atoms = 2 x (96 rows x 8 float4) = 1536 float4 -> 6 per thread
    constexpr int kPerHalf = 2 * kRows * 8, kIters = kPerHalf / (kXfWarps * 32);
    float4* hi4 = reinterpret_cast<float4*>(smem_gen + kHiOff);
    float4* lo4 = reinterpret_cast<float4*>(smem_gen + kLoOff);
    int k = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++k) {
      const int r = k % kRaw, pr = (k / kRaw) & 1, pa = k & 1;
      mbar_wait(raw_full(r), pr);
      const float4* raw4 = reinterpret_cast<const float4*>(smem_gen + kRawOff + r * kTileBytes);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        mbar_wait(ab_empty(c), pa ^ 1);
        float4 v[kIters];
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
          const int idx = it * (kXfWarps * 32) + t;                    // 0..1535: (atom, row in half, chunk)
          const int atom = idx / (kRows * 8), rem = idx - atom * (kRows * 8);
          v[it] = raw4[atom * (kHalfBytes / 16) + c * (kRows * 8) + rem];
        }
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
          const int idx = it * (kXfWarps * 32) + t;
          const int atom = idx / (kRows * 8), rem = idx - atom * (kRows * 8);
          const int o = atom * (kHalfBytes / 16) + c * (kRows * 8) + rem;
          float4 sq, hh, ll;
          sq.x = __fmul_rn(v[it].x, v[it].x); sq.y = __fmul_rn(v[it].y, v[it].y);
          sq.z = __fmul_rn(v[it].z, v[it].z); sq.w = __fmul_rn(v[it].w, v[it].w);
          hh.x = to_tf32_rna(sq.x); hh.y = to_tf32_rna(sq.y); hh.z = to_tf32_rna(sq.z); hh.w = to_tf32_rna(sq.w);
          ll.x = __fsub_rn(sq.x, hh.x); ll.y = __fsub_rn(sq.y, hh.y);
          ll.z = __fsub_rn(sq.z, hh.z); ll.w = __fsub_rn(sq.w, hh.w);
          hi4[o] = hh;
          lo4[o] = ll;
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(ab_full(c));
          if (c == 1) mbar_arrive(raw_empty(r));   // this warp is done with the raw tile
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int hp = (warp - kFirstEpi) >> 2; // 32-position half of the tile
    const int i = 32 * q + lane;            // local output channel == TMEM lane
    const bool real = i < kRows;            // lane quarter 3 is the M padding: it only keeps the barriers balanced
    const int ch = kRows * h + (real ? i : 0);
    const float beta = __ldg(params + ch);
    const bool leader = (threadIdx.x == kFirstEpi * 32);
    const bool flip = (i >> 2) & 1;
    float4* out4 = reinterpret_cast<float4*>(smem_gen + kOutOff) + hp * (kOutHalfBytes / 16) + (real ? i : 0) * 8;
    int k = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++k) {
      const int r = k % kRaw, d = k & 1, pd = (k >> 1) & 1;
      const int row0 = (tile / tiles_per_sample) * kC + kRows * h;
      const int p0 = (tile % tiles_per_sample) * kTileP;
      // x row segment into registers, then the raw slot is released (the result leaves through the staging tile)
      const float4* raw4 = reinterpret_cast<const float4*>(smem_gen + kRawOff + r * kTileBytes) +
                           hp * (kHalfBytes / 16) + ch * 8;
      mbar_wait(raw_full(r), (k / kRaw) & 1);
      float4 xv[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int cc = c ^ (int)flip;
        xv[c] = raw4[(((cc >> 1) ^ (i & 3)) << 1) | (cc & 1)];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(raw_empty(r));
      mbar_wait(d_full(d), pd);
      tc_fence_after();
      const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + kColD + kColDStage * d + 32 * hp;
      const float* arow = (addend && real) ? addend + ((int64_t)row0 + i) * HW + p0 + 32 * hp : nullptr;
      const int64_t pbase = (int64_t)p0 + 32 * hp;
      // the previous tile's store must have finished reading the staging tile before it is overwritten
      if (leader && k > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      uint32_t v1[2][8];
      tmem_ld8_issue(taddr, v1[0]);
#pragma unroll
      for (int pc = 0; pc < 4; ++pc) {
        tmem_ld_wait();
        if (pc < 3) {
          tmem_ld8_issue(taddr + 8 * (pc + 1), v1[(pc + 1) & 1]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d_empty(d));  // accumulator drained: the MMA warp may start tile k+2
        }
        float nr[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) nr[j] = __fadd_rn(__uint_as_float(v1[pc & 1][j]), beta);
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const int c = 2 * pc + c2;
          const int cc = c ^ (int)flip;
          const float xe[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w};
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float n = flip ? nr[4 * (c2 ^ 1) + e] : nr[4 * c2 + e];
            o[e] = INV == 2 ? n : __fmul_rn(xe[e], INV ? sqrt_approx(n) : rsqrt_approx(n));
          }
          if (arow != nullptr && pbase + 4 * cc < HW) {
            const float4 a4 = *reinterpret_cast<const float4*>(arow + 4 * cc);
            o[0] = __fadd_rn(o[0], a4.x); o[1] = __fadd_rn(o[1], a4.y);
            o[2] = __fadd_rn(o[2], a4.z); o[3] = __fadd_rn(o[3], a4.w);
          }
          if (real) out4[(((cc >> 1) ^ (i & 3)) << 1) | (cc & 1)] = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
      fence_proxy_async();  // staging tile (generic writes) -> visible to the TMA store
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (leader) {
        const uint32_t src = smem_base + kOutOff;
        if (accumulate) {
          tma_reduce_add_2d(&map_out, src, p0, row0);
          if ((int64_t)p0 + 32 < HW) tma_reduce_add_2d(&map_out, src + kOutHalfBytes, p0 + 32, row0);
        } else {
          tma_store_2d(&map_out, src, p0, row0);
          if ((int64_t)p0 + 32 < HW) tma_store_2d(&map_out, src + kOutHalfBytes, p0 + 32, row0);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (leader && k > 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

template <int INV>
static int launch_inv(const CUtensorMap& map_x, const CUtensorMap& map_out, const float* params, const float* addend,
                      int64_t HW, int tps, int total, int accumulate, int grid, cudaStream_t st) {
  static std::once_flag once[64];
  static bool ok[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return B200VC_EUNSUPPORTED;
  std::call_once(once[dev], [&]() {
    ok[dev] = cudaFuncSetAttribute(gdn_tc192_kernel<INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) ==
              cudaSuccess;
    if (!ok[dev]) (void)cudaGetLastError();
  });
  if (!ok[dev]) {
    set_error("gdn_f32: cannot reserve %d B of shared memory", kSmemBytes);
    return B200VC_EUNSUPPORTED;
  }
  gdn_tc192_kernel<INV><<<grid, kThreads, kSmemBytes, st>>>(map_x, map_out, params, addend, HW, tps, total, accumulate);
  return check_launch("gdn_f32(tcgen05, C=192)");
}

}  // namespace tc192

int launch_gdn_tc192(const float* x, const float* params, const float* addend, float* out, int N, int64_t HW,
                     int inverse, cudaStream_t st) {
  using namespace tc192;
  if (HW % 4 != 0 || HW >= (1ll << 31) || (int64_t)N * kC >= (1ll << 31)) {
    set_error("gdn_f32: the C = 192 tcgen05 kernel needs HW %% 4 == 0 (got HW=%lld)", (long long)HW);
    return B200VC_EUNSUPPORTED;
  }
  const int accumulate = (addend != nullptr && addend == out) ? 1 : 0;
  if (accumulate) addend = nullptr;
  CUtensorMap map_x, map_out;
  if (!make_rows_map(&map_x, x, (int64_t)N * kC, HW, kC) || !make_rows_map(&map_out, out, (int64_t)N * kC, HW, kRows)) {
    set_error("gdn_f32: cuTensorMapEncodeTiled failed");
    return B200VC_EUNSUPPORTED;
  }
  const int64_t tps = (HW + kTileP - 1) / kTileP;
  const int64_t total = tps * N;
  if (total >= (1ll << 30)) {
    set_error("gdn_f32: too many tiles");
    return B200VC_EINVAL;
  }
  const int pairs = (int)(total < sm_count() / 2 ? total : sm_count() / 2);
  const int grid = 2 * (pairs > 0 ? pairs : 1);
  switch (inverse) {
    case 0: return launch_inv<0>(map_x, map_out, params, addend, HW, (int)tps, (int)total, accumulate, grid, st);
    case 1: return launch_inv<1>(map_x, map_out, params, addend, HW, (int)tps, (int)total, accumulate, grid, st);
    default: return launch_inv<2>(map_x, map_out, params, addend, HW, (int)tps, (int)total, accumulate, grid, st);
  }
}

}  // namespace b200vc
