// K-GDN / K-IGDN on the 5th-gen tensor cores (tcgen05, sm_100a), C = 128.
//
//   norm[i, p] = beta_i + sum_j gamma[i, j] * x[j, p]^2       (the one dense contraction of the hot path)
//   out[i, p]  = x[i, p] * (rsqrt | sqrt)(norm[i, p]) [+ addend[i, p]]
//
// fp32-class accuracy from TF32 tensor cores ("3xTF32"): x^2 = hi + lo and gamma = ghi + glo with hi/ghi rounded
// to TF32 (cvt.rna) and lo/glo the exact remainders; D = ghi*hi + ghi*lo + glo*hi, fp32 accumulation in TMEM
// (dropped term glo*lo ~ 2^-22 relative).  See DESIGN.md for the error budget.
//
// Mapping (one persistent CTA per SM, 576 threads, warp-specialised, mbarrier pipelines):
//   * A operand  = gamma (M = 128 output channels x K = 128 input channels), resident in TMEM for the whole
//     kernel (hi: columns [0,128), lo: [128,256)), written once with tcgen05.st.
//   * B operand  = x^2 tile (K = 128 channels x N = 64 positions), MN-major (positions contiguous, exactly the
//     NCHW layout), SWIZZLE_128B_BASE32B.  TMA (SWIZZLE_128B_ATOM_32B) loads the raw x tile as two [128 x 32]
//     boxes whose swizzled image *is* the canonical MN-major UMMA layout, so the square/split pass is a pure
//     elementwise smem->smem copy with no index arithmetic.
//   * D          = 128 lanes (channels) x 64 columns (positions) fp32 in TMEM; two accumulators per tile
//     (D1 = ghi*hi, D2 = the small cross terms), double buffered: 256 + 2*128 = all 512 TMEM columns.
//   * warp 0: TMA producer (ring of R raw tiles) | warp 1: TMEM alloc + MMA issuer (48 tcgen05.mma per tile) |
//     warps 2-9: square + hi/lo split (ring of A operand pairs) | warps 10-17: epilogue (tcgen05.ld ->
//     D1+D2+beta -> rsqrt/sqrt -> * x -> in place into the raw tile -> TMA store; the raw slot is released
//     one tile later so the store drains behind the next tile's epilogue).
// Algorithmic HBM traffic: 2*128*4 B per position (+128*4 with addend); 3 * 2*128*128 tensor flop per position.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "gdn_tc_common.cuh"

namespace b200vc {

namespace tc {

constexpr int kC = 128;            // channels (UMMA M and K)
constexpr int kTileP = 64;         // positions per tile (UMMA N)
constexpr int kHalfBytes = kC * 32 * 4;          // one [128 x 32] fp32 box = 16 KB
constexpr int kTileBytes = 2 * kHalfBytes;       // 32 KB
constexpr int kXfWarps = 8;                      // square + split warps
constexpr int kEpiWarps = 8;                     // 2 per TMEM lane quarter (one per 32-column half)
constexpr int kFirstXf = 2, kFirstEpi = kFirstXf + kXfWarps;
constexpr int kThreads = (kFirstEpi + kEpiWarps) * 32;   // 576
constexpr uint32_t kTmemCols = 512;
// TMEM columns: gamma hi | gamma lo | 2 accumulator stages x (D1: ghi*hi, D2: ghi*lo + glo*hi)
constexpr uint32_t kColGhi = 0, kColGlo = 128, kColD = 256, kColDStage = 2 * kTileP;

// Ring depths: R raw tiles (TMA load .. TMA store), A hi/lo operand pairs (square/split .. MMA done).
// (A half-tile operand ring with N = 32 MMAs was measured SLOWER: a tcgen05.mma of N = 32 costs as much as
// one of N = 64, so halving N doubles tensor time.)
template <int R, int A>
struct Cfg {
  static constexpr int kRaw = R, kAb = A;
  static constexpr int kRawOff = 0;
  static constexpr int kAbOff = R * kTileBytes;              // A x (hi | lo)
  static constexpr int kBarOff = kAbOff + A * 2 * kTileBytes;
  static constexpr int kNumBars = 2 * R + 4 * A + 5;
  static constexpr int kSmemBytes = kBarOff + 8 * kNumBars + 16 + 1024 /*alignment slack*/;
};

// MN-major TF32 operand: the only legal smem layout is SWIZZLE_128B_BASE32B (cutlass sm100_common.inl:
// "for mn-major tf32 operands, SW128_32B is the only available smem layout"): rows of 128 B (32 positions),
// 32-byte chunk index XOR (row & 3), canonical ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-byte units
// (cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::MN>).  LBO = distance between the two 32-position
// atoms, SBO = distance between groups of 4 channels.  TMA produces exactly this image with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((kHalfBytes >> 4) & 0x3FFF) << 16;  // leading byte offset  (16 KB between the two atoms)
  d |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;         // stride byte offset   (4 rows x 128 B)
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)1 << 61;                              // SWIZZLE_128B_BASE32B
  return d;
}
// kind::tf32, D = F32, A = B = TF32, A K-major (TMEM), B MN-major, N = 64, M = 128.
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) |
                            ((uint32_t)(kTileP >> 3) << 17) | ((uint32_t)(kC >> 4) << 24);

__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
// params layout (gdn.cu): [0,C) beta | gamma | gammaT | hi[i][j] (C*C) | lo[i][j] (C*C)
template <class CFG, int INV>
__global__ void __launch_bounds__(kThreads, 1)
gdn_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_out,
              const __grid_constant__ CUtensorMap map_add, const float* __restrict__ params,
              const float* __restrict__ addend, int64_t HW, int tiles_per_sample, int total_tiles, int accumulate,
              long long* __restrict__ trace) {
  // trace (diagnostics, normally null): CTA 0 stamps clock64() per tile and pipeline event, 16 slots per tile
#define B200VC_TRACE(slot)                                                               \
  do {                                                                                   \
    if (trace != nullptr && blockIdx.x == 0 && k < 256) trace[k * 16 + (slot)] = clock64(); \
  } while (0)
  // addend handling: accumulate != 0  => `out` already holds the addend (in-place residual add): the result
  //                                      tile leaves through a TMA reduce-add, no addend traffic in the SM;
  //                  addend != null   => the epilogue reads it from global (L2-prefetched by the producer).
  constexpr int R = CFG::kRaw, A = CFG::kAb;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  auto raw_addr = [&](int r) { return smem_base + CFG::kRawOff + r * kTileBytes; };
  auto hi_addr = [&](int a) { return smem_base + CFG::kAbOff + a * 2 * kTileBytes; };
  auto lo_addr = [&](int a) { return smem_base + CFG::kAbOff + a * 2 * kTileBytes + kTileBytes; };
  const uint32_t bar_base = smem_base + CFG::kBarOff;
  auto raw_full = [&](int r) { return bar_base + 8 * r; };
  auto raw_empty = [&](int r) { return bar_base + 8 * (R + r); };
  // operand barriers exist per channel half (c = 0: channels 0-63 = k-steps 0-7, c = 1: the rest), so that the split
  // of tile k+1 overlaps the second half of tile k's MMAs although there is a single operand stage
  auto ab_full = [&](int a, int c) { return bar_base + 8 * (2 * R + 2 * a + c); };
  auto ab_empty = [&](int a, int c) { return bar_base + 8 * (2 * R + 2 * A + 2 * a + c); };
  auto d_full = [&](int d) { return bar_base + 8 * (2 * R + 4 * A + d); };
  auto d_empty = [&](int d) { return bar_base + 8 * (2 * R + 4 * A + 2 + d); };
  const uint32_t gamma_ready = bar_base + 8 * (2 * R + 4 * A + 4);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + CFG::kBarOff + 8 * CFG::kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int r = 0; r < R; ++r) {
      mbar_init(raw_full(r), 1);
      mbar_init(raw_empty(r), 1);
    }
    for (int a = 0; a < A; ++a)
      for (int c = 0; c < 2; ++c) {
        mbar_init(ab_full(a, c), kXfWarps);   // one arrival per warp (lane 0 after __syncwarp)
        mbar_init(ab_empty(a, c), 1);
      }
    for (int d = 0; d < 2; ++d) {
      mbar_init(d_full(d), 1);
      mbar_init(d_empty(d), kEpiWarps);
    }
    mbar_init(gamma_ready, kXfWarps + kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_out)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_add)) : "memory");
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

  // gamma hi/lo -> TMEM (lane = output channel i, column = input channel j), by the 16 split + epilogue warps (all
  // idle until the first tile lands): warp w owns lane quarter (w & 3); the four warps of a quarter take
  // {hi, lo} x {columns 0-63, 64-127}: two dependent L2 round trips per warp instead of four.
  if (warp >= kFirstXf) {
    const int q = warp & 3, grp = (warp - kFirstXf) >> 2;   // kFirstXf = 2: warps 2..17 -> groups 0..3
    const int which = grp & 1, half = grp >> 1;
    const int i = 32 * q + lane;
    const float* src_row = params + kC + 2 * kC * kC + (int64_t)which * kC * kC + (int64_t)i * kC;
#pragma unroll 1
    for (int part = 2 * half; part < 2 * half + 2; ++part) {
      uint32_t v[32];
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(src_row + part * 32 + e));
        v[e] = __float_as_uint(t.x); v[e + 1] = __float_as_uint(t.y);
        v[e + 2] = __float_as_uint(t.z); v[e + 3] = __float_as_uint(t.w);
      }
      tmem_st32(tmem + ((uint32_t)(32 * q) << 16) + (which ? kColGlo : kColGhi) + part * 32, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(gamma_ready);  // only the MMA issuer waits for it: loads and the split start now
  }

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int k = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++k) {
        const int r = k % R, ph = (k / R) & 1;
        const int row0 = (tile / tiles_per_sample) * kC;
        const int p0 = (tile % tiles_per_sample) * kTileP;
        mbar_wait(raw_empty(r), ph ^ 1);
        B200VC_TRACE(0);  // load issued
        mbar_arrive_expect_tx(raw_full(r), kTileBytes);
        tma_load_2d(raw_addr(r), &map_x, p0, row0, raw_full(r));
        tma_load_2d(raw_addr(r) + kHalfBytes, &map_x, p0 + 32, row0, raw_full(r));
        if (addend != nullptr) {
          tma_prefetch_l2_2d(&map_add, p0, row0);
          if ((int64_t)p0 + 32 < HW) tma_prefetch_l2_2d(&map_add, p0 + 32, row0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs the loop with warp-uniform values (the TMEM base goes through a shuffle so that ptxas
    // knows it is uniform) and one elected lane issues.  Written as `if (lane == 0) { loop }` the operands live in
    // per-thread registers and ptxas wraps EVERY tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall
    // loop, which costs more than the MMA itself (tools/mma_bench.cu).
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    mbar_wait(gamma_ready, 0);
    tc_fence_after();
    int k = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++k) {
      const int a = k % A, pa = (k / A) & 1, d = k & 1, pd = (k >> 1) & 1;
      mbar_wait(d_empty(d), pd ^ 1);
      if (lane == 0) B200VC_TRACE(14);  // accumulator stage free
      // Two accumulators: the tensor core's fp32 accumulation truncates, so the 32 small cross-term steps
      // go to their own accumulator and never disturb the 16-step main sum; the epilogue adds them (RN).
      const uint32_t d1 = tmem_u + kColD + kColDStage * d, d2 = d1 + kTileP;
      const uint64_t hi_desc = make_b_desc(hi_addr(a)), lo_desc = make_b_desc(lo_addr(a));
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        mbar_wait(ab_full(a, c), pa);
        tc_fence_after();
        if (elect_one()) {
          if (c == 0) B200VC_TRACE(3);  // operands ready, MMA issue starts
          // descriptor start address advances 1024 B (8 channel rows) per k-step
#pragma unroll
          for (int g = 8 * c; g < 8 * c + 8; ++g)   // D1 += ghi * hi
            umma_tf32_ts(d1, tmem_u + kColGhi + 8 * g, hi_desc + (uint64_t)(g * (1024 >> 4)), g != 0 ? 1u : 0u);
#pragma unroll
          for (int g = 8 * c; g < 8 * c + 8; ++g)   // D2 += ghi * lo
            umma_tf32_ts(d2, tmem_u + kColGhi + 8 * g, lo_desc + (uint64_t)(g * (1024 >> 4)), g != 0 ? 1u : 0u);
#pragma unroll
          for (int g = 8 * c; g < 8 * c + 8; ++g)   // D2 += glo * hi
            umma_tf32_ts(d2, tmem_u + kColGlo + 8 * g, hi_desc + (uint64_t)(g * (1024 >> 4)), 1u);
          umma_commit(ab_empty(a, c));  // this half of the operand pair is consumed
          if (c == 1) {
            umma_commit(d_full(d));     // accumulators ready
            B200VC_TRACE(15);           // all MMAs of the tile issued
          }
        }
        __syncwarp();
      }
    }
  } else if (warp < kFirstEpi) {
    // ------------------------------------------------------------------ square + TF32 hi/lo split (8 warps)
    const int t = threadIdx.x - kFirstXf * 32;
    constexpr int kIters = kTileBytes / 16 / (kXfWarps * 32);  // 8 float4 per thread
    int k = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++k) {
      const int r = k % R, pr = (k / R) & 1, a = k % A, pa = (k / A) & 1;
      mbar_wait(raw_full(r), pr);
      if (t == 0) B200VC_TRACE(1);  // raw tile landed
      const float4* raw4 = reinterpret_cast<const float4*>(smem_gen + CFG::kRawOff + r * kTileBytes);
      float4* hi4 = reinterpret_cast<float4*>(smem_gen + CFG::kAbOff + a * 2 * kTileBytes);
      float4* lo4 = hi4 + kTileBytes / 16;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        mbar_wait(ab_empty(a, c), pa ^ 1);
        if (t == 0 && c == 0) B200VC_TRACE(2);  // operand slot free: split starts
        // channel half c: rows 64c..64c+63 of both 32-position atoms = 2 x 512 float4
        float4 v[kIters / 2];
#pragma unroll
        for (int it = 0; it < kIters / 2; ++it) {
          const int idx = it * (kXfWarps * 32) + t;            // 0..1023
          v[it] = raw4[(idx >> 9) * 1024 + c * 512 + (idx & 511)];   // all loads in flight
        }
#pragma unroll
        for (int it = 0; it < kIters / 2; ++it) {
          const int idx = it * (kXfWarps * 32) + t;
          const int o = (idx >> 9) * 1024 + c * 512 + (idx & 511);
          float4 sq, hh, ll;
          sq.x = __fmul_rn(v[it].x, v[it].x); sq.y = __fmul_rn(v[it].y, v[it].y);
          sq.z = __fmul_rn(v[it].z, v[it].z); sq.w = __fmul_rn(v[it].w, v[it].w);
          hh.x = to_tf32_rna(sq.x); hh.y = to_tf32_rna(sq.y); hh.z = to_tf32_rna(sq.z); hh.w = to_tf32_rna(sq.w);
          ll.x = __fsub_rn(sq.x, hh.x); ll.y = __fsub_rn(sq.y, hh.y);
          ll.z = __fsub_rn(sq.z, hh.z); ll.w = __fsub_rn(sq.w, hh.w);
          hi4[o] = hh;
          lo4[o] = ll;
        }
        if (t == 0 && c == 1) B200VC_TRACE(12);  // split stores issued
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (t == 0 && c == 1) B200VC_TRACE(13);  // fence done
        if (lane == 0) mbar_arrive(ab_full(a, c));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int h = (warp - kFirstEpi) >> 2;  // 32-column (= 32-position) half of the tile
    const int i = 32 * q + lane;      // output channel == TMEM lane
    const float beta = __ldg(params + i);
    const bool leader = (threadIdx.x == kFirstEpi * 32);
    const bool flip = (i >> 2) & 1;   // rows i and i+4 share a swizzle phase: visit chunk pairs in opposite order
    int k = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++k) {
      const int r = k % R, d = k & 1, pd = (k >> 1) & 1;
      const int row0 = (tile / tiles_per_sample) * kC;
      const int p0 = (tile % tiles_per_sample) * kTileP;
      // The x row segment is fetched before the accumulators are awaited (the raw tile landed long ago), so the
      // shared-memory latency hides behind the d_full wait.  Logical 16-byte chunk cc of row i lives at 32-byte
      // chunk ((cc >> 1) ^ (i & 3)), same 16-byte half; lanes with `flip` take the odd chunk of each pair first
      // => the 8 rows of a quarter-warp hit 8 distinct 16-byte bank groups (conflict-free LDS.128 / STS.128).
      float4* raw4 = reinterpret_cast<float4*>(smem_gen + CFG::kRawOff + r * kTileBytes) + h * (kHalfBytes / 16) + i * 8;
      mbar_wait(raw_full(r), (k / R) & 1);
      float4 xv[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int cc = c ^ (int)flip;
        xv[c] = raw4[(((cc >> 1) ^ (i & 3)) << 1) | (cc & 1)];
      }
      mbar_wait(d_full(d), pd);
      tc_fence_after();
      if (leader) B200VC_TRACE(4);  // accumulators complete: epilogue starts
      // The accumulators come in four 8-column pieces (D1 and D2 each), the next piece loading while the current
      // one is finished: with the 32 registers of x this stays inside the 96-register budget of a 576-thread CTA.
      const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + kColD + kColDStage * d + 32 * h;
      const float* arow = addend ? addend + ((int64_t)row0 + i) * HW + p0 + 32 * h : nullptr;
      const int64_t pbase = (int64_t)p0 + 32 * h;
      uint32_t v1[2][8], v2[2][8];
      tmem_ld8_issue(taddr, v1[0]);
      tmem_ld8_issue(taddr + kTileP, v2[0]);
#pragma unroll
      for (int pc = 0; pc < 4; ++pc) {
        tmem_ld_wait();
        if (pc == 0 && leader) B200VC_TRACE(8);  // first accumulators in registers
        if (pc < 3) {
          tmem_ld8_issue(taddr + 8 * (pc + 1), v1[(pc + 1) & 1]);
          tmem_ld8_issue(taddr + kTileP + 8 * (pc + 1), v2[(pc + 1) & 1]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d_empty(d));  // accumulators drained: the MMA warp may start tile k+2
        }
        float nr[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          nr[j] = __fadd_rn(__fadd_rn(__uint_as_float(v1[pc & 1][j]), __uint_as_float(v2[pc & 1][j])), beta);
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
          raw4[(((cc >> 1) ^ (i & 3)) << 1) | (cc & 1)] = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
      if (leader) B200VC_TRACE(9);  // result written to the raw slot
      fence_proxy_async();  // result tile (generic writes) -> visible to the TMA store
      if (leader) B200VC_TRACE(10);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (leader) {
        B200VC_TRACE(5);  // epilogue math done, store issued
        if (accumulate) {
          tma_reduce_add_2d(&map_out, raw_addr(r), p0, row0);
          if ((int64_t)p0 + 32 < HW) tma_reduce_add_2d(&map_out, raw_addr(r) + kHalfBytes, p0 + 32, row0);
        } else {
          tma_store_2d(&map_out, raw_addr(r), p0, row0);
          if ((int64_t)p0 + 32 < HW) tma_store_2d(&map_out, raw_addr(r) + kHalfBytes, p0 + 32, row0);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        // release the PREVIOUS tile's raw slot once its store has finished reading shared memory; the store
        // just issued keeps draining while the next tile's epilogue runs
        if (k > 0) {
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          mbar_arrive(raw_empty((k - 1) % R));
        }
        B200VC_TRACE(6);  // previous tile's slot released
      }
    }
    if (leader && k > 0) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      mbar_arrive(raw_empty((k - 1) % R));
      tma_store_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static bool make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t HW) {
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr) return false;
  cuuint64_t dims[2] = {(cuuint64_t)HW, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)HW * sizeof(float)};
  cuuint32_t box[2] = {32, (cuuint32_t)kC};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

#ifdef B200VC_ENABLE_GDN_TRACE
long long* g_trace = nullptr;  // set through b200vc_debug_set_gdn_trace (debug builds only: include/b200vc_debug.h)
#else
static long long* const g_trace = nullptr;  // production builds carry no mutable global: the trace stores are dead code
#endif

template <class CFG, int INV>
static int launch_inv(const CUtensorMap& map_x, const CUtensorMap& map_out, const CUtensorMap& map_add,
                      const float* params, const float* addend, int64_t HW, int tps, int total, int accumulate,
                      int grid, cudaStream_t st) {
  // once per device and instantiation, safe under concurrent host threads
  static std::once_flag once[64];
  static bool ok[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return B200VC_EUNSUPPORTED;
  std::call_once(once[dev], [&]() {
    ok[dev] = cudaFuncSetAttribute(gdn_tc_kernel<CFG, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   CFG::kSmemBytes) == cudaSuccess;
    if (!ok[dev]) (void)cudaGetLastError();
  });
  if (!ok[dev]) {
    set_error("gdn_f32: cannot reserve %d B of shared memory", CFG::kSmemBytes);
    return B200VC_EUNSUPPORTED;
  }
  gdn_tc_kernel<CFG, INV><<<grid, kThreads, CFG::kSmemBytes, st>>>(map_x, map_out, map_add, params, addend, HW, tps,
                                                                   total, accumulate, g_trace);
  return check_launch("gdn_f32(tcgen05)");
}

// inverse: 0 = GDN (x * rsqrt(norm)), 1 = IGDN (x * sqrt(norm)), 2 = norm only (debug)
template <class CFG>
static int launch_cfg(const CUtensorMap& map_x, const CUtensorMap& map_out, const CUtensorMap& map_add,
                      const float* params, const float* addend, int64_t HW, int tps, int total, int inverse,
                      int accumulate, int grid, cudaStream_t st) {
  switch (inverse) {
    case 0: return launch_inv<CFG, 0>(map_x, map_out, map_add, params, addend, HW, tps, total, accumulate, grid, st);
    case 1: return launch_inv<CFG, 1>(map_x, map_out, map_add, params, addend, HW, tps, total, accumulate, grid, st);
    default: return launch_inv<CFG, 2>(map_x, map_out, map_add, params, addend, HW, tps, total, accumulate, grid, st);
  }
}

}  // namespace tc

int launch_gdn_tc(const float* x, const float* params, const float* addend, float* out, int N, int C, int64_t HW,
                  int inverse, cudaStream_t st) {
  using namespace tc;
  if (C != kC || HW % 4 != 0 || HW >= (1ll << 31) || (int64_t)N * C >= (1ll << 31)) {
    set_error("gdn_f32: tcgen05 kernel needs C == 128 and HW %% 4 == 0 (got C=%d, HW=%lld)", C, (long long)HW);
    return B200VC_EUNSUPPORTED;
  }
  // in-place residual add: `out` already holds the addend => accumulate through a TMA reduce-add
  const int accumulate = (addend != nullptr && addend == out) ? 1 : 0;
  if (accumulate) addend = nullptr;
  CUtensorMap map_x, map_out, map_add;
  if (!make_map(&map_x, x, (int64_t)N * C, HW) || !make_map(&map_out, out, (int64_t)N * C, HW) ||
      !make_map(&map_add, addend ? addend : x, (int64_t)N * C, HW)) {
    set_error("gdn_f32: cuTensorMapEncodeTiled failed");
    return B200VC_EUNSUPPORTED;
  }
  const int64_t tps = (HW + kTileP - 1) / kTileP;
  const int64_t total = tps * N;
  if (total >= (1ll << 31)) {
    set_error("gdn_f32: too many tiles");
    return B200VC_EINVAL;
  }
  const int grid = (int)(total < sm_count() ? total : sm_count());
  static const int cfg = []() {
    const char* e = getenv("B200VC_GDN_TC_CFG");
    return e ? atoi(e) : 0;
  }();
  switch (cfg) {
    // raw-ring depth wins: <5,1> is the default; the other two stay for experiments (B200VC_GDN_TC_CFG)
    case 1: return launch_cfg<Cfg<3, 2>>(map_x, map_out, map_add, params, addend, HW, (int)tps, (int)total, inverse,
                                          accumulate, grid, st);
    case 2: return launch_cfg<Cfg<4, 1>>(map_x, map_out, map_add, params, addend, HW, (int)tps, (int)total, inverse,
                                          accumulate, grid, st);
    default: return launch_cfg<Cfg<5, 1>>(map_x, map_out, map_add, params, addend, HW, (int)tps, (int)total, inverse,
                                          accumulate, grid, st);
  }
}

}  // namespace b200vc

#ifdef B200VC_ENABLE_GDN_TRACE
// Diagnostics (debug builds: NVCC_FLAGS=-DB200VC_ENABLE_GDN_TRACE, include/b200vc_debug.h): device buffer of 256 x 16
// clock64() stamps written by CTA 0 of the next tcgen05 GDN launches (slots: 0 load issued, 1 raw landed, 2 split
// starts, 3 MMA issue, 4 epilogue starts, 5 store issued, 6 slot released).  The buffer must outlive every launch.
extern "C" __attribute__((visibility("default"))) void b200vc_debug_set_gdn_trace(long long* device_buffer) {
  b200vc::tc::g_trace = device_buffer;
}
#endif
