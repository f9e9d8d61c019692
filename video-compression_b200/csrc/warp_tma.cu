// K-WARP, TMA-staged variant (C = 3): the reference-frame tile around each CTA's flow footprint is brought into
// shared memory by ONE 3-D TMA box copy (cp.async.bulk.tensor.3d, zero fill outside the frame), then the 12
// bilinear taps per pixel are shared-memory reads.  Replaces the same reference sites as warp.cu
// (LHBDC/model/m.py:111-126, flow.py:15-25, Flex .../b_model.py:99-112, ICIP2024/src/model/m.py:262-282).
//
// Why: the direct-gather kernel is latency-bound (ncu: 79 % long-scoreboard stalls, two dependent DRAM round
// trips per pixel, thousands of 32-byte L1 misses in flight per SM).  Here a CTA issues one bulk request for its
// whole footprint and every tap is a conflict-free LDS for smooth flows.
//
// CTA = 64 x 32 output pixels (256 threads x 8 pixels).  Flow -> exact source coordinates (same arithmetic as
// warp.cu, bit-identical results) -> block-wide bounding box of the footprint -> if it fits the 96 x 48 box the tile
// is staged, otherwise this CTA falls back to global gathers (large motion, Flex vectors far outside the frame).
#include <cuda.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "warp_common.cuh"

namespace b200vc {
namespace wt {

constexpr int kTW = 64, kTH = 32, kPX = 8;   // tile and pixels per thread (rows ty, ty+4, ...)
constexpr int kBW = 96, kBH = 48;            // staged box (fp32 elements); 3 planes = 55 296 B => 4 CTAs / SM
constexpr int kThreads = 256;
constexpr int kBoxBytes = 3 * kBW * kBH * 4;
constexpr int kSmemBytes = kBoxBytes + 128;  // + alignment slack

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// SPY = SPyNet level form (spynet.cu, LHBDC/model/flow.py:93-98): `flow` is the PREVIOUS level's flow [N,2,hp,wp]
// (or NULL = zeros); the displacement is 2 * its x2 align_corners=True upsample, replicate-padded; `out` is the
// 8-plane conv input [first | warped second | up] and `img` is `second`.
struct SpyArgs {
  const float* first;
  int64_t first_bs;
  int hp, wp;
  float sy, sx;
};

__device__ __forceinline__ void up_ac(int dst, int in_size, float scale, int& i0, int& i1, float& l0, float& l1) {
  // ATen upsample_bilinear2d, align_corners=True: src = scale * dst (see spynet.cu)
  const float src = __fmul_rn(scale, (float)dst);
  i0 = (int)src;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = __fsub_rn(src, (float)i0);
  l0 = __fsub_rn(1.f, l1);
}

template <int VARIANT, bool SPY>
__global__ void __launch_bounds__(kThreads, SPY ? 3 : 0)
warp_tma_kernel(const __grid_constant__ CUtensorMap map_img, const float* __restrict__ img,
                const float* __restrict__ flow, const float* __restrict__ tab_x, const float* __restrict__ tab_y,
                float* __restrict__ out, int64_t out_bs, WarpGeom g, SpyArgs spy) {
  constexpr bool BORDER = (VARIANT != B200VC_WARP_FLEX);
  extern __shared__ uint8_t smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));  // [3][kBH][kBW]
  __shared__ int s_red[kThreads / 32][4];
  __shared__ int s_box[3];                         // bx0, by0, fits
  __shared__ __align__(8) uint64_t s_bar;

  const int tid = threadIdx.x;
  const int tx = tid & (kTW - 1), ty = tid >> 6;
  const int x = blockIdx.x * kTW + tx;
  const int n = blockIdx.z;
  const int HW = g.H * g.W;
  const bool xin = x < g.W;
  const int xc = xin ? x : g.W - 1;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- phase A: flow -> source coordinates (all 16 loads of the thread in flight together)
  float u[kPX], v[kPX];
  if (!SPY) {
    const float* fbase = flow + (int64_t)n * 2 * HW;
#pragma unroll
    for (int k = 0; k < kPX; ++k) {
      const int y = min(blockIdx.y * kTH + ty + 4 * k, g.H - 1);
      u[k] = __ldg(fbase + y * g.W + xc);
      v[k] = __ldg(fbase + HW + y * g.W + xc);
    }
  } else {
#pragma unroll
    for (int k = 0; k < kPX; ++k) u[k] = v[k] = 0.f;
    if (flow != nullptr) {
      const int hp = spy.hp, wp = spy.wp;
      const float* fu = flow + (int64_t)n * 2 * hp * wp;
      const float* fv = fu + hp * wp;
      int xi0, xi1;
      float xl0, xl1;
      up_ac(min(xc, 2 * wp - 1), wp, spy.sx, xi0, xi1, xl0, xl1);  // replicate pad = clamp of the destination index
      // two batches of 4 pixels: 32 taps in flight per thread without the register cost of all 64
#pragma unroll
      for (int k0 = 0; k0 < kPX; k0 += 4) {
        float ta[4][4], tb[4][4], yl0[4], yl1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int y = min(min(blockIdx.y * kTH + ty + 4 * (k0 + j), g.H - 1), 2 * hp - 1);
          int yi0, yi1;
          up_ac(y, hp, spy.sy, yi0, yi1, yl0[j], yl1[j]);
          const int o00 = yi0 * wp + xi0, o01 = yi0 * wp + xi1, o10 = yi1 * wp + xi0, o11 = yi1 * wp + xi1;
          ta[j][0] = __ldg(fu + o00); ta[j][1] = __ldg(fu + o01); ta[j][2] = __ldg(fu + o10); ta[j][3] = __ldg(fu + o11);
          tb[j][0] = __ldg(fv + o00); tb[j][1] = __ldg(fv + o01); tb[j][2] = __ldg(fv + o10); tb[j][3] = __ldg(fv + o11);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float ui = __fmaf_rn(yl0[j], __fmaf_rn(xl0, ta[j][0], __fmul_rn(xl1, ta[j][1])),
                                     __fmul_rn(yl1[j], __fmaf_rn(xl0, ta[j][2], __fmul_rn(xl1, ta[j][3]))));
          const float vi = __fmaf_rn(yl0[j], __fmaf_rn(xl0, tb[j][0], __fmul_rn(xl1, tb[j][1])),
                                     __fmul_rn(yl1[j], __fmaf_rn(xl0, tb[j][2], __fmul_rn(xl1, tb[j][3]))));
          u[k0 + j] = __fmul_rn(ui, 2.0f);
          v[k0 + j] = __fmul_rn(vi, 2.0f);
        }
      }
    }
    // planes 6, 7 of the level's conv input = the upsampled flow itself
    float* up = out + (int64_t)n * out_bs + (int64_t)6 * HW;
#pragma unroll
    for (int k = 0; k < kPX; ++k) {
      const int y = blockIdx.y * kTH + ty + 4 * k;
      if (xin && y < g.H) {
        up[y * g.W + x] = u[k];
        up[HW + y * g.W + x] = v[k];
      }
    }
  }
  const float txv = BORDER ? __ldg(tab_x + xc) : 0.f;
  float ix[kPX], iy[kPX];
  int mnx = 0x7fffffff, mxx = -0x7fffffff - 1, mny = 0x7fffffff, mxy = -0x7fffffff - 1;
#pragma unroll
  for (int k = 0; k < kPX; ++k) {
    const int y = min(blockIdx.y * kTH + ty + 4 * k, g.H - 1);
    const float tyv = BORDER ? __ldg(tab_y + y) : 0.f;
    coords<VARIANT, true>(g, xc, y, u[k], v[k], txv, tyv, ix[k], iy[k]);
    const int x0 = (int)floorf(ix[k]), y0 = (int)floorf(iy[k]);
    mnx = min(mnx, x0); mxx = max(mxx, x0);
    mny = min(mny, y0); mxy = max(mxy, y0);
  }
  // ---- block-wide bounding box of the footprint (taps x0..x0+1, y0..y0+1)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  if ((tid & 31) == 0) {
    s_red[tid >> 5][0] = mnx; s_red[tid >> 5][1] = mxx; s_red[tid >> 5][2] = mny; s_red[tid >> 5][3] = mxy;
  }
  __syncthreads();
  if (tid == 0) {
    int a = s_red[0][0], b = s_red[0][1], c = s_red[0][2], d = s_red[0][3];
    for (int w = 1; w < kThreads / 32; ++w) {
      a = min(a, s_red[w][0]); b = max(b, s_red[w][1]); c = min(c, s_red[w][2]); d = max(d, s_red[w][3]);
    }
    if (g.arith == 0) a &= ~3;  // start the box on a 16-byte boundary of the row (arith != 0 here: debug switch)
    const bool fits = (b + 1 - a + 1 <= kBW) && (d + 1 - c + 1 <= kBH) && a > -(1 << 20) && c > -(1 << 20) &&
                      b < (1 << 20) && d < (1 << 20);
    s_box[0] = a; s_box[1] = c; s_box[2] = fits ? 1 : 0;
    if (fits) {
      const uint32_t bar = smem_u32(&s_bar);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kBoxBytes) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
          ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&map_img)), "r"(bar), "r"(a), "r"(c), "r"(n * 3)
          : "memory");
    }
  }
  __syncthreads();
  const int bx0 = s_box[0], by0 = s_box[1];
  const bool staged = s_box[2] != 0;
  float* op = out + (int64_t)n * out_bs + (SPY ? (int64_t)3 * HW : 0);
  if (SPY) {
    // planes 0-2 = `first`, copied while the box is in flight
    const float* fp = spy.first + (int64_t)n * spy.first_bs;
    float* o0 = out + (int64_t)n * out_bs;
    float f[kPX][3];
#pragma unroll
    for (int k = 0; k < kPX; ++k) {
      const int y = min(blockIdx.y * kTH + ty + 4 * k, g.H - 1);
#pragma unroll
      for (int c = 0; c < 3; ++c) f[k][c] = __ldg(fp + (int64_t)c * HW + y * g.W + xc);
    }
#pragma unroll
    for (int k = 0; k < kPX; ++k) {
      const int y = blockIdx.y * kTH + ty + 4 * k;
      if (xin && y < g.H) {
#pragma unroll
        for (int c = 0; c < 3; ++c) o0[(int64_t)c * HW + y * g.W + x] = f[k][c];
      }
    }
  }

  if (staged) {
    // wait for the box (phase 0 of a one-shot barrier)
    {
      const uint32_t bar = smem_u32(&s_bar);
      uint32_t done = 0;
      for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar) : "memory");
        if (spins > (1u << 26)) __trap();
      }
    }
#pragma unroll
    for (int k = 0; k < kPX; ++k) {
      const int y = blockIdx.y * kTH + ty + 4 * k;
      const float fx = floorf(ix[k]), fy = floorf(iy[k]);
      const int x0 = (int)fx, y0 = (int)fy;
      const float dx1 = __fsub_rn((float)(x0 + 1), ix[k]), dx0 = __fsub_rn(ix[k], (float)x0);
      const float dy1 = __fsub_rn((float)(y0 + 1), iy[k]), dy0 = __fsub_rn(iy[k], (float)y0);
      const float w00 = __fmul_rn(dx1, dy1), w01 = __fmul_rn(dx0, dy1);
      const float w10 = __fmul_rn(dx1, dy0), w11 = __fmul_rn(dx0, dy0);
      // taps outside the frame read TMA's zero fill: v * w == 0 exactly, the same as ATen skipping the tap
      const float* t0 = tile + (y0 - by0) * kBW + (x0 - bx0);
      float r[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* t = t0 + c * (kBH * kBW);
        float acc = __fmaf_rn(t[0], w00, 0.f);
        acc = __fmaf_rn(t[1], w01, acc);
        acc = __fmaf_rn(t[kBW], w10, acc);
        r[c] = __fmaf_rn(t[kBW + 1], w11, acc);
      }
      if (xin && y < g.H) {
#pragma unroll
        for (int c = 0; c < 3; ++c) op[(int64_t)c * HW + y * g.W + x] = r[c];
      }
    }
  } else {
    const float* ip = img + (int64_t)n * 3 * HW;
#pragma unroll
    for (int k = 0; k < kPX; ++k) {
      const int y = blockIdx.y * kTH + ty + 4 * k;
      const Taps t = make_taps<BORDER>(ix[k], iy[k], g.H, g.W);
      float r[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) r[c] = sample<BORDER>(ip + (int64_t)c * HW, t);
      if (xin && y < g.H) {
#pragma unroll
        for (int c = 0; c < 3; ++c) op[(int64_t)c * HW + y * g.W + x] = r[c];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Fused LHBDC motion compensation (LHBDC/model/m.py:55-63) with TMA-staged taps: per direction, the x4-upsampled
// flow (from the shared-memory quarter-resolution tile, as in warp.cu's warp2 kernel) -> coordinates -> bounding
// box -> one 3-D TMA box -> taps from shared memory.  The two directions reuse the box buffer (barrier phase = dir).
constexpr int kQW2 = kTW / 4 + 2, kQH2 = kTH / 4 + 2;  // 18 x 10 quarter-res points per channel

__global__ void __launch_bounds__(kThreads, 3)
warp2_tma_kernel(const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_a,
                 const float* __restrict__ xb, const float* __restrict__ xa, const float* __restrict__ flow_hat,
                 const float* __restrict__ flow_ab, const float* __restrict__ flow_ba,
                 const float* __restrict__ tab_x, const float* __restrict__ tab_y, float* __restrict__ out,
                 float* __restrict__ flows_out, int h4, int w4, WarpGeom g) {
  extern __shared__ uint8_t smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
  __shared__ float s_q[4][kQH2][kQW2];
  __shared__ int s_red[kThreads / 32][4];
  __shared__ int s_box[3];
  __shared__ __align__(8) uint64_t s_bar;

  const int tid = threadIdx.x;
  const int tx = tid & (kTW - 1), ty = tid >> 6;
  const int bx = blockIdx.x * kTW, by = blockIdx.y * kTH;
  const int x = bx + tx;
  const int n = blockIdx.z;
  const int HW = g.H * g.W;
  const int hh = g.H / 4, ww = g.W / 4;
  const int q = h4 * w4;
  const bool xin = x < g.W;
  const int xc = xin ? x : g.W - 1;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // quarter-res flow = mv x_hat chunk + linear-motion prior (one rounded add each, as torch does)
  const int qx0 = up4_index(bx, ww).i0, qy0 = up4_index(by, hh).i0;
  for (int e = tid; e < 4 * kQH2 * kQW2; e += kThreads) {
    const int ch = e / (kQH2 * kQW2), r = (e / kQW2) % kQH2, c = e % kQW2;
    const int sy = min(qy0 + r, hh - 1), sx = min(qx0 + c, ww - 1);
    const float* pri = (ch < 2 ? flow_ab : flow_ba) + ((int64_t)n * 2 + (ch & 1)) * q;
    const float* hat = flow_hat + ((int64_t)n * 4 + ch) * q;
    s_q[ch][r][c] = __fadd_rn(__ldg(hat + sy * w4 + sx), __ldg(pri + sy * w4 + sx));
  }
  __syncthreads();

  const Up4 ux = up4_index(xc, ww);
  const int cx0 = ux.i0 - qx0, cx1 = ux.i1 - qx0;
  const float txv = __ldg(tab_x + xc);
  float* op = out + (int64_t)n * 6 * HW;

  uint32_t phase = 0;
#pragma unroll 1
  for (int dir = 0; dir < 2; ++dir) {
    float ix[kPX], iy[kPX];
    int mnx = 0x7fffffff, mxx = -0x7fffffff - 1, mny = 0x7fffffff, mxy = -0x7fffffff - 1;
#pragma unroll
    for (int k = 0; k < kPX; ++k) {
      const int yr = by + ty + 4 * k;
      const int y = min(yr, g.H - 1);
      const Up4 uy = up4_index(y, hh);
      const int cy0 = uy.i0 - qy0, cy1 = uy.i1 - qy0;
      const float u = up4_value(uy, ux, s_q[2 * dir][cy0][cx0], s_q[2 * dir][cy0][cx1], s_q[2 * dir][cy1][cx0],
                                s_q[2 * dir][cy1][cx1], 0);
      const float v = up4_value(uy, ux, s_q[2 * dir + 1][cy0][cx0], s_q[2 * dir + 1][cy0][cx1],
                                s_q[2 * dir + 1][cy1][cx0], s_q[2 * dir + 1][cy1][cx1], 0);
      if (flows_out != nullptr && xin && yr < g.H) {
        float* fo = flows_out + ((int64_t)n * 4 + dir * 2) * HW + yr * g.W + x;
        fo[0] = u;
        fo[HW] = v;
      }
      coords<B200VC_WARP_LHBDC, true>(g, xc, y, u, v, txv, __ldg(tab_y + y), ix[k], iy[k]);
      const int x0 = (int)floorf(ix[k]), y0 = (int)floorf(iy[k]);
      mnx = min(mnx, x0); mxx = max(mxx, x0);
      mny = min(mny, y0); mxy = max(mxy, y0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
      mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if ((tid & 31) == 0) {
      s_red[tid >> 5][0] = mnx; s_red[tid >> 5][1] = mxx; s_red[tid >> 5][2] = mny; s_red[tid >> 5][3] = mxy;
    }
    __syncthreads();  // also: every thread has finished reading the box of the previous direction
    if (tid == 0) {
      int a = s_red[0][0], b = s_red[0][1], c = s_red[0][2], d = s_red[0][3];
      for (int w = 1; w < kThreads / 32; ++w) {
        a = min(a, s_red[w][0]); b = max(b, s_red[w][1]); c = min(c, s_red[w][2]); d = max(d, s_red[w][3]);
      }
      a &= ~3;  // TMA: the innermost start coordinate must sit on a 16-byte boundary (unaligned starts fault)
      const bool fits = (b + 1 - a + 1 <= kBW) && (d + 1 - c + 1 <= kBH);
      s_box[0] = a; s_box[1] = c; s_box[2] = fits ? 1 : 0;
      if (fits) {
        const uint32_t bar = smem_u32(&s_bar);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kBoxBytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(dir == 0 ? &map_b : &map_a)), "r"(bar), "r"(a),
              "r"(c), "r"(n * 3)
            : "memory");
      }
    }
    __syncthreads();
    const int bx0 = s_box[0], by0 = s_box[1];
    const bool staged = s_box[2] != 0;
    if (staged) {
      const uint32_t bar = smem_u32(&s_bar);
      uint32_t done = 0;
      for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(phase) : "memory");
        if (spins > (1u << 26)) __trap();
      }
      phase ^= 1u;  // a direction that fell back to gathers never used the barrier: parity follows the copies, not `dir`
    }
    const float* ip = (dir == 0 ? xb : xa) + (int64_t)n * 3 * HW;
#pragma unroll
    for (int k = 0; k < kPX; ++k) {
      const int y = by + ty + 4 * k;
      float r[3];
      if (staged) {
        const float fx = floorf(ix[k]), fy = floorf(iy[k]);
        const int x0 = (int)fx, y0 = (int)fy;
        const float dx1 = __fsub_rn((float)(x0 + 1), ix[k]), dx0 = __fsub_rn(ix[k], (float)x0);
        const float dy1 = __fsub_rn((float)(y0 + 1), iy[k]), dy0 = __fsub_rn(iy[k], (float)y0);
        const float w00 = __fmul_rn(dx1, dy1), w01 = __fmul_rn(dx0, dy1);
        const float w10 = __fmul_rn(dx1, dy0), w11 = __fmul_rn(dx0, dy0);
        const float* t0 = tile + (y0 - by0) * kBW + (x0 - bx0);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float* t = t0 + c * (kBH * kBW);
          float acc = __fmaf_rn(t[0], w00, 0.f);
          acc = __fmaf_rn(t[1], w01, acc);
          acc = __fmaf_rn(t[kBW], w10, acc);
          r[c] = __fmaf_rn(t[kBW + 1], w11, acc);
        }
      } else {
        const Taps t = make_taps<true>(ix[k], iy[k], g.H, g.W);
#pragma unroll
        for (int c = 0; c < 3; ++c) r[c] = sample<true>(ip + (int64_t)c * HW, t);
      }
      if (xin && y < g.H) {
#pragma unroll
        for (int c = 0; c < 3; ++c) op[(int64_t)(dir * 3 + c) * HW + y * g.W + x] = r[c];
      }
    }
  }
}

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

static bool make_img_map(CUtensorMap* map, const float* img, int N, int H, int W) {
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr) return false;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N * 3};
  cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  cuuint32_t box[3] = {kBW, kBH, 3};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(img), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace wt

// Returns B200VC_EUNSUPPORTED when the shape / layout does not qualify (the caller then uses the gather kernel).
int launch_warp_tma(const float* img, int64_t img_bs, const float* flow, const float* tab_x, const float* tab_y,
                    float* out, int64_t out_bs, int N, int C, int H, int W, const WarpGeom& g, cudaStream_t st) {
  using namespace wt;
  static const int enabled = []() {
    const char* e = getenv("B200VC_WARP_TMA");
    return e ? atoi(e) : 1;
  }();
  if (!enabled || C != 3 || img_bs != (int64_t)3 * H * W || W % 4 != 0 || (int64_t)H * W < 128 * 128 ||
      (reinterpret_cast<uintptr_t>(img) & 15u) != 0 || (int64_t)N * 3 >= (1 << 30))
    return B200VC_EUNSUPPORTED;
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr) return B200VC_EUNSUPPORTED;
  CUtensorMap map;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N * 3};
  cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  cuuint32_t box[3] = {kBW, kBH, 3};
  cuuint32_t estr[3] = {1, 1, 1};
  if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(img), dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return B200VC_EUNSUPPORTED;
  static std::atomic<bool> configured[64][3];  // zero-initialised; idempotent set-up, safe under concurrent hosts
  int dev = 0;
  cudaGetDevice(&dev);
  dim3 grid((W + kTW - 1) / kTW, (H + kTH - 1) / kTH, N);
  WarpGeom gd = g;
  static const int unaligned = []() {
    const char* e = getenv("B200VC_WARP_TMA_UNALIGNED");
    return e ? atoi(e) : 0;
  }();
  gd.arith = unaligned;  // the TMA kernel always uses the production arithmetic; the field is reused as a debug switch
#define B200VC_WT_LAUNCH(V)                                                                                     \
  do {                                                                                                          \
    if (dev >= 0 && dev < 64 && !configured[dev][V]) {                                                          \
      if (cudaFuncSetAttribute(warp_tma_kernel<V, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) !=  \
          cudaSuccess) {                                                                                        \
        (void)cudaGetLastError();                                                                               \
        return B200VC_EUNSUPPORTED;                                                                             \
      }                                                                                                         \
      configured[dev][V] = true;                                                                                \
    }                                                                                                           \
    warp_tma_kernel<V, false><<<grid, kThreads, kSmemBytes, st>>>(map, img, flow, tab_x, tab_y, out, out_bs, gd, SpyArgs{});  \
  } while (0)
  if (g.variant == B200VC_WARP_LHBDC) B200VC_WT_LAUNCH(0);
  else if (g.variant == B200VC_WARP_FLEX) B200VC_WT_LAUNCH(1);
  else B200VC_WT_LAUNCH(2);
#undef B200VC_WT_LAUNCH
  return check_launch("warp_f32(tma)");
}

// SPyNet level through the staged kernel; B200VC_EUNSUPPORTED when the level does not qualify (spynet.cu then uses
// its gather kernel).  Upsampled flows are smooth by construction, so the footprint test practically always passes.
int launch_spynet_level_tma(const float* first, int64_t first_bs, const float* second, int64_t second_bs,
                            const float* flow_prev, const float* tab_x, const float* tab_y, float* feat, int N, int H,
                            int W, int hp, int wp, float sy, float sx, const WarpGeom& g, cudaStream_t st) {
  using namespace wt;
  static const int enabled = []() {
    const char* e = getenv("B200VC_SPYNET_TMA");
    return e ? atoi(e) : 1;
  }();
  // measured on B200 (tools/spynet_time.py): staged wins from ~2 Mpx per launch (N=4 1088x1920: 105 vs 121 us);
  // below that the launch is too short for the box round trip to pay and the gather kernel is equal or better
  if (!enabled || second_bs != (int64_t)3 * H * W || W % 4 != 0 || (int64_t)N * H * W < (1 << 21) ||
      (int64_t)H * W < 128 * 128 || (reinterpret_cast<uintptr_t>(second) & 15u) != 0 || (int64_t)N * 3 >= (1 << 30))
    return B200VC_EUNSUPPORTED;
  CUtensorMap map;
  if (!make_img_map(&map, second, N, H, W)) return B200VC_EUNSUPPORTED;
  static std::atomic<bool> configured[64];  // zero-initialised; idempotent set-up, safe under concurrent hosts
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    if (cudaFuncSetAttribute(warp_tma_kernel<B200VC_WARP_LHBDC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kSmemBytes) != cudaSuccess) {
      (void)cudaGetLastError();
      return B200VC_EUNSUPPORTED;
    }
    configured[dev] = true;
  }
  dim3 grid((W + kTW - 1) / kTW, (H + kTH - 1) / kTH, N);
  WarpGeom gd = g;
  gd.arith = 0;
  SpyArgs spy{first, first_bs, hp, wp, sy, sx};
  warp_tma_kernel<B200VC_WARP_LHBDC, true><<<grid, kThreads, kSmemBytes, st>>>(map, second, flow_prev, tab_x, tab_y,
                                                                             feat, (int64_t)8 * H * W, gd, spy);
  return check_launch("spynet_level_f32(tma)");
}

int launch_warp2_tma(const float* xb, const float* xa, const float* flow_hat, const float* flow_ab,
                     const float* flow_ba, const float* tab_x, const float* tab_y, float* out, float* flows_out, int N,
                     int H, int W, int h4, int w4, const WarpGeom& g, cudaStream_t st) {
  using namespace wt;
  // opt-in: on this pool's B200 the staged fused kernel measured no faster than the gather version on the bench's
  // (rough, random-weight) flows (0.306 vs 0.294 ms/step); it is kept, bit-exact and tested, for smooth flows
  static const int enabled = []() {
    const char* e = getenv("B200VC_WARP2_TMA");
    return e ? atoi(e) : 0;
  }();
  if (!enabled || W % 4 != 0 || (int64_t)H * W < 128 * 128 || (int64_t)N * 3 >= (1 << 30) ||
      ((reinterpret_cast<uintptr_t>(xb) | reinterpret_cast<uintptr_t>(xa)) & 15u) != 0)
    return B200VC_EUNSUPPORTED;
  CUtensorMap map_b, map_a;
  if (!make_img_map(&map_b, xb, N, H, W) || !make_img_map(&map_a, xa, N, H, W)) return B200VC_EUNSUPPORTED;
  static std::atomic<bool> configured[64];  // zero-initialised; idempotent set-up, safe under concurrent hosts
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    if (cudaFuncSetAttribute(warp2_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) {
      (void)cudaGetLastError();
      return B200VC_EUNSUPPORTED;
    }
    configured[dev] = true;
  }
  dim3 grid((W + kTW - 1) / kTW, (H + kTH - 1) / kTH, N);
  warp2_tma_kernel<<<grid, kThreads, kSmemBytes, st>>>(map_b, map_a, xb, xa, flow_hat, flow_ab, flow_ba, tab_x, tab_y,
                                                        out, flows_out, h4, w4, g);
  return check_launch("warp2_lhbdc_f32(tma)");
}

}  // namespace b200vc
