// K-WARP / K-WARP2: flow-driven bilinear backward warp.
//
// Replaces torch.grid_sample + the flow-normalisation glue of
//   LHBDC/model/m.py:111-126, LHBDC/model/flow.py:15-25            (WARP_LHBDC)
//   Flex-Rate.../b_model/b_model.py:99-112                         (WARP_FLEX)
//   ICIP2024/src/model/m.py:262-282, OJSP2025/video_model.py:668-676 (WARP_AC1)
// and, for warp2, LHBDC/model/m.py:55-63 (x4 flow upsample + both warps + concat).
//
// HBM-bound gather.  Layout NCHW planar fp32 (the reference layout).  One thread owns one output pixel of a
// 32 x 8 CTA tile: a warp's flow loads and plane stores are single fully-coalesced 128-byte transactions, and
// its four bilinear taps per plane touch 1-2 L1 lines for smooth flows (a 4-pixel-per-thread layout was
// measured 3x slower: 16-byte lane stride => 4-5 L1 wavefronts per gather, and 98-143 registers).
// Algorithmic bytes: (2C+2)*4 per output pixel (C=3: 32 B/px); fused warp2: 50 B/px.
//
// The coordinate path reproduces ATen's CUDA arithmetic operation by operation:
// every step is pinned with __f*_rn intrinsics so nvcc cannot re-associate or contract differently.
#include <stdlib.h>

#include "common.cuh"
#include "warp_common.cuh"

namespace b200vc {

int launch_warp2_tma(const float* xb, const float* xa, const float* flow_hat, const float* flow_ab,
                     const float* flow_ba, const float* tab_x, const float* tab_y, float* out, float* flows_out, int N,
                     int H, int W, int h4, int w4, const WarpGeom& g, cudaStream_t st);  // warp_tma.cu
int launch_warp2_v2(const float* xb, const float* xa, const float* flow_hat, const float* flow_ab,
                    const float* flow_ba, const float* tab_x, const float* tab_y, float* out, float* flows_out, int N,
                    int H, int W, int h4, int w4, const WarpGeom& g, cudaStream_t st);  // warp2.cu
int launch_warp2_staged(const float* xb, const float* xa, const float* flow_hat, const float* flow_ab,
                        const float* flow_ba, const float* tab_x, const float* tab_y, float* out, float* flows_out,
                        int N, int H, int W, int h4, int w4, const WarpGeom& g, cudaStream_t st);  // warp2.cu
int launch_warp_tma(const float* img, int64_t img_bs, const float* flow, const float* tab_x, const float* tab_y,
                    float* out, int64_t out_bs, int N, int C, int H, int W, const WarpGeom& g, cudaStream_t st);  // warp_tma.cu

// CT > 0: channel count known at compile time (planes unrolled, all gathers of a pixel in flight together).
// PX: output pixels per thread (rows y, y+8, ...): the kernel is latency-bound (two dependent DRAM round trips per
// pixel: flow, then taps), so each thread keeps PX independent chains in flight.
template <int CT, int VARIANT, bool ARITH0, int PX>
__global__ void __launch_bounds__(kWarpThreads)
warp_kernel(const float* __restrict__ img, int64_t img_bs, const float* __restrict__ flow,
            const float* __restrict__ tab_x, const float* __restrict__ tab_y, float* __restrict__ out,
            int64_t out_bs, int C, WarpGeom g) {
  constexpr bool BORDER = (VARIANT != B200VC_WARP_FLEX);
  constexpr int kRows = kWarpThreads / 32;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y0 = blockIdx.y * (kRows * PX) + (threadIdx.x >> 5);
  const int n = blockIdx.z;
  if (x >= g.W) return;
  const int HW = g.H * g.W;
  const float* fbase = flow + (int64_t)n * 2 * HW;
  float u[PX], v[PX];
#pragma unroll
  for (int k = 0; k < PX; ++k) {
    const int y = y0 + k * kRows;
    const int o = min(y, g.H - 1) * g.W + x;
    u[k] = __ldg(fbase + o);
    v[k] = __ldg(fbase + HW + o);
  }
  const float tx = BORDER ? __ldg(tab_x + x) : 0.f;
  Taps t[PX];
#pragma unroll
  for (int k = 0; k < PX; ++k) {
    const int y = min(y0 + k * kRows, g.H - 1);
    const float ty = BORDER ? __ldg(tab_y + y) : 0.f;
    float ix, iy;
    coords<VARIANT, ARITH0>(g, x, y, u[k], v[k], tx, ty, ix, iy);
    t[k] = make_taps<BORDER>(ix, iy, g.H, g.W);
  }
  const float* ip = img + (int64_t)n * img_bs;
  float* op = out + (int64_t)n * out_bs;
  if (CT > 0) {
    float r[PX][CT > 0 ? CT : 1];
#pragma unroll
    for (int k = 0; k < PX; ++k)
#pragma unroll
      for (int c = 0; c < CT; ++c) r[k][c] = sample<BORDER>(ip + (int64_t)c * HW, t[k]);
#pragma unroll
    for (int k = 0; k < PX; ++k) {
      const int y = y0 + k * kRows;
      if (y < g.H) {
#pragma unroll
        for (int c = 0; c < CT; ++c) op[(int64_t)c * HW + y * g.W + x] = r[k][c];
      }
    }
  } else {
    const int y = y0;  // PX == 1 in the generic-C instantiation
    if (y < g.H) {
#pragma unroll 4
      for (int c = 0; c < C; ++c) op[(int64_t)c * HW + y * g.W + x] = sample<BORDER>(ip + (int64_t)c * HW, t[0]);
    }
  }
}

// Many-channel form (ICIP feature warps, SURVEY 8a W4: [1,64,544,960], [1,96,272,480], [1,128,136,240]): the channel
// loop of the generic kernel is a serial chain of C gathers per thread and leaves small planes with fewer CTAs than
// SMs.  Here grid.z = sample x channel chunk and a thread owns one pixel of CC channels: all 4*CC taps in flight, the
// coordinate chain (cheap next to 4*CC gathers) recomputed per chunk.
template <int VARIANT, int CC>
__global__ void __launch_bounds__(kWarpThreads)
warp_mc_kernel(const float* __restrict__ img, int64_t img_bs, const float* __restrict__ flow,
               const float* __restrict__ tab_x, const float* __restrict__ tab_y, float* __restrict__ out,
               int64_t out_bs, int C, int chunks, WarpGeom g) {
  constexpr bool BORDER = (VARIANT != B200VC_WARP_FLEX);
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * (kWarpThreads / 32) + (threadIdx.x >> 5);
  const int n = blockIdx.z / chunks, c0 = (blockIdx.z % chunks) * CC;
  if (x >= g.W || y >= g.H) return;
  const int HW = g.H * g.W;
  const int o = y * g.W + x;
  const float* f = flow + (int64_t)n * 2 * HW + o;
  float ix, iy;
  coords<VARIANT, true>(g, x, y, __ldg(f), __ldg(f + HW), BORDER ? __ldg(tab_x + x) : 0.f,
                        BORDER ? __ldg(tab_y + y) : 0.f, ix, iy);
  const Taps t = make_taps<BORDER>(ix, iy, g.H, g.W);
  const float* ip = img + (int64_t)n * img_bs + (int64_t)c0 * HW;
  float* op = out + (int64_t)n * out_bs + (int64_t)c0 * HW + o;
  float r[CC];
  if (c0 + CC <= C) {
#pragma unroll
    for (int c = 0; c < CC; ++c) r[c] = sample<BORDER>(ip + (int64_t)c * HW, t);
#pragma unroll
    for (int c = 0; c < CC; ++c) op[(int64_t)c * HW] = r[c];
  } else {
    for (int c = 0; c0 + c < C; ++c) op[(int64_t)c * HW] = sample<BORDER>(ip + (int64_t)c * HW, t);
  }
}

// ---------------------------------------------------------------------------------------------------
// warp2: LHBDC flow glue + two warps + concat in one pass.
// ATen upsample_bilinear2d (align_corners=False, scale_factor=4 => rscale = 0.25):
//   src = max(0.25*(dst+0.5) - 0.5, 0); i0 = (int)src; ip = i0 < in-1; l1 = src - i0; l0 = 1 - l1
//   val = l0y*(l0x*a + l1x*b) + l1y*(l0x*c + l1x*d)
// CTA = 32 x 8 output pixels.  The quarter-resolution flow (mv x_hat + linear-motion prior, m.py:56,58) that the
// tile's x4 upsample touches -- at most 10 x 4 points x 4 channels -- is summed once into shared memory; every
// pixel then reads its four corners from there (warp-wide broadcasts) instead of 32 global loads.
constexpr int kQW = 10, kQH = 4;
template <bool ARITH0>
__global__ void __launch_bounds__(kWarpThreads)
warp2_lhbdc_kernel(const float* __restrict__ xb, const float* __restrict__ xa,
                   const float* __restrict__ flow_hat, const float* __restrict__ flow_ab,
                   const float* __restrict__ flow_ba, const float* __restrict__ tab_x,
                   const float* __restrict__ tab_y, float* __restrict__ out, float* __restrict__ flows_out,
                   int h4, int w4, WarpGeom g) {
  __shared__ float s_q[4][kQH][kQW];
  const int n = blockIdx.z;
  const int bx = blockIdx.x * 32, by = blockIdx.y * (kWarpThreads / 32);
  const int hh = g.H / 4, ww = g.W / 4;
  const int q = h4 * w4;
  // first quarter-res column / row the tile can touch (i0 of its first pixel)
  const int qx0 = up4_index(bx, ww).i0, qy0 = up4_index(by, hh).i0;
  if (threadIdx.x < 4 * kQH * kQW) {
    const int ch = threadIdx.x / (kQH * kQW), r = (threadIdx.x / kQW) % kQH, c = threadIdx.x % kQW;
    const int sy = min(qy0 + r, hh - 1), sx = min(qx0 + c, ww - 1);
    // ch 0,1 = flow_cb (x, y) = x_hat[0:2] + flow_ab ; ch 2,3 = flow_ca = x_hat[2:4] + flow_ba
    const float* pri = (ch < 2 ? flow_ab : flow_ba) + ((int64_t)n * 2 + (ch & 1)) * q;
    const float* hat = flow_hat + ((int64_t)n * 4 + ch) * q;
    s_q[ch][r][c] = __fadd_rn(__ldg(hat + sy * w4 + sx), __ldg(pri + sy * w4 + sx));
  }
  __syncthreads();
  const int x = bx + (threadIdx.x & 31), y = by + (threadIdx.x >> 5);
  if (x >= g.W || y >= g.H) return;
  const int HW = g.H * g.W;
  const int o = y * g.W + x;
  const Up4 ux = up4_index(x, ww), uy = up4_index(y, hh);
  const int cx0 = ux.i0 - qx0, cx1 = ux.i1 - qx0, cy0 = uy.i0 - qy0, cy1 = uy.i1 - qy0;
  const float tx = __ldg(tab_x + x), ty = __ldg(tab_y + y);
  const int arith = ARITH0 ? 0 : g.arith;
  float* op = out + (int64_t)n * 6 * HW + o;
#pragma unroll
  for (int dir = 0; dir < 2; ++dir) {
    const float u = up4_value(uy, ux, s_q[2 * dir][cy0][cx0], s_q[2 * dir][cy0][cx1], s_q[2 * dir][cy1][cx0],
                              s_q[2 * dir][cy1][cx1], arith);
    const float v = up4_value(uy, ux, s_q[2 * dir + 1][cy0][cx0], s_q[2 * dir + 1][cy0][cx1],
                              s_q[2 * dir + 1][cy1][cx0], s_q[2 * dir + 1][cy1][cx1], arith);
    if (flows_out != nullptr) {
      float* fo = flows_out + ((int64_t)n * 4 + dir * 2) * HW + o;
      fo[0] = u;
      fo[HW] = v;
    }
    float ix, iy;
    coords<B200VC_WARP_LHBDC, ARITH0>(g, x, y, u, v, tx, ty, ix, iy);
    const Taps t = make_taps<true>(ix, iy, g.H, g.W);
    const float* ip = (dir == 0 ? xb : xa) + (int64_t)n * 3 * HW;
    float r[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) r[c] = sample<true>(ip + (int64_t)c * HW, t);
#pragma unroll
    for (int c = 0; c < 3; ++c) op[(int64_t)(dir * 3 + c) * HW] = r[c];
  }
}

// ---------------------------------------------------------------------------------------------------
// Search form (ICIP2024/src/opt_helpers.py:23-51, OJSP2025/video_model.py:621-666): for every candidate
// down-ratio the reference warps both references, blends 0.5/0.5, clamps and takes the MSE against the current
// frame -- four full-frame tensors written and re-read per candidate.  Here: both warps + blend + clamp + squared
// error in one pass, nothing but per-CTA fp64 partials written (52 B/px algorithmic instead of 156+).
template <int VARIANT>
__global__ void __launch_bounds__(kWarpThreads)
warp2_half_sse_kernel(const float* __restrict__ x1, const float* __restrict__ x2, const float* __restrict__ flow1,
                      const float* __restrict__ flow2, const float* __restrict__ x_cur,
                      const float* __restrict__ tab_x, const float* __restrict__ tab_y, float* __restrict__ pred_out,
                      double* __restrict__ partials, WarpGeom g, Finish fin) {
  constexpr bool BORDER = (VARIANT != B200VC_WARP_FLEX);
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * (kWarpThreads / 32) + (threadIdx.x >> 5);
  const int n = blockIdx.z;
  const int HW = g.H * g.W;
  float sse = 0.f;
  if (x < g.W && y < g.H) {
    const int o = y * g.W + x;
    const float tx = BORDER ? __ldg(tab_x + x) : 0.f, ty = BORDER ? __ldg(tab_y + y) : 0.f;
    float r[2][3];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float* f = (k == 0 ? flow1 : flow2) + (int64_t)n * 2 * HW + o;
      float ix, iy;
      coords<VARIANT, true>(g, x, y, __ldg(f), __ldg(f + HW), tx, ty, ix, iy);
      const Taps t = make_taps<BORDER>(ix, iy, g.H, g.W);
      const float* ip = (k == 0 ? x1 : x2) + (int64_t)n * 3 * HW;
#pragma unroll
      for (int c = 0; c < 3; ++c) r[k][c] = sample<BORDER>(ip + (int64_t)c * HW, t);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // mask*wref1 + (1-mask)*wref2 with mask = 0.5: both products exact, one rounding
      const float pred = __fadd_rn(__fmul_rn(0.5f, r[0][c]), __fmul_rn(0.5f, r[1][c]));
      if (pred_out != nullptr) pred_out[((int64_t)n * 3 + c) * HW + o] = pred;
      const float d = __fsub_rn(fminf(fmaxf(pred, 0.f), 1.f), __ldg(x_cur + ((int64_t)n * 3 + c) * HW + o));
      sse = __fmaf_rn(d, d, sse);
    }
  }
  const double tot = block_sum_to_f64<kWarpThreads>(sse);
  const int per = gridDim.x * gridDim.y;
  publish_partial<kWarpThreads>(tot, partials + (int64_t)n * per, blockIdx.y * gridDim.x + blockIdx.x, per, n, fin);
}

// Single-reference search form (OJSP2025/video_model.py:621-666: x_hat = warp(ref_frame, est_mv) ; PSNR(x, x_hat)
// = 10 log10(1 / mean((x - x_hat)^2)) for each of 32 candidate ratios at 2160 x 3840): warp + squared error in one
// pass; the warped frame is written only when asked for (32 B/px algorithmic instead of 56 + the torch MSE chain).
template <int VARIANT>
__global__ void __launch_bounds__(kWarpThreads)
warp_sse_kernel(const float* __restrict__ img, const float* __restrict__ flow, const float* __restrict__ x_cur,
                const float* __restrict__ tab_x, const float* __restrict__ tab_y, float* __restrict__ pred_out,
                double* __restrict__ partials, WarpGeom g, Finish fin) {
  constexpr bool BORDER = (VARIANT != B200VC_WARP_FLEX);
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * (kWarpThreads / 32) + (threadIdx.x >> 5);
  const int n = blockIdx.z;
  const int HW = g.H * g.W;
  float sse = 0.f;
  if (x < g.W && y < g.H) {
    const int o = y * g.W + x;
    const float tx = BORDER ? __ldg(tab_x + x) : 0.f, ty = BORDER ? __ldg(tab_y + y) : 0.f;
    const float* f = flow + (int64_t)n * 2 * HW + o;
    float ix, iy;
    coords<VARIANT, true>(g, x, y, __ldg(f), __ldg(f + HW), tx, ty, ix, iy);
    const Taps t = make_taps<BORDER>(ix, iy, g.H, g.W);
    const float* ip = img + (int64_t)n * 3 * HW;
    float r[3], c0[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      r[c] = sample<BORDER>(ip + (int64_t)c * HW, t);
      c0[c] = __ldg(x_cur + ((int64_t)n * 3 + c) * HW + o);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (pred_out != nullptr) pred_out[((int64_t)n * 3 + c) * HW + o] = r[c];
      const float d = __fsub_rn(c0[c], r[c]);
      sse = __fmaf_rn(d, d, sse);
    }
  }
  const double tot = block_sum_to_f64<kWarpThreads>(sse);
  const int per = gridDim.x * gridDim.y;
  publish_partial<kWarpThreads>(tot, partials + (int64_t)n * per, blockIdx.y * gridDim.x + blockIdx.x, per, n, fin);
}

}  // namespace b200vc

using namespace b200vc;

extern "C" int b200vc_warp_f32(const float* img, int64_t img_bs, const float* flow, const float* tab_x,
                               const float* tab_y, float* out, int64_t out_bs, int N, int C, int H, int W,
                               int variant, int arith, void* stream) {
  B200VC_REQUIRE(img && flow && out, "warp_f32: null pointer");
  B200VC_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, "warp_f32: bad shape N=%d C=%d H=%d W=%d", N, C, H, W);
  B200VC_REQUIRE(variant >= 0 && variant <= 2, "warp_f32: unknown variant %d", variant);
  B200VC_REQUIRE(variant == B200VC_WARP_FLEX || (tab_x && tab_y), "warp_f32: grid tables required");
  B200VC_REQUIRE((int64_t)H * W < (1ll << 31), "warp_f32: plane too large");
  B200VC_REQUIRE(N <= 65535, "warp_f32: N too large");
  const WarpGeom g = make_geom(H, W, variant, arith);
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = kWarpThreads / 32;
  dim3 grid((W + 31) / 32, (H + rows - 1) / rows, N);
  if (arith == 0) {
    const int rc = launch_warp_tma(img, img_bs, flow, tab_x, tab_y, out, out_bs, N, C, H, W, g, st);
    if (rc != B200VC_EUNSUPPORTED) return rc;
  }
#define B200VC_WARP_ARGS img, img_bs, flow, tab_x, tab_y, out, out_bs, C, g
  static const int px_env = []() {
    const char* e = getenv("B200VC_WARP_PX");
    return e ? atoi(e) : 0;
  }();
  // pixels per thread: 2 when the plane is large enough to still fill the machine (tuned on B200)
  const int px = (C > 4 || arith != 0) ? 1 : (px_env > 0 ? px_env : ((int64_t)N * H * W >= (1 << 19) ? 2 : 1));
  const int rows_per_cta = rows * px;
  grid = dim3((W + 31) / 32, (H + rows_per_cta - 1) / rows_per_cta, N);
#define B200VC_WARP_LAUNCH_PX(CT, PX)                                                                      \
  do {                                                                                                    \
    if (variant == B200VC_WARP_LHBDC)                                                                     \
      warp_kernel<CT, 0, true, PX><<<grid, kWarpThreads, 0, st>>>(B200VC_WARP_ARGS);                      \
    else if (variant == B200VC_WARP_FLEX)                                                                 \
      warp_kernel<CT, 1, true, PX><<<grid, kWarpThreads, 0, st>>>(B200VC_WARP_ARGS);                      \
    else                                                                                                  \
      warp_kernel<CT, 2, true, PX><<<grid, kWarpThreads, 0, st>>>(B200VC_WARP_ARGS);                      \
  } while (0)
#define B200VC_WARP_LAUNCH(CT)                                                                            \
  do {                                                                                                    \
    if (px >= 4) B200VC_WARP_LAUNCH_PX(CT, 4);                                                            \
    else if (px == 2) B200VC_WARP_LAUNCH_PX(CT, 2);                                                       \
    else B200VC_WARP_LAUNCH_PX(CT, 1);                                                                    \
  } while (0)
  if (arith != 0) {
    if (variant == B200VC_WARP_LHBDC) warp_kernel<0, 0, false, 1><<<grid, kWarpThreads, 0, st>>>(B200VC_WARP_ARGS);
    else if (variant == B200VC_WARP_FLEX) warp_kernel<0, 1, false, 1><<<grid, kWarpThreads, 0, st>>>(B200VC_WARP_ARGS);
    else warp_kernel<0, 2, false, 1><<<grid, kWarpThreads, 0, st>>>(B200VC_WARP_ARGS);
  } else {
    switch (C) {
      case 1: B200VC_WARP_LAUNCH(1); break;
      case 2: B200VC_WARP_LAUNCH(2); break;
      case 3: B200VC_WARP_LAUNCH(3); break;
      case 4: B200VC_WARP_LAUNCH(4); break;
      default: {
        constexpr int CC = 8;  // 16 measured 5-8 % slower (tools/warp_feat_time.py)
        const int chunks = (C + CC - 1) / CC;
        if ((int64_t)N * chunks <= 65535) {
          const dim3 gmc((W + 31) / 32, (H + rows - 1) / rows, N * chunks);
          if (variant == B200VC_WARP_LHBDC)
            warp_mc_kernel<0, CC><<<gmc, kWarpThreads, 0, st>>>(img, img_bs, flow, tab_x, tab_y, out, out_bs, C, chunks, g);
          else if (variant == B200VC_WARP_FLEX)
            warp_mc_kernel<1, CC><<<gmc, kWarpThreads, 0, st>>>(img, img_bs, flow, tab_x, tab_y, out, out_bs, C, chunks, g);
          else
            warp_mc_kernel<2, CC><<<gmc, kWarpThreads, 0, st>>>(img, img_bs, flow, tab_x, tab_y, out, out_bs, C, chunks, g);
        } else {
          B200VC_WARP_LAUNCH_PX(0, 1);
        }
        break;
      }
    }
  }
#undef B200VC_WARP_LAUNCH
#undef B200VC_WARP_LAUNCH_PX
#undef B200VC_WARP_ARGS
  return check_launch("warp_f32");
}

extern "C" int b200vc_warp2_lhbdc_f32(const float* x_before, const float* x_after, const float* flow_hat,
                                      const float* flow_ab, const float* flow_ba, const float* tab_x,
                                      const float* tab_y, float* out, float* flows_out, int N, int H, int W,
                                      int h4, int w4, int arith, void* stream) {
  B200VC_REQUIRE(x_before && x_after && flow_hat && flow_ab && flow_ba && tab_x && tab_y && out,
                 "warp2_lhbdc_f32: null pointer");
  B200VC_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0, "warp2_lhbdc_f32: bad shape");
  B200VC_REQUIRE(H % 4 == 0 && W % 4 == 0, "warp2_lhbdc_f32: H, W must be multiples of 4 (got %d x %d)", H, W);
  B200VC_REQUIRE(h4 >= H / 4 && w4 >= W / 4, "warp2_lhbdc_f32: quarter-res tensors smaller than the crop");
  B200VC_REQUIRE((int64_t)H * W < (1ll << 31), "warp2_lhbdc_f32: plane too large");
  const WarpGeom g = make_geom(H, W, B200VC_WARP_LHBDC, arith);
  const int rows = kWarpThreads / 32;
  dim3 grid((W + 31) / 32, (H + rows - 1) / rows, N);
  if (arith == 0) {
    const int rc = launch_warp2_tma(x_before, x_after, flow_hat, flow_ab, flow_ba, tab_x, tab_y, out, flows_out, N, H, W,
                                    h4, w4, g, (cudaStream_t)stream);
    if (rc != B200VC_EUNSUPPORTED) return rc;
    const int rc3 = launch_warp2_staged(x_before, x_after, flow_hat, flow_ab, flow_ba, tab_x, tab_y, out, flows_out, N, H,
                                        W, h4, w4, g, (cudaStream_t)stream);
    if (rc3 != B200VC_EUNSUPPORTED) return rc3;
    const int rc2 = launch_warp2_v2(x_before, x_after, flow_hat, flow_ab, flow_ba, tab_x, tab_y, out, flows_out, N, H, W,
                                    h4, w4, g, (cudaStream_t)stream);
    if (rc2 != B200VC_EUNSUPPORTED) return rc2;
  }
  if (arith == 0)
    warp2_lhbdc_kernel<true><<<grid, kWarpThreads, 0, (cudaStream_t)stream>>>(
        x_before, x_after, flow_hat, flow_ab, flow_ba, tab_x, tab_y, out, flows_out, h4, w4, g);
  else
    warp2_lhbdc_kernel<false><<<grid, kWarpThreads, 0, (cudaStream_t)stream>>>(
        x_before, x_after, flow_hat, flow_ab, flow_ba, tab_x, tab_y, out, flows_out, h4, w4, g);
  return check_launch("warp2_lhbdc_f32");
}

extern "C" int b200vc_warp2_half_sse_blocks(int H, int W) {
  if (H <= 0 || W <= 0) return 0;
  return ((W + 31) / 32) * ((H + kWarpThreads / 32 - 1) / (kWarpThreads / 32));
}

extern "C" int b200vc_warp2_half_sse_f32(const float* x1, const float* x2, const float* flow1, const float* flow2,
                                         const float* x_cur, const float* tab_x, const float* tab_y, float* pred,
                                         double* partials, double* totals, int32_t* counters, int N, int H, int W,
                                         int variant, void* stream) {
  B200VC_REQUIRE(!totals || counters, "warp2_half_sse_f32: totals need counters");
  B200VC_REQUIRE(x1 && x2 && flow1 && flow2 && x_cur && partials, "warp2_half_sse_f32: null pointer");
  B200VC_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0, "warp2_half_sse_f32: bad shape");
  B200VC_REQUIRE(variant >= 0 && variant <= 2, "warp2_half_sse_f32: unknown variant %d", variant);
  B200VC_REQUIRE(variant == B200VC_WARP_FLEX || (tab_x && tab_y), "warp2_half_sse_f32: grid tables required");
  B200VC_REQUIRE((int64_t)H * W < (1ll << 31), "warp2_half_sse_f32: plane too large");
  const WarpGeom g = make_geom(H, W, variant, 0);
  const int rows = kWarpThreads / 32;
  dim3 grid((W + 31) / 32, (H + rows - 1) / rows, N);
  cudaStream_t st = (cudaStream_t)stream;
  if (variant == B200VC_WARP_LHBDC)
    warp2_half_sse_kernel<0><<<grid, kWarpThreads, 0, st>>>(x1, x2, flow1, flow2, x_cur, tab_x, tab_y, pred, partials, g, Finish{totals, counters});
  else if (variant == B200VC_WARP_FLEX)
    warp2_half_sse_kernel<1><<<grid, kWarpThreads, 0, st>>>(x1, x2, flow1, flow2, x_cur, tab_x, tab_y, pred, partials, g, Finish{totals, counters});
  else
    warp2_half_sse_kernel<2><<<grid, kWarpThreads, 0, st>>>(x1, x2, flow1, flow2, x_cur, tab_x, tab_y, pred, partials, g, Finish{totals, counters});
  return check_launch("warp2_half_sse_f32");
}

extern "C" int b200vc_warp_sse_f32(const float* img, const float* flow, const float* x_cur, const float* tab_x,
                                   const float* tab_y, float* pred, double* partials, double* totals,
                                   int32_t* counters, int N, int H, int W, int variant, void* stream) {
  B200VC_REQUIRE(!totals || counters, "warp_sse_f32: totals need counters");
  B200VC_REQUIRE(img && flow && x_cur && partials, "warp_sse_f32: null pointer");
  B200VC_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0, "warp_sse_f32: bad shape");
  B200VC_REQUIRE(variant >= 0 && variant <= 2, "warp_sse_f32: unknown variant %d", variant);
  B200VC_REQUIRE(variant == B200VC_WARP_FLEX || (tab_x && tab_y), "warp_sse_f32: grid tables required");
  B200VC_REQUIRE((int64_t)H * W < (1ll << 31), "warp_sse_f32: plane too large");
  const WarpGeom g = make_geom(H, W, variant, 0);
  const int rows = kWarpThreads / 32;
  dim3 grid((W + 31) / 32, (H + rows - 1) / rows, N);
  cudaStream_t st = (cudaStream_t)stream;
  if (variant == B200VC_WARP_LHBDC)
    warp_sse_kernel<0><<<grid, kWarpThreads, 0, st>>>(img, flow, x_cur, tab_x, tab_y, pred, partials, g, Finish{totals, counters});
  else if (variant == B200VC_WARP_FLEX)
    warp_sse_kernel<1><<<grid, kWarpThreads, 0, st>>>(img, flow, x_cur, tab_x, tab_y, pred, partials, g, Finish{totals, counters});
  else
    warp_sse_kernel<2><<<grid, kWarpThreads, 0, st>>>(img, flow, x_cur, tab_x, tab_y, pred, partials, g, Finish{totals, counters});
  return check_launch("warp_sse_f32");
}
