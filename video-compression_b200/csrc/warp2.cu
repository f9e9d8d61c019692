// K-WARP2 (second generation): LHBDC flow glue + both backward warps + concat in one pass
// (LHBDC/model/m.py:55-63: chunk, + linear-motion prior, crop, nn.Upsample(x4, bilinear), two grid_sample, cat).
//
// Why a rewrite: the first kernel (warp.cu, one pixel per thread, 32 x 8 tile) ran at 38 % of the HBM peak.  ncu:
// ~470 thread instructions and ~48 load/store instructions per pixel; on this part the LSU accepts one warp-level
// LDG/LDS/STG roughly every 1.8 clk per SM, so 48 of them per 32 pixels is a 20 us floor per 1088 x 1920 frame on
// its own, and the 470-instruction chain another 27 us -- against 16 us of HBM time for the 50 algorithmic bytes per
// pixel.  This version attacks both counts while keeping every rounding of the ATen chain (bit-exact, same tests):
//   * the four corners of a pixel's x4-upsample cell come from ONE 128-bit shared load per flow channel (the staging
//     pass stores (r,c), (r,c+1), (r+1,c), (r+1,c+1) of `x_hat + prior` side by side, edge replication = ATen's
//     `i1 = i0 + (i0 < in-1)`): 4 LDS instead of 16;
//   * a thread owns 4 pixels of one row, 32 apart (lanes stay adjacent => every gather / store instruction of a warp
//     still covers one or two 128-byte lines), so the row's upsample weights, grid value and offsets are computed
//     once per 4 pixels;
//   * tap offsets are 32-bit against warp-uniform plane bases; the border variant needs no int-range guard (the
//     clip maps NaN to 0 and everything else into [0, size-1]); `(float)(x0 + 1)` is `floor(ix) + 1`.
// Algorithmic bytes: 2 x 3 planes read + 6 planes written + quarter-resolution flows = 50 B/px.
#include <stdlib.h>

#include "common.cuh"
#include "warp_common.cuh"

namespace b200vc {
namespace w2 {

constexpr int kTX = 128, kTY = 8, kThreads = 256, kPX = 4;
constexpr int kQC = 34;  // quarter-resolution cells a 128-pixel span can start in (33) + 1
constexpr int kQR = 3;   // ... an 8-row span can start in

struct Cell {  // upsample source cell of one output coordinate
  int i0;
  float l0, l1;
};
__device__ __forceinline__ Cell cell_of(int dst) {
  // ATen area_pixel_compute_source_index(scale = 0.25, align_corners = false, cubic = false)
  Cell r;
  float src = __fmaf_rn(0.25f, (float)dst + 0.5f, -0.5f);
  src = src < 0.f ? 0.f : src;
  r.i0 = (int)src;
  r.l1 = __fsub_rn(src, (float)r.i0);
  r.l0 = __fsub_rn(1.f, r.l1);
  return r;
}

template <bool FLOWS>
__global__ void __launch_bounds__(kThreads)
warp2_kernel(const float* __restrict__ xb, const float* __restrict__ xa, const float* __restrict__ flow_hat,
             const float* __restrict__ flow_ab, const float* __restrict__ flow_ba, const float* __restrict__ tab_x,
             const float* __restrict__ tab_y, float* __restrict__ out, float* __restrict__ flows_out, int h4, int w4,
             WarpGeom g) {
  __shared__ float4 s_c[4][kQR][kQC];
  const int n = blockIdx.z;
  const int bx = blockIdx.x * kTX, by = blockIdx.y * kTY;
  const int hh = g.H >> 2, ww = g.W >> 2;
  const int qx0 = max((bx >> 2) - 1, 0), qy0 = max((by >> 2) - 1, 0);
  {
    const int q = h4 * w4;
    for (int e = threadIdx.x; e < 4 * kQR * kQC; e += kThreads) {
      const int ch = e / (kQR * kQC), r = (e / kQC) % kQR, c = e % kQC;
      const int y0 = min(qy0 + r, hh - 1) * w4, y1 = min(qy0 + r + 1, hh - 1) * w4;
      const int x0 = min(qx0 + c, ww - 1), x1 = min(qx0 + c + 1, ww - 1);
      // ch 0,1 = flow_cb (x, y) = x_hat[0:2] + flow_ab ; ch 2,3 = flow_ca = x_hat[2:4] + flow_ba   (m.py:56,58)
      const float* pri = (ch < 2 ? flow_ab : flow_ba) + ((int64_t)n * 2 + (ch & 1)) * q;
      const float* hat = flow_hat + ((int64_t)n * 4 + ch) * q;
      float4 v;
      v.x = __fadd_rn(__ldg(hat + y0 + x0), __ldg(pri + y0 + x0));
      v.y = __fadd_rn(__ldg(hat + y0 + x1), __ldg(pri + y0 + x1));
      v.z = __fadd_rn(__ldg(hat + y1 + x0), __ldg(pri + y1 + x0));
      v.w = __fadd_rn(__ldg(hat + y1 + x1), __ldg(pri + y1 + x1));
      s_c[ch][r][c] = v;
    }
  }
  __syncthreads();
  const int y = by + (threadIdx.x >> 5);
  if (y >= g.H) return;
  const int lane = threadIdx.x & 31;
  const int HW = g.H * g.W;
  const Cell cy = cell_of(y);
  const int ry = cy.i0 - qy0;
  const float ty = __ldg(tab_y + y);
  // per-thread pointers to pixel (y, bx + lane) of every output plane: the 4 pixels of the thread are then
  // immediate offsets (+128 B each), so stores cost no address arithmetic
  const int o0 = y * g.W + bx + lane;
  float* po[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    po[c] = out + ((int64_t)n * 6 + c) * HW + o0;
    asm volatile("" : "+l"(po[c]));  // keep the pointer materialised: its 4 pixels are then immediate offsets
  }
  float* pf[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    pf[c] = FLOWS ? flows_out + ((int64_t)n * 4 + c) * HW + o0 : nullptr;
    if (FLOWS) asm volatile("" : "+l"(pf[c]));
  }
  const float* ptx = tab_x + bx + lane;
  // plane bases kept opaque so that every tap address is ONE `IMAD.WIDE base, offset, 4` (ptxas otherwise folds
  // n, c and the offset into a 64-bit element index: four integer instructions per gather)
  const float* pin[2][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    pin[0][c] = xb + ((int64_t)n * 3 + c) * HW;
    pin[1][c] = xa + ((int64_t)n * 3 + c) * HW;
    asm volatile("" : "+l"(pin[0][c]), "+l"(pin[1][c]));
  }
  const float wmax = (float)(g.W - 1), hmax = (float)(g.H - 1);
  const float fW = (float)g.W, fH = (float)g.H;
  const int W2 = g.W - 2, H2 = g.H - 2;
#pragma unroll
  for (int j = 0; j < kPX; ++j) {
    const int x = bx + lane + 32 * j;
    if (x >= g.W) break;
    const Cell cx = cell_of(x);
    const int rc = cx.i0 - qx0;
    const float tx = __ldg(ptx + 32 * j);
    float uv[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const float4 c = s_c[ch][ry][rc];
      // val = l0y*(l0x*a + l1x*b) + l1y*(l0x*c + l1x*d), contracted as ATen's CUDA kernel is
      const float top = __fmaf_rn(cx.l0, c.x, __fmul_rn(cx.l1, c.y));
      const float bot = __fmaf_rn(cx.l0, c.z, __fmul_rn(cx.l1, c.w));
      uv[ch] = __fmaf_rn(cy.l0, top, __fmul_rn(cy.l1, bot));
    }
    if (FLOWS) {
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) __stcg(pf[ch] + 32 * j, uv[ch]);
    }
    float t[2][3][4], w[2][4];
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      // grid + flow / ((W-1)/2)  ->  grid_sampler unnormalise (align_corners=False) -> border clip   (m.py:121-125)
      const float gx = __fadd_rn(tx, __fmul_rn(uv[2 * dir], g.inv_x));
      const float gy = __fadd_rn(ty, __fmul_rn(uv[2 * dir + 1], g.inv_y));
      float ix = __fmul_rn(__fmaf_rn(__fadd_rn(gx, 1.f), fW, -1.f), 0.5f);
      float iy = __fmul_rn(__fmaf_rn(__fadd_rn(gy, 1.f), fH, -1.f), 0.5f);
      ix = fminf(wmax, fmaxf(ix, 0.f));
      iy = fminf(hmax, fmaxf(iy, 0.f));
      const float fx = floorf(ix), fy = floorf(iy);
      const float dx1 = __fsub_rn(__fadd_rn(fx, 1.f), ix), dx0 = __fsub_rn(ix, fx);
      const float dy1 = __fsub_rn(__fadd_rn(fy, 1.f), iy), dy0 = __fsub_rn(iy, fy);
      const float w00 = __fmul_rn(dx1, dy1), w01 = __fmul_rn(dx0, dy1);
      const float w10 = __fmul_rn(dx1, dy0), w11 = __fmul_rn(dx0, dy0);
      // The 2x2 footprint is always read as (xq, xq+1) x (yq, yq+1) with xq <= W-2, yq <= H-2, so the three
      // neighbours are immediate / one-add offsets of one address.  ATen clamps the "+1" taps instead; they only
      // leave the plane when the clipped coordinate is exactly W-1 (H-1), where their weights are exactly 0: there
      // the footprint is shifted by one and the weights move with it (same products, same accumulation order).
      const int x0 = (int)fx, y0 = (int)fy;
      const bool sx = x0 > W2, sy = y0 > H2;
      const float a00 = sx ? 0.f : w00, a01 = sx ? w00 : w01, a10 = sx ? 0.f : w10, a11 = sx ? w10 : w11;
      w[dir][0] = sy ? 0.f : a00;
      w[dir][1] = sy ? 0.f : a01;
      w[dir][2] = sy ? a00 : a10;
      w[dir][3] = sy ? a01 : a11;
      const int q0 = min(y0, H2) * g.W + min(x0, W2), q1 = q0 + g.W;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* rc0 = pin[dir][c] + q0;
        const float* rc1 = pin[dir][c] + q1;
        t[dir][c][0] = __ldg(rc0);
        t[dir][c][1] = __ldg(rc0 + 1);
        t[dir][c][2] = __ldg(rc1);
        t[dir][c][3] = __ldg(rc1 + 1);
      }
    }
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        // ATen: out_acc = 0; out_acc += v * w per tap in nw, ne, sw, se order (FMA-contracted)
        float acc = __fmaf_rn(t[dir][c][0], w[dir][0], 0.f);
        acc = __fmaf_rn(t[dir][c][1], w[dir][1], acc);
        acc = __fmaf_rn(t[dir][c][2], w[dir][2], acc);
        acc = __fmaf_rn(t[dir][c][3], w[dir][3], acc);
        __stcg(po[dir * 3 + c] + 32 * j, acc);
      }
    }
  }
}

}  // namespace w2

// Production arithmetic only (arith == 0); any H, W that are multiples of 4; planes below 2^31 elements / 6.
int launch_warp2_v2(const float* xb, const float* xa, const float* flow_hat, const float* flow_ab,
                    const float* flow_ba, const float* tab_x, const float* tab_y, float* out, float* flows_out, int N,
                    int H, int W, int h4, int w4, const WarpGeom& g, cudaStream_t st) {
  using namespace w2;
  static const int enabled = []() {
    const char* e = getenv("B200VC_WARP2_V2");
    return e ? atoi(e) : 1;
  }();
  if (!enabled || H < 2 || W < 2) return B200VC_EUNSUPPORTED;
  dim3 grid((W + kTX - 1) / kTX, (H + kTY - 1) / kTY, N);
  if (flows_out != nullptr)
    warp2_kernel<true><<<grid, kThreads, 0, st>>>(xb, xa, flow_hat, flow_ab, flow_ba, tab_x, tab_y, out, flows_out, h4,
                                                   w4, g);
  else
    warp2_kernel<false><<<grid, kThreads, 0, st>>>(xb, xa, flow_hat, flow_ab, flow_ba, tab_x, tab_y, out, flows_out,
                                                    h4, w4, g);
  return check_launch("warp2_lhbdc_f32(v2)");
}

}  // namespace b200vc
