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
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

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

__device__ __forceinline__ float ldg_tap(const float* p) { return __ldg(p); }

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
        t[dir][c][0] = ldg_tap(rc0);
        t[dir][c][1] = ldg_tap(rc0 + 1);
        t[dir][c][2] = ldg_tap(rc1);
        t[dir][c][3] = ldg_tap(rc1 + 1);
      }
    }
    // Scheduling fence that survives ptxas: the accumulators start from a +0.0f that is *data-dependent on every
    // gather of the pixel* (OR of their bits AND a kernel argument that is 0 in this instantiation), so no FMA chain can be placed between
    // the loads -- the SM issues in order, and a consumer scheduled early stalls the issue of the remaining gathers
    // on its scoreboard (first build: 4-6 of 24 taps in flight, 80 % long-scoreboard stalls, 0.37 IPC).
    int any = 0;
#pragma unroll
    for (int dir = 0; dir < 2; ++dir)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        any |= __float_as_int(t[dir][c][0]) | __float_as_int(t[dir][c][1]) | __float_as_int(t[dir][c][2]) |
               __float_as_int(t[dir][c][3]);
    const float zero = __int_as_float(any & g.zero);
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        // ATen: out_acc = 0; out_acc += v * w per tap in nw, ne, sw, se order (FMA-contracted)
        float acc = __fmaf_rn(t[dir][c][0], w[dir][0], zero);
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

// ---------------------------------------------------------------------------------------------------------------
// K-WARP2, TMA-staged form (the default for frames >= 128 x 128): what the gather form above cannot fix is its L1
// traffic -- a warp's gather of 32 neighbouring, unaligned pixels is two L1 wavefronts, 24 of them per pixel, and
// ncu shows both generations of the gather kernel converging on the same ~36 us per 1088 x 1920 frame with the L1
// data pipe as the busiest unit.  Here the reference-frame tile around a CTA's flow footprint arrives by ONE 3-D TMA
// box copy per direction (no LSU traffic, zero fill outside the frame = ATen's skipped taps), and every tap is a
// conflict-free shared-memory read at an immediate offset from one address per pixel.  The older staged kernel
// (warp_tma.cu: warp2_tma_kernel, kept behind B200VC_WARP2_TMA=1) has this structure but ~470 instructions per pixel
// (issue-bound, ncu IPC 2.7); this one keeps the lean chains of the gather form.
namespace w3 {

constexpr int kTW = 64, kTH = 32, kThreads = 256, kPX = 8;     // tile; thread = 8 pixels of one column, rows ty + 4k
constexpr int kBW = 96, kBH = 48;                               // staged box: tile + 16 / 8 pixels of flow spread
// (measured: a 64 x 16 tile with a 96 x 32 box -- 4 CTAs / SM instead of 3 -- is 9 % slower: more staging and a 3x
// instead of 2.25x L2 over-fetch per pixel)
constexpr int kBoxBytes = 3 * kBW * kBH * 4;
constexpr int kSmemBytes = kBoxBytes + 128;
constexpr int kQC = kTW / 4 + 2, kQR = kTH / 4 + 1;             // quarter-resolution cells a tile can start in

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool FLOWS>
__global__ void __launch_bounds__(kThreads, 3)
warp2_staged_kernel(const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_a,
                    const float* __restrict__ xb, const float* __restrict__ xa, const float* __restrict__ flow_hat,
                    const float* __restrict__ flow_ab, const float* __restrict__ flow_ba,
                    const float* __restrict__ tab_x, const float* __restrict__ tab_y, float* __restrict__ out,
                    float* __restrict__ flows_out, int h4, int w4, WarpGeom g) {
  extern __shared__ uint8_t smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));  // [3][kBH][kBW]
  __shared__ float4 s_c[4][kQR][kQC];
  __shared__ int s_red[kThreads / 32][4];
  __shared__ int s_box[3];
  __shared__ __align__(8) uint64_t s_bar;

  const int tid = threadIdx.x;
  const int tx = tid & (kTW - 1), ty = tid >> 6;
  const int bx = blockIdx.x * kTW, by = blockIdx.y * kTH;
  const int n = blockIdx.z;
  const int HW = g.H * g.W;
  const int hh = g.H >> 2, ww = g.W >> 2;
  const bool xin = bx + tx < g.W;
  const int x = xin ? bx + tx : g.W - 1;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int qx0 = max((bx >> 2) - 1, 0), qy0 = max((by >> 2) - 1, 0);
  {
    // 32-bit offsets from per-sample bases (a quarter-resolution plane set is far below 2^31 elements); the entry
    // index is split with constant divisions once per entry, not per address
    const int q = h4 * w4;
    const float* hat0 = flow_hat + (int64_t)n * 4 * q;
    const float* ab0 = flow_ab + (int64_t)n * 2 * q;
    const float* ba0 = flow_ba + (int64_t)n * 2 * q;
    for (int e = tid; e < 4 * kQR * kQC; e += kThreads) {
      const int ch = e / (kQR * kQC), rem = e - ch * (kQR * kQC), r = rem / kQC, c = rem - r * kQC;
      const int y0 = min(qy0 + r, hh - 1) * w4, y1 = min(qy0 + r + 1, hh - 1) * w4;
      const int x0 = min(qx0 + c, ww - 1), x1 = min(qx0 + c + 1, ww - 1);
      const float* pri = (ch < 2 ? ab0 : ba0) + (ch & 1) * q;
      const float* hat = hat0 + ch * q;
      const int o00 = y0 + x0, o01 = y0 + x1, o10 = y1 + x0, o11 = y1 + x1;
      float4 v;
      v.x = __fadd_rn(__ldg(hat + o00), __ldg(pri + o00));
      v.y = __fadd_rn(__ldg(hat + o01), __ldg(pri + o01));
      v.z = __fadd_rn(__ldg(hat + o10), __ldg(pri + o10));
      v.w = __fadd_rn(__ldg(hat + o11), __ldg(pri + o11));
      s_c[ch][r][c] = v;
    }
  }
  __syncthreads();

  const w2::Cell cx = w2::cell_of(x);
  const int rc = cx.i0 - qx0;
  const float txv = __ldg(tab_x + x);
  const float wmax = (float)(g.W - 1), hmax = (float)(g.H - 1);
  const float fW = (float)g.W, fH = (float)g.H;
  // row state of the thread's 8 pixels: rows by + ty + 4k share their upsample weights from k = 1 on (same y mod 4,
  // no top-border clamp), and their source cell advances by one per k
  const int yk0 = min(by + ty, g.H - 1), yk1 = min(by + ty + 4, g.H - 1);
  const w2::Cell cy0 = w2::cell_of(yk0), cy1 = w2::cell_of(yk1);
  const int o0 = (by + ty) * g.W + x;

  uint32_t phase = 0;  // parity of the next box copy (a direction that falls back to gathers does not use the barrier)
#pragma unroll 1
  for (int dir = 0; dir < 2; ++dir) {
    float ix[kPX], iy[kPX];
    float fmnx = 3.0e38f, fmxx = -3.0e38f, fmny = 3.0e38f, fmxy = -3.0e38f;
    // The thread's 8 pixels sit in one column, 4 rows apart: consecutive upsample cells.  The x-interpolated lower
    // edge of cell k IS the upper edge of cell k+1 (same inputs, same fma) -- computed once and carried along.
    float topu = 0.f, topv = 0.f;
    int prev_cell = -2;
#pragma unroll
    for (int k = 0; k < kPX; ++k) {
      const int yr = by + ty + 4 * k;
      const int y = min(yr, g.H - 1);
      const float l0y = k == 0 ? cy0.l0 : cy1.l0, l1y = k == 0 ? cy0.l1 : cy1.l1;
      // rows past the frame (partial bottom tile) repeat the last row's cell; their results are never stored
      const int ry = min(min((k == 0 ? cy0.i0 : cy1.i0 + (k - 1)), hh - 1) - qy0, kQR - 1);
      const float4 cu = s_c[2 * dir][ry][rc], cv = s_c[2 * dir + 1][ry][rc];
      if (ry != prev_cell + 1) {   // first pixel, top-border clamp or bottom clamp: no edge to reuse
        topu = __fmaf_rn(cx.l0, cu.x, __fmul_rn(cx.l1, cu.y));
        topv = __fmaf_rn(cx.l0, cv.x, __fmul_rn(cx.l1, cv.y));
      }
      const float botu = __fmaf_rn(cx.l0, cu.z, __fmul_rn(cx.l1, cu.w));
      const float botv = __fmaf_rn(cx.l0, cv.z, __fmul_rn(cx.l1, cv.w));
      const float u = __fmaf_rn(l0y, topu, __fmul_rn(l1y, botu));
      const float v = __fmaf_rn(l0y, topv, __fmul_rn(l1y, botv));
      topu = botu; topv = botv; prev_cell = ry;
      if (FLOWS && xin && yr < g.H) {
        float* fo = flows_out + ((int64_t)n * 4 + dir * 2) * HW + o0 + 4 * k * g.W;
        __stcg(fo, u);
        __stcg(fo + HW, v);
      }
      const float gx = __fadd_rn(txv, __fmul_rn(u, g.inv_x));
      const float gy = __fadd_rn(__ldg(tab_y + y), __fmul_rn(v, g.inv_y));
      const float px = __fmul_rn(__fmaf_rn(__fadd_rn(gx, 1.f), fW, -1.f), 0.5f);
      const float py = __fmul_rn(__fmaf_rn(__fadd_rn(gy, 1.f), fH, -1.f), 0.5f);
      ix[k] = fminf(wmax, fmaxf(px, 0.f));
      iy[k] = fminf(hmax, fmaxf(py, 0.f));
      fmnx = fminf(fmnx, ix[k]); fmxx = fmaxf(fmxx, ix[k]);
      fmny = fminf(fmny, iy[k]); fmxy = fmaxf(fmxy, iy[k]);
    }
    // coordinates are >= 0 after the clip: truncation == floor, and floor is monotonic, so the box of the floors is
    // the floor of the box
    int mnx = (int)fmnx, mxx = (int)fmxx, mny = (int)fmny, mxy = (int)fmxy;
    // ---- block-wide bounding box of the footprint (taps x0..x0+1, y0..y0+1)
    mnx = __reduce_min_sync(0xffffffffu, mnx); mxx = __reduce_max_sync(0xffffffffu, mxx);
    mny = __reduce_min_sync(0xffffffffu, mny); mxy = __reduce_max_sync(0xffffffffu, mxy);
    if ((tid & 31) == 0) {
      s_red[tid >> 5][0] = mnx; s_red[tid >> 5][1] = mxx; s_red[tid >> 5][2] = mny; s_red[tid >> 5][3] = mxy;
    }
    __syncthreads();  // also: every thread has finished reading the box of the previous direction
    if (tid == 0) {
      int a = s_red[0][0], b = s_red[0][1], c = s_red[0][2], d = s_red[0][3];
#pragma unroll
      for (int w = 1; w < kThreads / 32; ++w) {
        a = min(a, s_red[w][0]); b = max(b, s_red[w][1]); c = min(c, s_red[w][2]); d = max(d, s_red[w][3]);
      }
      a &= ~3;  // TMA: the innermost start coordinate must sit on a 16-byte boundary (unaligned starts fault)
      const bool fits = (b + 2 - a <= kBW) && (d + 2 - c <= kBH);
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
    float* op = out + ((int64_t)n * 6 + dir * 3) * HW + o0;
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
      phase ^= 1u;
      const float* tb = tile - by0 * kBW - bx0;
#pragma unroll
      for (int k = 0; k < kPX; ++k) {
        const int x0 = (int)ix[k], y0 = (int)iy[k];
        const float fx = (float)x0, fy = (float)y0;
        const float dx1 = __fsub_rn(__fadd_rn(fx, 1.f), ix[k]), dx0 = __fsub_rn(ix[k], fx);
        const float dy1 = __fsub_rn(__fadd_rn(fy, 1.f), iy[k]), dy0 = __fsub_rn(iy[k], fy);
        const float w00 = __fmul_rn(dx1, dy1), w01 = __fmul_rn(dx0, dy1);
        const float w10 = __fmul_rn(dx1, dy0), w11 = __fmul_rn(dx0, dy0);
        // taps outside the frame read TMA's zero fill: v * w == 0 exactly, the same as ATen skipping the tap
        const float* t = tb + y0 * kBW + x0;
        float r[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float acc = __fmaf_rn(t[c * (kBH * kBW)], w00, 0.f);
          acc = __fmaf_rn(t[c * (kBH * kBW) + 1], w01, acc);
          acc = __fmaf_rn(t[c * (kBH * kBW) + kBW], w10, acc);
          r[c] = __fmaf_rn(t[c * (kBH * kBW) + kBW + 1], w11, acc);
        }
        if (xin && by + ty + 4 * k < g.H) {
#pragma unroll
          for (int c = 0; c < 3; ++c) __stcg(op + c * HW + 4 * k * g.W, r[c]);
        }
      }
    } else {
      // footprint wider than the box (divergent flow): direct gathers, same arithmetic
      const float* ip = (dir == 0 ? xb : xa) + (int64_t)n * 3 * HW;
#pragma unroll
      for (int k = 0; k < kPX; ++k) {
        const Taps t = make_taps<true>(ix[k], iy[k], g.H, g.W);
        if (xin && by + ty + 4 * k < g.H) {
#pragma unroll
          for (int c = 0; c < 3; ++c) __stcg(op + c * HW + 4 * k * g.W, sample<true>(ip + (int64_t)c * HW, t));
        }
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(p);
    return nullptr;
  }();
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

}  // namespace w3

int launch_warp2_staged(const float* xb, const float* xa, const float* flow_hat, const float* flow_ab,
                        const float* flow_ba, const float* tab_x, const float* tab_y, float* out, float* flows_out,
                        int N, int H, int W, int h4, int w4, const WarpGeom& g, cudaStream_t st) {
  using namespace w3;
  static const int enabled = []() {
    const char* e = getenv("B200VC_WARP2_STAGED");
    return e ? atoi(e) : 1;
  }();
  if (!enabled || W % 4 != 0 || H < 2 || W < 2 || (int64_t)H * W < 128 * 128 || (int64_t)N * 3 >= (1 << 30) ||
      ((reinterpret_cast<uintptr_t>(xb) | reinterpret_cast<uintptr_t>(xa)) & 15u) != 0)
    return B200VC_EUNSUPPORTED;
  CUtensorMap map_b, map_a;
  if (!make_img_map(&map_b, xb, N, H, W) || !make_img_map(&map_a, xa, N, H, W)) return B200VC_EUNSUPPORTED;
  static std::once_flag once[64];
  static bool ok[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return B200VC_EUNSUPPORTED;
  std::call_once(once[dev], [&]() {
    ok[dev] = cudaFuncSetAttribute(warp2_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kSmemBytes) == cudaSuccess &&
              cudaFuncSetAttribute(warp2_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kSmemBytes) == cudaSuccess;
    if (!ok[dev]) (void)cudaGetLastError();
  });
  if (!ok[dev]) return B200VC_EUNSUPPORTED;
  dim3 grid((W + kTW - 1) / kTW, (H + kTH - 1) / kTH, N);
  if (flows_out != nullptr)
    warp2_staged_kernel<true><<<grid, kThreads, kSmemBytes, st>>>(map_b, map_a, xb, xa, flow_hat, flow_ab, flow_ba,
                                                                  tab_x, tab_y, out, flows_out, h4, w4, g);
  else
    warp2_staged_kernel<false><<<grid, kThreads, kSmemBytes, st>>>(map_b, map_a, xb, xa, flow_hat, flow_ab, flow_ba,
                                                                   tab_x, tab_y, out, flows_out, h4, w4, g);
  return check_launch("warp2_lhbdc_f32(staged)");
}

// ---------------------------------------------------------------------------------------------------------------
// K-WARP2 (Flex-Rate form): the flow glue of BidirFlowRef + both FLEX backward warps + the 16-channel concat the
// flow compressor / mask U-Net consume, in one pass (Flex-Rate.../b_model/b_model.py:34-45 `process` and :58-66):
//   LINEAR  ft0 = c.a0*f01 + c.b0*f10 ; ft1 = c.a1*f01 + c.b1*f10   (linear-motion flows to the middle frame; the
//           reference's python-float coefficients -(1-t)t, t^2, (1-t)^2, -t(1-t) arrive as fp32, each product and
//           each sum rounded once, as torch's separate elementwise kernels round)
//   REFINE  ft0 = base[0:2] + delta[0:2] ; ft1 = base[2:4] + delta[2:4]   (mv_before + flow_hat[:, :2], ...)
//   out16   = cat(ft0, ft1, x0, x1, backwarp(x0, ft0), backwarp(x1, ft1))
// The reference materialises the two flows, two warped frames and the concat separately (and the first port copied
// the flow channel slices with .contiguous() before each warp); here x0 / x1 are read once for both the copy and the
// gathers.  FLEX warp = zeros padding, half-pixel grid 2((x+u)/W - 0.5) (b_model.py:99-112): taps are predicated, and
// ATen's int-range guard on the unnormalised coordinate is kept because nothing clips it.
namespace wf {

constexpr int kTX = 128, kTY = 8, kThreads = 256, kPX = 4;

struct Coef {
  float a0, b0, a1, b1;
};

template <bool LINEAR>
__global__ void __launch_bounds__(kThreads, 2)
warp2_flex_kernel(const float* __restrict__ x0p, const float* __restrict__ x1p, const float* __restrict__ fa,
                  int64_t fa_bs, const float* __restrict__ fb, int64_t fb_bs, Coef cf, float* __restrict__ out, int N,
                  WarpGeom g) {
  const int n = blockIdx.z;
  const int y = blockIdx.y * kTY + (threadIdx.x >> 5);
  if (y >= g.H) return;
  const int lane = threadIdx.x & 31;
  const int bx = blockIdx.x * kTX;
  const int HW = g.H * g.W;
  const float* ia[2] = {x0p + (int64_t)n * 3 * HW, x1p + (int64_t)n * 3 * HW};
  const float* pa = fa + (int64_t)n * fa_bs;
  const float* pb = fb + (int64_t)n * fb_bs;
  float* po = out + (int64_t)n * 16 * HW;
#pragma unroll 1
  for (int j = 0; j < kPX; ++j) {
    const int x = bx + lane + 32 * j;
    if (x >= g.W) break;
    const int o = y * g.W + x;
    float f[4];   // ft0.x, ft0.y, ft1.x, ft1.y
    if (LINEAR) {
      const float ax = __ldg(pa + o), ay = __ldg(pa + HW + o), bxv = __ldg(pb + o), byv = __ldg(pb + HW + o);
      f[0] = __fadd_rn(__fmul_rn(cf.a0, ax), __fmul_rn(cf.b0, bxv));
      f[1] = __fadd_rn(__fmul_rn(cf.a0, ay), __fmul_rn(cf.b0, byv));
      f[2] = __fadd_rn(__fmul_rn(cf.a1, ax), __fmul_rn(cf.b1, bxv));
      f[3] = __fadd_rn(__fmul_rn(cf.a1, ay), __fmul_rn(cf.b1, byv));
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) f[c] = __fadd_rn(__ldg(pa + c * HW + o), __ldg(pb + c * HW + o));
    }
    Taps t[2];
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      float ix, iy;
      coords<B200VC_WARP_FLEX, true>(g, x, y, f[2 * dir], f[2 * dir + 1], 0.f, 0.f, ix, iy);
      t[dir] = make_taps<false>(ix, iy, g.H, g.W);
    }
    float ctr[2][3], v[2][3][4];
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* p = ia[dir] + c * HW;
        ctr[dir][c] = __ldg(p + o);                 // the frame itself (channels 4..9 of the concat)
        v[dir][c][0] = __ldg(p + t[dir].o00);       // invalid taps read element 0 and are not accumulated
        v[dir][c][1] = __ldg(p + t[dir].o01);
        v[dir][c][2] = __ldg(p + t[dir].o10);
        v[dir][c][3] = __ldg(p + t[dir].o11);
      }
    }
    // scheduling fence (see w2::warp2_kernel): every gather of the pixel is issued before the first FMA chain
    int any = 0;
#pragma unroll
    for (int dir = 0; dir < 2; ++dir)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        any |= __float_as_int(v[dir][c][0]) | __float_as_int(v[dir][c][1]) | __float_as_int(v[dir][c][2]) |
               __float_as_int(v[dir][c][3]);
    const float zero = __int_as_float(any & g.zero);
#pragma unroll
    for (int c = 0; c < 4; ++c) __stcg(po + c * HW + o, f[c]);
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        __stcg(po + (4 + 3 * dir + c) * HW + o, ctr[dir][c]);
        // ATen: out_acc = 0; out_acc += v * w for every in-bounds tap, in nw, ne, sw, se order
        float acc = zero;
        if (t[dir].v00) acc = __fmaf_rn(v[dir][c][0], t[dir].w00, acc);
        if (t[dir].v01) acc = __fmaf_rn(v[dir][c][1], t[dir].w01, acc);
        if (t[dir].v10) acc = __fmaf_rn(v[dir][c][2], t[dir].w10, acc);
        if (t[dir].v11) acc = __fmaf_rn(v[dir][c][3], t[dir].w11, acc);
        __stcg(po + (10 + 3 * dir + c) * HW + o, acc);
      }
    }
  }
}

}  // namespace wf
}  // namespace b200vc

using namespace b200vc;

extern "C" int b200vc_warp2_flex_f32(const float* x0, const float* x1, const float* fa, int64_t fa_bs, const float* fb,
                                     int64_t fb_bs, int mode, float a0, float b0, float a1, float b1, float* out16, int N,
                                     int H, int W, void* stream) {
  B200VC_REQUIRE(x0 && x1 && fa && fb && out16, "warp2_flex_f32: null pointer");
  B200VC_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0, "warp2_flex_f32: bad shape");
  B200VC_REQUIRE(mode == 0 || mode == 1, "warp2_flex_f32: mode must be 0 (linear motion) or 1 (refine), got %d", mode);
  B200VC_REQUIRE((int64_t)H * W * 16 < (1ll << 31), "warp2_flex_f32: plane too large");
  const WarpGeom g = make_geom(H, W, B200VC_WARP_FLEX, 0);
  dim3 grid((W + wf::kTX - 1) / wf::kTX, (H + wf::kTY - 1) / wf::kTY, N);
  const wf::Coef cf{a0, b0, a1, b1};
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0)
    wf::warp2_flex_kernel<true><<<grid, wf::kThreads, 0, st>>>(x0, x1, fa, fa_bs, fb, fb_bs, cf, out16, N, g);
  else
    wf::warp2_flex_kernel<false><<<grid, wf::kThreads, 0, st>>>(x0, x1, fa, fa_bs, fb, fb_bs, cf, out16, N, g);
  return check_launch("warp2_flex_f32");
}
