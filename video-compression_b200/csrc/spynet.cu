// K-SPYLEVEL / K-SPYPYR: the glue of one SPyNet pyramid level and the image pyramid.
//
// LHBDC/model/flow.py:
//   :39-44  Preprocess      (x[c] - mean[c]) / std[c] with the channel order reversed, 3 sub + 3 div + cat
//   :83-88  pyramid         avg_pool2d(2, 2, count_include_pad=False) while a side is > 32 (at most 5 times)
//   :92-99  per level       up = interpolate(flow, x2, bilinear, align_corners=True) * 2.0 ; replicate-pad to the
//                           level's size ; cat([first, backwarp(second, up), up], 1)   -> the 8-channel conv input
// The reference runs ~12 kernels per image for the pyramid and 6-7 per level (x6 levels x 4 FlowNet calls per
// B-frame); the coarse levels are pure launch overhead.  Here: one kernel per image pyramid, one per level.
//
// Arithmetic is ATen's CUDA arithmetic, operation by operation (all results bit-identical to the torch chain):
//   * tensor - python_scalar = x - float(s);  tensor / python_scalar = x * float(1.0 / double(s))
//   * avg_pool2d: acc = 0; acc += v in (h, w) window order; acc / 4
//   * upsample_bilinear2d(align_corners=True): scale = float(in-1)/(out-1); src = scale*dst; i0 = (int)src;
//     i1 = i0 + (i0 < in-1); l1 = src - i0; l0 = 1 - l1; value = fma(l0y, fma(l0x,a,l1x*b), l1y*fma(l0x,c,l1x*d))
//     (same kernel template as the x4 align_corners=False upsample verified in warp2)
//   * backwarp = WARP_LHBDC coordinate chain of warp_common.cuh.
#include "common.cuh"
#include "warp_common.cuh"

namespace b200vc {

int launch_spynet_level_tma(const float* first, int64_t first_bs, const float* second, int64_t second_bs,
                            const float* flow_prev, const float* tab_x, const float* tab_y, float* feat, int N, int H,
                            int W, int hp, int wp, float sy, float sx, const WarpGeom& g, cudaStream_t st);  // warp_tma.cu

// ------------------------------------------------------------------------------------------------ pyramid
constexpr int kPyrMaxLevels = 5;
constexpr int kPyrTileW = 64, kPyrTileH = 32;  // level-0 pixels per CTA: 2^5-aligned, so every level stays CTA-local

struct PyrArgs {
  float* lvl[kPyrMaxLevels + 1];  // lvl[0] = full resolution (preprocessed), lvl[l] = l times pooled; [N,3,h,w] each
  int h[kPyrMaxLevels + 1], w[kPyrMaxLevels + 1];
  int levels;  // number of poolings (0..5)
};

__device__ __forceinline__ float pool4(float a, float b, float c, float d) {
  // ATen avg_pool2d_out_cuda_frame: aveval = 0; aveval += ...; aveval / divide_factor
  return __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(0.f, a), b), c), d), 4.f);
}

template <bool PRE, bool VEC>
__global__ void __launch_bounds__(256)
spynet_pyramid_kernel(const float* __restrict__ in, int64_t in_bs, PyrArgs a) {
  __shared__ float s[2][kPyrTileH / 2][kPyrTileW / 2];  // ping-pong of the pooled tile
  const int n = blockIdx.z / 3, c_out = blockIdx.z % 3;
  // Preprocess reverses the channel order (flow.py:40-44) and normalises with the ImageNet statistics
  const int c_in = PRE ? 2 - c_out : c_out;
  const float mean = c_out == 0 ? 0.485f : (c_out == 1 ? 0.456f : 0.406f);
  // ATen: tensor / python_scalar = tensor * float(1.0 / double(scalar))  (reciprocal in double, then cast; for
  // 0.224 this differs from 1.0f / 0.224f in the last bit -- measured, tools/spynet_probe.py)
  const float inv = c_out == 0 ? (float)(1.0 / 0.229) : (c_out == 1 ? (float)(1.0 / 0.224) : (float)(1.0 / 0.225));
  const int H = a.h[0], W = a.w[0];
  const float* src = in + (int64_t)n * in_bs + (int64_t)c_in * H * W;
  float* dst0 = a.lvl[0] + ((int64_t)n * 3 + c_out) * H * W;
  const int x0 = blockIdx.x * kPyrTileW, y0 = blockIdx.y * kPyrTileH;

  if (VEC) {
    // level 0 (copy / preprocess) and level 1: every thread owns 4 x 2 pixels = two 2x2 blocks (W % 4 == 0: 16-byte
    // loads / stores, a row of the tile is 16 threads x 16 B)
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads
    const int x = x0 + 4 * tx, y = y0 + 2 * ty;
    float4 r[2];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const bool ok = (x < W) && (y + dy < H);
      float4 t = ok ? ld_stream4(src + (int64_t)(y + dy) * W + x) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (PRE) {
        t.x = __fmul_rn(__fsub_rn(t.x, mean), inv); t.y = __fmul_rn(__fsub_rn(t.y, mean), inv);
        t.z = __fmul_rn(__fsub_rn(t.z, mean), inv); t.w = __fmul_rn(__fsub_rn(t.w, mean), inv);
        if (ok) *reinterpret_cast<float4*>(dst0 + (int64_t)(y + dy) * W + x) = t;
      }
      r[dy] = t;
    }
    const float p0 = pool4(r[0].x, r[0].y, r[1].x, r[1].y), p1 = pool4(r[0].z, r[0].w, r[1].z, r[1].w);
    s[0][ty][2 * tx] = p0;
    s[0][ty][2 * tx + 1] = p1;
    if (a.levels >= 1) {
      const int px = (x0 >> 1) + 2 * tx, py = (y0 >> 1) + ty;
      float* d1 = a.lvl[1] + (((int64_t)n * 3 + c_out) * a.h[1] + py) * a.w[1] + px;
      if (py < a.h[1]) {
        if (px + 1 < a.w[1]) *reinterpret_cast<float2*>(d1) = make_float2(p0, p1);  // w[1] = W/2 is even: 8-byte aligned
        else if (px < a.w[1]) d1[0] = p0;
      }
    }
  } else {
    // level 0 (copy / preprocess) and level 1: every thread owns two 2x2 blocks (32 x 16 blocks per tile)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int bx = threadIdx.x & 31, by = (threadIdx.x >> 5) + 8 * k;
      const int x = x0 + 2 * bx, y = y0 + 2 * by;
      float v[2][2];
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const bool ok = (x + dx < W) && (y + dy < H);
          float t = ok ? __ldg(src + (int64_t)(y + dy) * W + x + dx) : 0.f;
          if (PRE) t = __fmul_rn(__fsub_rn(t, mean), inv);
          v[dy][dx] = t;
          if (PRE && ok) dst0[(int64_t)(y + dy) * W + x + dx] = t;
        }
      const float p = pool4(v[0][0], v[0][1], v[1][0], v[1][1]);
      s[0][by][bx] = p;
      if (a.levels >= 1) {
        const int px = (x0 >> 1) + bx, py = (y0 >> 1) + by;
        if (px < a.w[1] && py < a.h[1]) a.lvl[1][(((int64_t)n * 3 + c_out) * a.h[1] + py) * a.w[1] + px] = p;
      }
    }
  }
  // levels 2..5 from shared memory: level l reads buffer (l & 1) and writes the other one (level 1 sits in s[0])
  // (fully unrolled: a runtime index into the by-value argument struct would force a local-memory copy of it)
#pragma unroll
  for (int l = 2; l <= kPyrMaxLevels; ++l) {
    if (l > a.levels) break;
    __syncthreads();
    const int tw = kPyrTileW >> l, th = kPyrTileH >> l;
    const int from = l & 1, to = from ^ 1;
    if (threadIdx.x < tw * th) {
      const int bx = threadIdx.x % tw, by = threadIdx.x / tw;
      const float p = pool4(s[from][2 * by][2 * bx], s[from][2 * by][2 * bx + 1], s[from][2 * by + 1][2 * bx],
                            s[from][2 * by + 1][2 * bx + 1]);
      s[to][by][bx] = p;
      const int px = (x0 >> l) + bx, py = (y0 >> l) + by;
      if (px < a.w[l] && py < a.h[l]) a.lvl[l][(((int64_t)n * 3 + c_out) * a.h[l] + py) * a.w[l] + px] = p;
    }
  }
}

// ------------------------------------------------------------------------------------------------ level
struct UpAC {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ UpAC up_ac_index(int dst, int in_size, float scale) {
  UpAC r;
  const float src = __fmul_rn(scale, (float)dst);
  r.i0 = (int)src;
  r.i1 = r.i0 + ((r.i0 < in_size - 1) ? 1 : 0);
  r.l1 = __fsub_rn(src, (float)r.i0);
  r.l0 = __fsub_rn(1.f, r.l1);
  return r;
}

// One thread per PX output pixels (rows y, y+8, ...) of a 32 x 8*PX tile (the gather-kernel geometry of warp.cu;
// the kernel is latency-bound -- flow taps, then image taps -- so PX independent chains are kept in flight).
// feat[n] = [ first (3) | backwarp(second, up) (3) | up (2) ],  up = 2 * upsample_x2(flow_prev), replicate-padded.
template <int PX>
__global__ void __launch_bounds__(kWarpThreads)
spynet_level_kernel(const float* __restrict__ first, int64_t first_bs, const float* __restrict__ second,
                    int64_t second_bs, const float* __restrict__ flow_prev, const float* __restrict__ tab_x,
                    const float* __restrict__ tab_y, float* __restrict__ feat, int hp, int wp, float scale_y,
                    float scale_x, WarpGeom g) {
  constexpr int kRows = kWarpThreads / 32;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y0 = blockIdx.y * (kRows * PX) + (threadIdx.x >> 5);
  const int n = blockIdx.z;
  if (x >= g.W) return;
  const int HW = g.H * g.W;
  float u[PX], v[PX];
#pragma unroll
  for (int k = 0; k < PX; ++k) u[k] = v[k] = 0.f;
  if (flow_prev != nullptr) {
    // replicate pad (flow.py:95-96): the last row / column of the upsampled flow is repeated once
    const int xs = min(x, 2 * wp - 1);
    const UpAC ux = up_ac_index(xs, wp, scale_x);
    const float* fu = flow_prev + (int64_t)n * 2 * hp * wp;
    const float* fv = fu + hp * wp;
    float ta[PX][4], tb[PX][4];
    UpAC uy[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) {
      const int ys = min(min(y0 + k * kRows, g.H - 1), 2 * hp - 1);
      uy[k] = up_ac_index(ys, hp, scale_y);
      const int o00 = uy[k].i0 * wp + ux.i0, o01 = uy[k].i0 * wp + ux.i1, o10 = uy[k].i1 * wp + ux.i0,
                o11 = uy[k].i1 * wp + ux.i1;
      ta[k][0] = __ldg(fu + o00); ta[k][1] = __ldg(fu + o01); ta[k][2] = __ldg(fu + o10); ta[k][3] = __ldg(fu + o11);
      tb[k][0] = __ldg(fv + o00); tb[k][1] = __ldg(fv + o01); tb[k][2] = __ldg(fv + o10); tb[k][3] = __ldg(fv + o11);
    }
#pragma unroll
    for (int k = 0; k < PX; ++k) {
      const float ui = __fmaf_rn(uy[k].l0, __fmaf_rn(ux.l0, ta[k][0], __fmul_rn(ux.l1, ta[k][1])),
                                 __fmul_rn(uy[k].l1, __fmaf_rn(ux.l0, ta[k][2], __fmul_rn(ux.l1, ta[k][3]))));
      const float vi = __fmaf_rn(uy[k].l0, __fmaf_rn(ux.l0, tb[k][0], __fmul_rn(ux.l1, tb[k][1])),
                                 __fmul_rn(uy[k].l1, __fmaf_rn(ux.l0, tb[k][2], __fmul_rn(ux.l1, tb[k][3]))));
      u[k] = __fmul_rn(ui, 2.0f);
      v[k] = __fmul_rn(vi, 2.0f);
    }
  }
  const float tx = __ldg(tab_x + x);
  Taps t[PX];
#pragma unroll
  for (int k = 0; k < PX; ++k) {
    const int y = min(y0 + k * kRows, g.H - 1);
    float ix, iy;
    coords<B200VC_WARP_LHBDC, true>(g, x, y, u[k], v[k], tx, __ldg(tab_y + y), ix, iy);
    t[k] = make_taps<true>(ix, iy, g.H, g.W);
  }
  const float* sp = second + (int64_t)n * second_bs;
  const float* fp = first + (int64_t)n * first_bs;
  float r[PX][3], f[PX][3];
#pragma unroll
  for (int k = 0; k < PX; ++k) {
    const int y = min(y0 + k * kRows, g.H - 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      r[k][c] = sample<true>(sp + (int64_t)c * HW, t[k]);
      f[k][c] = __ldg(fp + (int64_t)c * HW + y * g.W + x);
    }
  }
#pragma unroll
  for (int k = 0; k < PX; ++k) {
    const int y = y0 + k * kRows;
    if (y < g.H) {
      float* op = feat + (int64_t)n * 8 * HW + y * g.W + x;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        op[(int64_t)c * HW] = f[k][c];
        op[(int64_t)(3 + c) * HW] = r[k][c];
      }
      op[(int64_t)6 * HW] = u[k];
      op[(int64_t)7 * HW] = v[k];
    }
  }
}

}  // namespace b200vc

using namespace b200vc;

extern "C" int b200vc_spynet_pyramid_f32(const float* frame, int64_t frame_bs, float* const* levels, int N, int H,
                                         int W, int n_levels, int preprocess, void* stream) {
  B200VC_REQUIRE(frame && levels, "spynet_pyramid_f32: null pointer");
  B200VC_REQUIRE(N > 0 && H > 0 && W > 0, "spynet_pyramid_f32: bad shape N=%d H=%d W=%d", N, H, W);
  B200VC_REQUIRE(n_levels >= 0 && n_levels <= kPyrMaxLevels, "spynet_pyramid_f32: %d poolings (max %d)", n_levels,
                 kPyrMaxLevels);
  B200VC_REQUIRE((int64_t)N * 3 <= 65535, "spynet_pyramid_f32: N too large");
  B200VC_REQUIRE((int64_t)H * W < (1ll << 31), "spynet_pyramid_f32: plane too large");
  B200VC_REQUIRE(preprocess || n_levels >= 1, "spynet_pyramid_f32: nothing to do");
  PyrArgs a;
  a.levels = n_levels;
  int h = H, w = W;
  for (int l = 0; l <= kPyrMaxLevels; ++l) {
    a.lvl[l] = nullptr;
    a.h[l] = a.w[l] = 0;
  }
  for (int l = 0; l <= n_levels; ++l) {
    a.h[l] = h;
    a.w[l] = w;
    a.lvl[l] = levels[l];
    B200VC_REQUIRE(l == 0 ? (levels[0] != nullptr || !preprocess) : levels[l] != nullptr,
                   "spynet_pyramid_f32: level %d output missing", l);
    h /= 2;
    w /= 2;
    B200VC_REQUIRE(l == n_levels || (h > 0 && w > 0), "spynet_pyramid_f32: level %d would be empty", l + 1);
  }
  dim3 grid((W + kPyrTileW - 1) / kPyrTileW, (H + kPyrTileH - 1) / kPyrTileH, N * 3);
  cudaStream_t st = (cudaStream_t)stream;
  // 16-byte path: rows of the frame and of level 0 must start 16-byte aligned
  const bool vec = (W % 4 == 0) && (frame_bs % 4 == 0) && ((reinterpret_cast<uintptr_t>(frame) & 15) == 0) &&
                   (!preprocess || (reinterpret_cast<uintptr_t>(levels[0]) & 15) == 0) &&
                   (n_levels < 1 || (reinterpret_cast<uintptr_t>(levels[1]) & 7) == 0);
  if (preprocess) {
    if (vec) spynet_pyramid_kernel<true, true><<<grid, 256, 0, st>>>(frame, frame_bs, a);
    else spynet_pyramid_kernel<true, false><<<grid, 256, 0, st>>>(frame, frame_bs, a);
  } else {
    if (vec) spynet_pyramid_kernel<false, true><<<grid, 256, 0, st>>>(frame, frame_bs, a);
    else spynet_pyramid_kernel<false, false><<<grid, 256, 0, st>>>(frame, frame_bs, a);
  }
  return check_launch("spynet_pyramid_f32");
}

extern "C" int b200vc_spynet_level_f32(const float* first, int64_t first_bs, const float* second, int64_t second_bs,
                                       const float* flow_prev, const float* tab_x, const float* tab_y, float* feat,
                                       int N, int H, int W, int hp, int wp, void* stream) {
  B200VC_REQUIRE(first && second && feat && tab_x && tab_y, "spynet_level_f32: null pointer");
  B200VC_REQUIRE(N > 0 && H > 0 && W > 0, "spynet_level_f32: bad shape N=%d H=%d W=%d", N, H, W);
  B200VC_REQUIRE(N <= 65535, "spynet_level_f32: N too large");
  B200VC_REQUIRE((int64_t)H * W < (1ll << 31), "spynet_level_f32: plane too large");
  if (flow_prev != nullptr) {
    // interpolate(scale_factor=2) then at most one replicated row / column (flow.py:95-96)
    B200VC_REQUIRE(hp > 0 && wp > 0 && (H == 2 * hp || H == 2 * hp + 1) && (W == 2 * wp || W == 2 * wp + 1),
                   "spynet_level_f32: previous flow %dx%d does not upsample to %dx%d", hp, wp, H, W);
  }
  const WarpGeom g = make_geom(H, W, B200VC_WARP_LHBDC, 0);
  // area_pixel_compute_scale(align_corners=True): (T)(in - 1) / (out - 1), 0 when out == 1
  const float sy = (flow_prev && 2 * hp > 1) ? (float)(hp - 1) / (float)(2 * hp - 1) : 0.f;
  const float sx = (flow_prev && 2 * wp > 1) ? (float)(wp - 1) / (float)(2 * wp - 1) : 0.f;
  {
    const int rc = launch_spynet_level_tma(first, first_bs, second, second_bs, flow_prev, tab_x, tab_y, feat, N, H, W,
                                           hp, wp, sy, sx, g, (cudaStream_t)stream);
    if (rc != B200VC_EUNSUPPORTED) return rc;
  }
  const int rows = kWarpThreads / 32;
  // 2 pixels per thread once the plane is large enough to still fill the machine (same rule as warp_f32)
  const int px = (int64_t)N * H * W >= (1 << 19) ? 2 : 1;
  dim3 grid((W + 31) / 32, (H + rows * px - 1) / (rows * px), N);
  if (px == 2)
    spynet_level_kernel<2><<<grid, kWarpThreads, 0, (cudaStream_t)stream>>>(first, first_bs, second, second_bs,
                                                                         flow_prev, tab_x, tab_y, feat, hp, wp, sy, sx, g);
  else
    spynet_level_kernel<1><<<grid, kWarpThreads, 0, (cudaStream_t)stream>>>(first, first_bs, second, second_bs,
                                                                         flow_prev, tab_x, tab_y, feat, hp, wp, sy, sx, g);
  return check_launch("spynet_level_f32");
}
