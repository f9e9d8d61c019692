// K-WARP / K-WARP2: flow-driven bilinear backward warp.
//
// Replaces torch.grid_sample + the flow-normalisation glue of
//   LHBDC/model/m.py:111-126, LHBDC/model/flow.py:15-25            (WARP_LHBDC)
//   Flex-Rate.../b_model/b_model.py:99-112                         (WARP_FLEX)
//   ICIP2024/src/model/m.py:262-282, OJSP2025/video_model.py:668-676 (WARP_AC1)
// and, for warp2, LHBDC/model/m.py:55-63 (x4 flow upsample + both warps + concat).
//
// HBM-bound gather.  Layout NCHW planar fp32 (the reference layout).  One thread owns VEC=4 horizontally
// adjacent output pixels: flow is read with one 128-bit load per plane, each output plane is written with
// one 128-bit store, the four bilinear taps per pixel are scalar read-only loads that hit L1/L2 (neighbouring
// threads sample neighbouring source pixels).  Algorithmic bytes: (2C+2)*4 per output pixel (C=3: 32 B/px).
//
// The coordinate path reproduces ATen's CUDA arithmetic operation by operation:
// every step is pinned with __f*_rn intrinsics so nvcc cannot re-associate or contract differently.
#include "common.cuh"

namespace b200vc {

struct WarpGeom {
  int H, W;
  float inv_x, inv_y;  // 1.0f / float((W-1)/2)  (ATen CUDA: tensor / python scalar == tensor * (1/scalar))
  float den_x, den_y;  // float((W-1)/2)         (ARITH_TRUE_DIV form)
  int variant, arith;
};

// Normalised grid coordinate -> source pixel coordinate, exactly as ATen's grid_sampler_compute_source_index.
__device__ __forceinline__ float unnormalize(float g, int size, bool align_corners, bool border, int arith) {
  float c;
  if (align_corners) {
    c = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), (float)(size - 1));
  } else if (arith & B200VC_ARITH_NO_FMA) {
    c = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)size), 1.f), 0.5f);
  } else {
    c = __fmul_rn(__fmaf_rn(__fadd_rn(g, 1.f), (float)size, -1.f), 0.5f);
  }
  if (border) c = fminf((float)(size - 1), fmaxf(c, 0.f));
  // safe_downgrade_to_int_range
  if (!(c <= 2147483520.f && c >= -2147483648.f)) c = -100.f;
  return c;
}

struct Taps {
  int off[4];   // plane offsets of nw, ne, sw, se (or -1 when out of bounds)
  float w[4];
};

__device__ __forceinline__ Taps make_taps(float ix, float iy, int H, int W) {
  Taps t;
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  // ATen: nw = (ix_se - ix)*(iy_se - iy), ne = (ix - ix_sw)*(iy_sw - iy), sw = (ix_ne - ix)*(iy - iy_ne), se = ...
  const float dx1 = __fsub_rn((float)x1, ix), dx0 = __fsub_rn(ix, (float)x0);
  const float dy1 = __fsub_rn((float)y1, iy), dy0 = __fsub_rn(iy, (float)y0);
  t.w[0] = __fmul_rn(dx1, dy1);
  t.w[1] = __fmul_rn(dx0, dy1);
  t.w[2] = __fmul_rn(dx1, dy0);
  t.w[3] = __fmul_rn(dx0, dy0);
  const bool vx0 = (x0 >= 0) & (x0 < W), vx1 = (x1 >= 0) & (x1 < W);
  const bool vy0 = (y0 >= 0) & (y0 < H), vy1 = (y1 >= 0) & (y1 < H);
  t.off[0] = (vx0 & vy0) ? y0 * W + x0 : -1;
  t.off[1] = (vx1 & vy0) ? y0 * W + x1 : -1;
  t.off[2] = (vx0 & vy1) ? y1 * W + x0 : -1;
  t.off[3] = (vx1 & vy1) ? y1 * W + x1 : -1;
  return t;
}

// out_acc = 0; out_acc += v*w per in-bounds tap in nw, ne, sw, se order (nvcc contracts ATen's += into FMA).
__device__ __forceinline__ float sample(const float* __restrict__ plane, const Taps& t) {
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (t.off[k] >= 0) acc = __fmaf_rn(__ldg(plane + t.off[k]), t.w[k], acc);
  return acc;
}

__device__ __forceinline__ void coords(const WarpGeom& g, int x, int y, float u, float v, float tx, float ty,
                                       float& ix, float& iy) {
  float gx, gy;
  if (g.variant == B200VC_WARP_FLEX) {
    // x = gridX.float() + u ; normx = 2*(x/W - 0.5)          (b_model.py:106-109)
    const float xs = __fadd_rn((float)x, u), ys = __fadd_rn((float)y, v);
    const float qx = (g.arith & B200VC_ARITH_TRUE_DIV) ? __fdiv_rn(xs, g.den_x) : __fmul_rn(xs, g.inv_x);
    const float qy = (g.arith & B200VC_ARITH_TRUE_DIV) ? __fdiv_rn(ys, g.den_y) : __fmul_rn(ys, g.inv_y);
    gx = __fmul_rn(2.f, __fsub_rn(qx, 0.5f));
    gy = __fmul_rn(2.f, __fsub_rn(qy, 0.5f));
    ix = unnormalize(gx, g.W, false, false, g.arith);
    iy = unnormalize(gy, g.H, false, false, g.arith);
  } else {
    // grid + flow / ((W-1)/2)                                  (m.py:121-125)
    const float nu = (g.arith & B200VC_ARITH_TRUE_DIV) ? __fdiv_rn(u, g.den_x) : __fmul_rn(u, g.inv_x);
    const float nv = (g.arith & B200VC_ARITH_TRUE_DIV) ? __fdiv_rn(v, g.den_y) : __fmul_rn(v, g.inv_y);
    gx = __fadd_rn(tx, nu);
    gy = __fadd_rn(ty, nv);
    const bool ac = (g.variant == B200VC_WARP_AC1);
    ix = unnormalize(gx, g.W, ac, true, g.arith);
    iy = unnormalize(gy, g.H, ac, true, g.arith);
  }
}

constexpr int kWarpThreads = 256;

template <int VEC>
__global__ void __launch_bounds__(kWarpThreads)
warp_kernel(const float* __restrict__ img, int64_t img_bs, const float* __restrict__ flow,
            const float* __restrict__ tab_x, const float* __restrict__ tab_y, float* __restrict__ out,
            int64_t out_bs, int C, WarpGeom g) {
  const int Wv = (g.W + VEC - 1) / VEC;
  const int xv = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * (kWarpThreads / 32) + (threadIdx.x >> 5);
  const int n = blockIdx.z;
  if (xv >= Wv || y >= g.H) return;
  const int x0 = xv * VEC;
  const int64_t HW = (int64_t)g.H * g.W;
  const float* fu = flow + (int64_t)n * 2 * HW + (int64_t)y * g.W + x0;
  const float* fv = fu + HW;
  float u[VEC], v[VEC];
  if constexpr (VEC == 4) {
    const float4 a = ld_stream4(fu), b = ld_stream4(fv);
    u[0] = a.x; u[1] = a.y; u[2] = a.z; u[3] = a.w;
    v[0] = b.x; v[1] = b.y; v[2] = b.z; v[3] = b.w;
  } else {
    u[0] = __ldg(fu);
    v[0] = __ldg(fv);
  }
  const bool flex = (g.variant == B200VC_WARP_FLEX);
  const float ty = flex ? 0.f : __ldg(tab_y + y);
  Taps t[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const float tx = flex ? 0.f : __ldg(tab_x + x0 + k);
    float ix, iy;
    coords(g, x0 + k, y, u[k], v[k], tx, ty, ix, iy);
    t[k] = make_taps(ix, iy, g.H, g.W);
  }
  const float* ip = img + (int64_t)n * img_bs;
  float* op = out + (int64_t)n * out_bs + (int64_t)y * g.W + x0;
  for (int c = 0; c < C; ++c) {
    const float* plane = ip + (int64_t)c * HW;
    float r[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) r[k] = sample(plane, t[k]);
    if constexpr (VEC == 4) {
      st_stream4(op + (int64_t)c * HW, make_float4(r[0], r[1], r[2], r[3]));
    } else {
      op[(int64_t)c * HW] = r[0];
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// warp2: LHBDC flow glue + two warps + concat in one pass.
// ATen upsample_bilinear2d (align_corners=False, scale_factor=4 => rscale = 0.25):
//   src = max(0.25*(dst+0.5) - 0.5, 0); i0 = (int)src; ip = i0 < in-1; l1 = src - i0; l0 = 1 - l1
//   val = l0y*(l0x*a + l1x*b) + l1y*(l0x*c + l1x*d)
struct Up4 {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Up4 up4_index(int dst, int in_size) {
  Up4 r;
  float src = __fmaf_rn(0.25f, (float)dst + 0.5f, -0.5f);
  src = src < 0.f ? 0.f : src;
  r.i0 = (int)src;
  r.i1 = r.i0 + ((r.i0 < in_size - 1) ? 1 : 0);
  r.l1 = __fsub_rn(src, (float)r.i0);
  r.l0 = __fsub_rn(1.f, r.l1);
  return r;
}

__device__ __forceinline__ float up4_value(const Up4& uy, const Up4& ux, float a, float b, float c, float d,
                                           int arith) {
  if (arith & B200VC_ARITH_NO_FMA) {
    const float top = __fadd_rn(__fmul_rn(ux.l0, a), __fmul_rn(ux.l1, b));
    const float bot = __fadd_rn(__fmul_rn(ux.l0, c), __fmul_rn(ux.l1, d));
    return __fadd_rn(__fmul_rn(uy.l0, top), __fmul_rn(uy.l1, bot));
  }
  const float top = __fmaf_rn(ux.l0, a, __fmul_rn(ux.l1, b));
  const float bot = __fmaf_rn(ux.l0, c, __fmul_rn(ux.l1, d));
  return __fmaf_rn(uy.l0, top, __fmul_rn(uy.l1, bot));
}

__global__ void __launch_bounds__(kWarpThreads)
warp2_lhbdc_kernel(const float* __restrict__ xb, const float* __restrict__ xa,
                   const float* __restrict__ flow_hat, const float* __restrict__ flow_ab,
                   const float* __restrict__ flow_ba, const float* __restrict__ tab_x,
                   const float* __restrict__ tab_y, float* __restrict__ out, float* __restrict__ flows_out,
                   int h4, int w4, WarpGeom g) {
  constexpr int VEC = 4;
  const int Wv = g.W / VEC;
  const int xv = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * (kWarpThreads / 32) + (threadIdx.x >> 5);
  const int n = blockIdx.z;
  if (xv >= Wv || y >= g.H) return;
  const int x0 = xv * VEC;
  const int hh = g.H / 4, ww = g.W / 4;
  const int64_t HW = (int64_t)g.H * g.W;
  const int64_t q = (int64_t)h4 * w4;
  const Up4 uy = up4_index(y, hh);
  const float ty = __ldg(tab_y + y);
  float* op = out + (int64_t)n * 6 * HW + (int64_t)y * g.W + x0;
#pragma unroll
  for (int dir = 0; dir < 2; ++dir) {
    // quarter-res value = mv x_hat chunk + linear-motion prior (m.py:56,58), rounded once like torch's add
    const float* hat = flow_hat + ((int64_t)n * 4 + dir * 2) * q;
    const float* pri = (dir == 0 ? flow_ab : flow_ba) + (int64_t)n * 2 * q;
    float u[VEC], v[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const Up4 ux = up4_index(x0 + k, ww);
      const int o00 = uy.i0 * w4 + ux.i0, o01 = uy.i0 * w4 + ux.i1;
      const int o10 = uy.i1 * w4 + ux.i0, o11 = uy.i1 * w4 + ux.i1;
      float f[2];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const float* h = hat + ch * q;
        const float* p = pri + ch * q;
        const float a = __fadd_rn(__ldg(h + o00), __ldg(p + o00));
        const float b = __fadd_rn(__ldg(h + o01), __ldg(p + o01));
        const float c = __fadd_rn(__ldg(h + o10), __ldg(p + o10));
        const float d = __fadd_rn(__ldg(h + o11), __ldg(p + o11));
        f[ch] = up4_value(uy, ux, a, b, c, d, g.arith);
      }
      u[k] = f[0];
      v[k] = f[1];
    }
    if (flows_out != nullptr) {
      float* fo = flows_out + ((int64_t)n * 4 + dir * 2) * HW + (int64_t)y * g.W + x0;
      st_stream4(fo, make_float4(u[0], u[1], u[2], u[3]));
      st_stream4(fo + HW, make_float4(v[0], v[1], v[2], v[3]));
    }
    Taps t[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float ix, iy;
      coords(g, x0 + k, y, u[k], v[k], __ldg(tab_x + x0 + k), ty, ix, iy);
      t[k] = make_taps(ix, iy, g.H, g.W);
    }
    const float* ip = (dir == 0 ? xb : xa) + (int64_t)n * 3 * HW;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* plane = ip + (int64_t)c * HW;
      float r[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) r[k] = sample(plane, t[k]);
      st_stream4(op + (int64_t)(dir * 3 + c) * HW, make_float4(r[0], r[1], r[2], r[3]));
    }
  }
}

static WarpGeom make_geom(int H, int W, int variant, int arith) {
  WarpGeom g;
  g.H = H;
  g.W = W;
  g.variant = variant;
  g.arith = arith;
  if (variant == B200VC_WARP_FLEX) {
    g.den_x = (float)W;
    g.den_y = (float)H;
  } else {
    g.den_x = (float)(((double)W - 1.0) / 2.0);
    g.den_y = (float)(((double)H - 1.0) / 2.0);
  }
  g.inv_x = 1.0f / g.den_x;
  g.inv_y = 1.0f / g.den_y;
  return g;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

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
  const int64_t HW = (int64_t)H * W;
  const bool vec = (W % 4 == 0) && aligned16(flow) && aligned16(out) && (out_bs % 4 == 0) && (HW % 4 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = kWarpThreads / 32;
  if (vec) {
    dim3 grid((W / 4 + 31) / 32, (H + rows - 1) / rows, N);
    warp_kernel<4><<<grid, kWarpThreads, 0, st>>>(img, img_bs, flow, tab_x, tab_y, out, out_bs, C, g);
  } else {
    dim3 grid((W + 31) / 32, (H + rows - 1) / rows, N);
    warp_kernel<1><<<grid, kWarpThreads, 0, st>>>(img, img_bs, flow, tab_x, tab_y, out, out_bs, C, g);
  }
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
  B200VC_REQUIRE(aligned16(out) && (flows_out == nullptr || aligned16(flows_out)),
                 "warp2_lhbdc_f32: outputs must be 16-byte aligned");
  B200VC_REQUIRE((int64_t)H * W < (1ll << 31), "warp2_lhbdc_f32: plane too large");
  const WarpGeom g = make_geom(H, W, B200VC_WARP_LHBDC, arith);
  const int rows = kWarpThreads / 32;
  dim3 grid((W / 4 + 31) / 32, (H + rows - 1) / rows, N);
  warp2_lhbdc_kernel<<<grid, kWarpThreads, 0, (cudaStream_t)stream>>>(
      x_before, x_after, flow_hat, flow_ab, flow_ba, tab_x, tab_y, out, flows_out, h4, w4, g);
  return check_launch("warp2_lhbdc_f32");
}
