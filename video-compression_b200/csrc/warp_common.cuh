// Shared device code of the warp kernels (warp.cu, warp_tma.cu): geometry, ATen-exact coordinate arithmetic,
// tap construction and sampling.  See warp.cu for the reference sites.
#pragma once
#include "common.cuh"

namespace b200vc {

struct WarpGeom {
  int H, W;
  float inv_x, inv_y;  // float(1.0 / ((W-1)/2))  (ATen CUDA: tensor / python scalar == tensor * float(1/scalar))
  float den_x, den_y;  // float((W-1)/2)         (ARITH_TRUE_DIV form)
  int variant, arith;
  int zero;            // always 0, but only known at run time: the operand of the gather fence (warp2.cu)
};

// Reference arithmetic of the flow normalisation (see coords()): the divisor is the python float (W-1)/2 cast to fp32.
inline WarpGeom make_geom(int H, int W, int variant, int arith) {
  WarpGeom g;
  g.H = H;
  g.W = W;
  g.variant = variant;
  g.arith = arith;
  g.zero = 0;
  const double dx = variant == B200VC_WARP_FLEX ? (double)W : ((double)W - 1.0) / 2.0;
  const double dy = variant == B200VC_WARP_FLEX ? (double)H : ((double)H - 1.0) / 2.0;
  g.den_x = (float)dx;
  g.den_y = (float)dy;
  // ATen CUDA div by a python scalar multiplies by float(1.0 / double(scalar)): reciprocal in double, then cast
  g.inv_x = (float)(1.0 / dx);
  g.inv_y = (float)(1.0 / dy);
  return g;
}

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

constexpr int kWarpThreads = 256;  // 32 (x) x 8 (y) output pixels per CTA, one pixel per thread

// Border variants never need tap predicates: the clipped coordinate lies in [0, size-1], so only the "+1" tap
// can leave the plane, and exactly then its weight is 0.  ATen skips that tap; we clamp its address and add
// v * 0 (== skipping, for finite v).  The zeros variant (Flex) keeps ATen's per-tap bounds tests.
struct Taps {
  unsigned o00, o01, o10, o11;  // plane offsets of nw, ne, sw, se (32-bit: lets ptxas use [R.U32 + UR.64] addressing)
  float w00, w01, w10, w11;
  bool v00, v01, v10, v11;
};

template <bool BORDER>
__device__ __forceinline__ Taps make_taps(float ix, float iy, int H, int W) {
  Taps t;
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  // ATen: nw = (ix_se - ix)*(iy_se - iy), ne = (ix - ix_sw)*(iy_sw - iy), sw = (ix_ne - ix)*(iy - iy_ne), se = ...
  const float dx1 = __fsub_rn((float)x1, ix), dx0 = __fsub_rn(ix, (float)x0);
  const float dy1 = __fsub_rn((float)y1, iy), dy0 = __fsub_rn(iy, (float)y0);
  t.w00 = __fmul_rn(dx1, dy1);
  t.w01 = __fmul_rn(dx0, dy1);
  t.w10 = __fmul_rn(dx1, dy0);
  t.w11 = __fmul_rn(dx0, dy0);
  if (BORDER) {
    const int x1c = min(x1, W - 1), y1c = min(y1, H - 1);
    t.o00 = y0 * W + x0;
    t.o01 = y0 * W + x1c;
    t.o10 = y1c * W + x0;
    t.o11 = y1c * W + x1c;
    t.v00 = t.v01 = t.v10 = t.v11 = true;
  } else {
    const bool vx0 = (x0 >= 0) & (x0 < W), vx1 = (x1 >= 0) & (x1 < W);
    const bool vy0 = (y0 >= 0) & (y0 < H), vy1 = (y1 >= 0) & (y1 < H);
    t.v00 = vx0 & vy0; t.v01 = vx1 & vy0; t.v10 = vx0 & vy1; t.v11 = vx1 & vy1;
    t.o00 = t.v00 ? y0 * W + x0 : 0;
    t.o01 = t.v01 ? y0 * W + x1 : 0;
    t.o10 = t.v10 ? y1 * W + x0 : 0;
    t.o11 = t.v11 ? y1 * W + x1 : 0;
  }
  return t;
}

// out_acc = 0; out_acc += v*w per in-bounds tap in nw, ne, sw, se order (nvcc contracts ATen's += into FMA).
template <bool BORDER>
__device__ __forceinline__ float sample(const float* __restrict__ plane, const Taps& t) {
  if (BORDER) {
    const float a = __ldg(plane + t.o00), b = __ldg(plane + t.o01), c = __ldg(plane + t.o10),
                d = __ldg(plane + t.o11);
    float acc = __fmaf_rn(a, t.w00, 0.f);
    acc = __fmaf_rn(b, t.w01, acc);
    acc = __fmaf_rn(c, t.w10, acc);
    return __fmaf_rn(d, t.w11, acc);
  }
  float acc = 0.f;
  if (t.v00) acc = __fmaf_rn(__ldg(plane + t.o00), t.w00, acc);
  if (t.v01) acc = __fmaf_rn(__ldg(plane + t.o01), t.w01, acc);
  if (t.v10) acc = __fmaf_rn(__ldg(plane + t.o10), t.w10, acc);
  if (t.v11) acc = __fmaf_rn(__ldg(plane + t.o11), t.w11, acc);
  return acc;
}

// VARIANT and (for the production path) arith == 0 are compile-time so each instantiation carries one
// straight-line coordinate chain.
template <int VARIANT, bool ARITH0>
__device__ __forceinline__ void coords(const WarpGeom& g, int x, int y, float u, float v, float tx, float ty,
                                       float& ix, float& iy) {
  float gx, gy;
  const int arith = ARITH0 ? 0 : g.arith;
  if (VARIANT == B200VC_WARP_FLEX) {
    // x = gridX.float() + u ; normx = 2*(x/W - 0.5)          (b_model.py:106-109)
    const float xs = __fadd_rn((float)x, u), ys = __fadd_rn((float)y, v);
    const float qx = (arith & B200VC_ARITH_TRUE_DIV) ? __fdiv_rn(xs, g.den_x) : __fmul_rn(xs, g.inv_x);
    const float qy = (arith & B200VC_ARITH_TRUE_DIV) ? __fdiv_rn(ys, g.den_y) : __fmul_rn(ys, g.inv_y);
    gx = __fmul_rn(2.f, __fsub_rn(qx, 0.5f));
    gy = __fmul_rn(2.f, __fsub_rn(qy, 0.5f));
    ix = unnormalize(gx, g.W, false, false, arith);
    iy = unnormalize(gy, g.H, false, false, arith);
  } else {
    // grid + flow / ((W-1)/2)                                  (m.py:121-125)
    const float nu = (arith & B200VC_ARITH_TRUE_DIV) ? __fdiv_rn(u, g.den_x) : __fmul_rn(u, g.inv_x);
    const float nv = (arith & B200VC_ARITH_TRUE_DIV) ? __fdiv_rn(v, g.den_y) : __fmul_rn(v, g.inv_y);
    gx = __fadd_rn(tx, nu);
    gy = __fadd_rn(ty, nv);
    constexpr bool ac = (VARIANT == B200VC_WARP_AC1);
    ix = unnormalize(gx, g.W, ac, true, arith);
    iy = unnormalize(gy, g.H, ac, true, arith);
  }
}


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


}  // namespace b200vc
