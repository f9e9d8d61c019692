// K-DCN: modulated deformable convolution.
//
// Replaces torchvision.ops.deform_conv2d behind the reference's DeformConv2d modules:
//   ICIP2023/src/model/m.py:29-34        DeformConv2d(96|64|32, same, kernel_size=3, padding=1, groups=8)
//   ICIP2024/src/model/helpers.py:40,57  OffsetDiversity.fusion = DeformConv2d(2C, C, 3, padding=1, groups=16), C = 64/96/128
// torchvision materialises the deformable im2col matrix ([N, Cin*9, Ho*Wo] fp32: 9x the input, written and re-read)
// and runs one small GEMM per weight group.  The reference's layers are *grouped* with 4-16 channels per group, so the
// GEMMs are tiny (K = 36..144, 4..12 output rows) and the operator is a gather, not a matrix product: here the
// samples are consumed as they are produced -- one thread owns one output position of one weight group, keeps the
// group's CO output channels in registers and walks (offset group, kernel point, input channel).  Offsets, masks and
// outputs are coalesced over positions; taps are L1 gathers around the position, like K-WARP.
// HBM traffic = input + offsets + mask + output once (no column matrix).
//
// Sampling arithmetic = torchvision/csrc/ops/cuda/deform_conv2d_kernel.cu (bilinear_interpolate + mask); the sum over
// (channel, kernel point) is sequential per thread where torchvision uses a cuBLAS GEMM, so results agree to fp32
// summation-order noise (tests hold 2e-5 of the output magnitude), not bit for bit.
#include "common.cuh"

namespace b200vc {

struct DcnParams {
  int N, Cin, H, W, Cout, Ho, Wo;
  int kh, kw, sh, sw, ph, pw, dh, dw;
  int groups, og;        // weight groups, offset groups
  int cin_g, cout_g;     // channels per weight group
  int ch_og;             // input channels per offset group
  int chunks;            // cout_g / CO
};

constexpr int kDcnThreads = 128;

template <int CO>
__global__ void __launch_bounds__(kDcnThreads)
deform_conv2d_kernel(const float* __restrict__ in, const float* __restrict__ offset, const float* __restrict__ mask,
                     const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ out,
                     DcnParams p) {
  extern __shared__ __align__(16) float s_w[];  // [cin_g][K][CO] of this (weight group, output chunk): one vector load per sample
  const int K = p.kh * p.kw;
  const int wg = blockIdx.y / p.chunks, chunk = blockIdx.y % p.chunks;
  const int n = blockIdx.z;
  const int co0 = wg * p.cout_g + chunk * CO;  // first output channel of this thread block
  const int nw = CO * p.cin_g * K;
  for (int i = threadIdx.x; i < nw; i += kDcnThreads) {
    const int o = i / (p.cin_g * K), ck = i - o * (p.cin_g * K);  // global order [o][c][k]
    s_w[ck * CO + o] = __ldg(weight + (int64_t)co0 * p.cin_g * K + i);
  }
  __syncthreads();

  const int P = p.Ho * p.Wo;
  const int pos = blockIdx.x * kDcnThreads + threadIdx.x;
  if (pos >= P) return;
  const int yo = pos / p.Wo, xo = pos % p.Wo;
  const int HW = p.H * p.W;
  float acc[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) acc[o] = bias ? __ldg(bias + co0 + o) : 0.f;

  const int c_begin = wg * p.cin_g, c_end = c_begin + p.cin_g;
  const float* in_n = in + (int64_t)n * p.Cin * HW;
  const int ybase = yo * p.sh - p.ph, xbase = xo * p.sw - p.pw;
  // offset groups touched by this weight group (one, when groups is a multiple of the offset groups)
  for (int g = c_begin / p.ch_og; g <= (c_end - 1) / p.ch_og; ++g) {
    const int c_lo = max(c_begin, g * p.ch_og), c_hi = min(c_end, (g + 1) * p.ch_og);
    const float* off_g = offset + ((int64_t)n * p.og + g) * 2 * K * P + pos;
    const float* msk_g = mask ? mask + ((int64_t)n * p.og + g) * K * P + pos : nullptr;
    for (int k = 0; k < K; ++k) {
      const int i = k / p.kw, j = k - i * p.kw;
      const float oy = __ldg(off_g + (int64_t)(2 * k) * P), ox = __ldg(off_g + (int64_t)(2 * k + 1) * P);
      const float m = msk_g ? __ldg(msk_g + (int64_t)k * P) : 1.f;
      const float y = __fadd_rn((float)(ybase + i * p.dh), oy), x = __fadd_rn((float)(xbase + j * p.dw), ox);
      // bilinear_interpolate(): zero outside (-1, H) x (-1, W); corners outside the image contribute zero
      const bool inside = !(y <= -1.f || (float)p.H <= y || x <= -1.f || (float)p.W <= x);
      const float fy = floorf(y), fx = floorf(x);
      const int y0 = (int)fy, x0 = (int)fx, y1 = y0 + 1, x1 = x0 + 1;
      const float lh = __fsub_rn(y, fy), lw = __fsub_rn(x, fx);
      const float hh = __fsub_rn(1.f, lh), hw = __fsub_rn(1.f, lw);
      const bool vy0 = inside && y0 >= 0, vy1 = inside && y1 <= p.H - 1;
      const bool vx0 = x0 >= 0, vx1 = x1 <= p.W - 1;
      // invalid corners: weight 0 and a clamped (always readable) address
      const float w1 = (vy0 && vx0) ? __fmul_rn(hh, hw) : 0.f, w2 = (vy0 && vx1) ? __fmul_rn(hh, lw) : 0.f;
      const float w3 = (vy1 && vx0) ? __fmul_rn(lh, hw) : 0.f, w4 = (vy1 && vx1) ? __fmul_rn(lh, lw) : 0.f;
      const int yc0 = min(max(y0, 0), p.H - 1), yc1 = min(max(y1, 0), p.H - 1);
      const int xc0 = min(max(x0, 0), p.W - 1), xc1 = min(max(x1, 0), p.W - 1);
      const int o1 = yc0 * p.W + xc0, o2 = yc0 * p.W + xc1, o3 = yc1 * p.W + xc0, o4 = yc1 * p.W + xc1;
      const float* ws = s_w + ((c_lo - c_begin) * K + k) * CO;
#pragma unroll 4
      for (int c = c_lo; c < c_hi; ++c) {
        const float* pl = in_n + (int64_t)c * HW;
        const float v1 = __ldg(pl + o1), v2 = __ldg(pl + o2), v3 = __ldg(pl + o3), v4 = __ldg(pl + o4);
        float val = __fmul_rn(w1, v1);
        val = __fmaf_rn(w2, v2, val);
        val = __fmaf_rn(w3, v3, val);
        val = __fmaf_rn(w4, v4, val);
        val = __fmul_rn(m, val);
        float wv[CO];
        if (CO % 4 == 0) {
#pragma unroll
          for (int q = 0; q < CO / 4; ++q) {
            const float4 t = reinterpret_cast<const float4*>(ws)[q];
            wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
          }
        } else if (CO % 2 == 0) {
#pragma unroll
          for (int q = 0; q < CO / 2; ++q) {
            const float2 t = reinterpret_cast<const float2*>(ws)[q];
            wv[2 * q] = t.x; wv[2 * q + 1] = t.y;
          }
        } else {
#pragma unroll
          for (int o = 0; o < CO; ++o) wv[o] = ws[o];
        }
#pragma unroll
        for (int o = 0; o < CO; ++o) acc[o] = __fmaf_rn(wv[o], val, acc[o]);
        ws += K * CO;
      }
    }
  }
  float* op = out + ((int64_t)n * p.Cout + co0) * P + pos;
#pragma unroll
  for (int o = 0; o < CO; ++o) op[(int64_t)o * P] = acc[o];
}

template <int CO>
static int launch_dcn(const float* in, const float* offset, const float* mask, const float* weight, const float* bias,
                      float* out, DcnParams p, cudaStream_t st) {
  p.chunks = p.cout_g / CO;
  const int P = p.Ho * p.Wo;
  const size_t smem = (size_t)CO * p.cin_g * p.kh * p.kw * sizeof(float);
  if (smem > 48 * 1024) {
    set_error("deform_conv2d_f32: %zu B of weights per block exceed 48 KB", smem);
    return B200VC_EUNSUPPORTED;
  }
  dim3 grid((P + kDcnThreads - 1) / kDcnThreads, p.groups * p.chunks, p.N);
  deform_conv2d_kernel<CO><<<grid, kDcnThreads, smem, st>>>(in, offset, mask, weight, bias, out, p);
  return check_launch("deform_conv2d_f32");
}

}  // namespace b200vc

using namespace b200vc;

extern "C" int b200vc_deform_conv2d_f32(const float* input, const float* offset, const float* mask,
                                        const float* weight, const float* bias, float* out, int N, int Cin, int H,
                                        int W, int Cout, int kh, int kw, int stride_h, int stride_w, int pad_h,
                                        int pad_w, int dil_h, int dil_w, int groups, int offset_groups, void* stream) {
  B200VC_REQUIRE(input && offset && weight && out, "deform_conv2d_f32: null pointer");
  B200VC_REQUIRE(N > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "deform_conv2d_f32: bad shape");
  B200VC_REQUIRE(kh > 0 && kw > 0 && stride_h > 0 && stride_w > 0 && dil_h > 0 && dil_w > 0 && pad_h >= 0 && pad_w >= 0,
                 "deform_conv2d_f32: bad kernel geometry");
  B200VC_REQUIRE(groups > 0 && Cin % groups == 0 && Cout % groups == 0, "deform_conv2d_f32: groups=%d does not divide "
                 "Cin=%d / Cout=%d", groups, Cin, Cout);
  B200VC_REQUIRE(offset_groups > 0 && Cin % offset_groups == 0, "deform_conv2d_f32: offset_groups=%d does not divide "
                 "Cin=%d", offset_groups, Cin);
  DcnParams p;
  p.N = N; p.Cin = Cin; p.H = H; p.W = W; p.Cout = Cout;
  p.kh = kh; p.kw = kw; p.sh = stride_h; p.sw = stride_w; p.ph = pad_h; p.pw = pad_w; p.dh = dil_h; p.dw = dil_w;
  p.Ho = (H + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
  p.Wo = (W + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
  B200VC_REQUIRE(p.Ho > 0 && p.Wo > 0, "deform_conv2d_f32: empty output");
  B200VC_REQUIRE((int64_t)H * W < (1ll << 31) && (int64_t)p.Ho * p.Wo < (1ll << 31), "deform_conv2d_f32: plane too large");
  B200VC_REQUIRE(N <= 65535 && (int64_t)groups * (Cout / groups) <= 65535, "deform_conv2d_f32: grid too large");
  p.groups = groups; p.og = offset_groups;
  p.cin_g = Cin / groups; p.cout_g = Cout / groups; p.ch_og = Cin / offset_groups;
  p.chunks = 1;
  cudaStream_t st = (cudaStream_t)stream;
  // all of a group's output channels in one thread when their count is one of the instantiated sizes (the
  // reference's layers: 4, 6, 8, 12), otherwise in slices (the samples are then recomputed per slice)
  const int cg = p.cout_g;
  if (cg == 16) return launch_dcn<16>(input, offset, mask, weight, bias, out, p, st);
  if (cg == 12) return launch_dcn<12>(input, offset, mask, weight, bias, out, p, st);
  if (cg == 8) return launch_dcn<8>(input, offset, mask, weight, bias, out, p, st);
  if (cg == 6) return launch_dcn<6>(input, offset, mask, weight, bias, out, p, st);
  if (cg % 4 == 0) return launch_dcn<4>(input, offset, mask, weight, bias, out, p, st);
  if (cg % 3 == 0) return launch_dcn<3>(input, offset, mask, weight, bias, out, p, st);
  if (cg % 2 == 0) return launch_dcn<2>(input, offset, mask, weight, bias, out, p, st);
  return launch_dcn<1>(input, offset, mask, weight, bias, out, p, st);
}
