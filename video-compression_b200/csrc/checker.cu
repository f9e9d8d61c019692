// K-CHK: checkerboard context glue of the ICIP ELIC-style compressors.
//
// ICIP2024/src/model/compression_bottlenecks.py:237-268 (Offset_ELIC; Res_ELIC :479-510 and ICIP2023's elic.py are the
// same loop): per channel group the reference re-quantises (`ste_round`), clones, zero-fills the anchor positions with
// two strided writes, runs the masked 5x5 convolution, zero-fills the non-anchor positions of its output with two
// more strided writes, re-quantises the concatenation of all earlier groups, and concatenates the parameter inputs.
// The glue is ~40 tiny torch kernels per compressor call.  Here:
//   round_checker  : y -> ste_round(y) for ALL channels and its anchor-zeroed copy, once per call (the groups and the
//                    "earlier groups" tensors are channel slices of these two);
//   checker_mask   : conv output -> checkerboard-zeroed copy written straight into a channel slice of the
//                    entropy-parameter network's input buffer (or in place).
// Arithmetic: ste_round(x) = (round(x) - x) + x with round = round-half-to-even (torch.round), evaluated exactly as
// written (the result equals round(x) except that a negative zero becomes +0, as in the reference).
#include "common.cuh"

namespace b200vc {

__device__ __forceinline__ float ste_round_f(float x) { return __fadd_rn(__fsub_rn(rintf(x), x), x); }

__global__ void __launch_bounds__(256)
round_checker_kernel(const float* __restrict__ y, float* __restrict__ y_hat, float* __restrict__ y_half, int W, int HW,
                     int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int p = (int)(i % HW);
  const int h = p / W, w = p - h * W;
  const float r = ste_round_f(__ldg(y + i));
  if (y_hat != nullptr) y_hat[i] = r;
  // y_half[:, :, 0::2, 0::2] = 0 ; y_half[:, :, 1::2, 1::2] = 0   (compression_bottlenecks.py:240-242)
  if (y_half != nullptr) y_half[i] = ((h + w) & 1) ? r : 0.f;
}

__global__ void __launch_bounds__(256)
checker_mask_kernel(const float* __restrict__ src, int64_t src_bs, float* __restrict__ dst, int64_t dst_bs, int W, int HW,
                    int64_t chw, int zero_parity) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;  // index inside one sample's [C,H,W] block
  if (i >= chw) return;
  const int n = blockIdx.y;
  const int p = (int)(i % HW);
  const int h = p / W, w = p - h * W;
  const float v = __ldg(src + (int64_t)n * src_bs + i);
  dst[(int64_t)n * dst_bs + i] = (((h + w) & 1) == zero_parity) ? 0.f : v;
}

}  // namespace b200vc

using namespace b200vc;

extern "C" int b200vc_round_checker_f32(const float* y, float* y_hat, float* y_half, int N, int C, int H, int W,
                                        void* stream) {
  B200VC_REQUIRE(y && (y_hat || y_half), "round_checker_f32: null pointer");
  B200VC_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, "round_checker_f32: bad shape");
  B200VC_REQUIRE((int64_t)H * W < (1ll << 31), "round_checker_f32: plane too large");
  const int64_t total = (int64_t)N * C * H * W;
  const int64_t blocks = (total + 255) / 256;
  B200VC_REQUIRE(blocks < (1ll << 31), "round_checker_f32: tensor too large");
  round_checker_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(y, y_hat, y_half, W, H * W, total);
  return check_launch("round_checker_f32");
}

extern "C" int b200vc_checker_mask_f32(const float* src, int64_t src_bs, float* dst, int64_t dst_bs, int N, int C, int H,
                                       int W, int zero_parity, void* stream) {
  B200VC_REQUIRE(src && dst, "checker_mask_f32: null pointer");
  B200VC_REQUIRE(N > 0 && N <= 65535 && C > 0 && H > 0 && W > 0, "checker_mask_f32: bad shape");
  B200VC_REQUIRE(zero_parity == 0 || zero_parity == 1, "checker_mask_f32: zero_parity must be 0 or 1");
  B200VC_REQUIRE((int64_t)H * W < (1ll << 31), "checker_mask_f32: plane too large");
  const int64_t chw = (int64_t)C * H * W;
  const int64_t blocks = (chw + 255) / 256;
  B200VC_REQUIRE(blocks < (1ll << 31), "checker_mask_f32: tensor too large");
  checker_mask_kernel<<<dim3((unsigned)blocks, N), 256, 0, (cudaStream_t)stream>>>(src, src_bs, dst, dst_bs, W, H * W, chw,
                                                                                zero_parity);
  return check_launch("checker_mask_f32");
}
