/*
 * b200vc.h -- C-ABI of the B200-native hot path for the KUIS-AI LHBDC / Flex-Rate B-frame codecs.
 *
 * The reference (KUIS-AI-Tekalp-Research-Group/video-compression) has NO FFI layer: its hot path is a
 * chain of eager torch / CompressAI calls.  Each entry point below replaces one such chain; the comment
 * above it cites the reference interface (file:line, relative to the reference root) it stands in for.
 * The Python binding a maintainer adds on the reference side is the ctypes stub in INTEGRATION.md
 * (shipped as video-compression_b200/b200vc/_lib.py).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless marked [host].
 *   - tensors are fp32, NCHW, with H*W contiguous per channel plane (channel stride = H*W).  Where a
 *     tensor may be a channel slice of a wider tensor the batch stride is passed explicitly (elements).
 *   - the library never allocates, never synchronises and keeps no mutable global state; every call is
 *     an asynchronous launch on `stream` (a cudaStream_t passed as void*), CUDA-graph capturable.
 *   - return value: 0 on success, a negative B200VC_E* code otherwise; b200vc_last_error() returns a
 *     thread-local message for the last failing call.
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns B200VC_ECUDA.
 */
#ifndef B200VC_H_
#define B200VC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VC_VERSION 100

#if defined(__GNUC__)
#define B200VC_API __attribute__((visibility("default")))
#else
#define B200VC_API
#endif

enum {
  B200VC_OK = 0,
  B200VC_EINVAL = -1,   /* bad argument (null pointer, non-positive size, unsupported channel count) */
  B200VC_ECUDA = -2,    /* CUDA runtime error at launch                                              */
  B200VC_EUNSUPPORTED = -3
};

/* Warp geometry variants. */
enum {
  B200VC_WARP_LHBDC = 0, /* LHBDC/model/m.py:111-126, LHBDC/model/flow.py:15-25: pixel-centre linspace grid,
                            flow/((W-1)/2), grid_sample(bilinear, border, align_corners=False)              */
  B200VC_WARP_FLEX = 1,  /* Flex-Rate.../b_model/b_model.py:99-112: int grid + flow, 2*(x/W-0.5),
                            grid_sample defaults (bilinear, zeros, align_corners=False)                      */
  B200VC_WARP_AC1 = 2    /* ICIP2024/src/model/m.py:262-282, helpers.py:61-69, OJSP2025/video_model.py:668-676:
                            linspace(-1,1) grid, grid_sample(bilinear, border, align_corners=True)           */
};

/* Arithmetic-form switches for the coordinate path (bit-parity experiments; default 0 = the form torch's
 * CUDA kernels use: reciprocal-multiply for tensor/python-scalar division, FMA-contracted unnormalise). */
enum {
  B200VC_ARITH_NO_FMA = 1,   /* unnormalise as mul-then-sub (ATen CPU form is (g+1)*(W/2)-0.5) */
  B200VC_ARITH_TRUE_DIV = 2  /* flow / ((W-1)/2) as an IEEE division (ATen CPU form)                           */
};

/* Blend modes. */
enum {
  B200VC_BLEND_MASK = 0,   /* LHBDC/model/m.py:63-67: pred = m*fw + (1-m)*bw ; res = x - pred (mask has 1 channel) */
  B200VC_BLEND_NORMW = 1,  /* Flex-Rate.../b_model/b_model.py:68-73: mask = sigmoid(logits[2ch]); w=0.5*mask;
                              pred = (w1*xb + w2*xa)/(w1+w2+1e-8) ; res = x - pred                                */
  B200VC_BLEND_HALF = 2    /* ICIP2024/src/opt_helpers.py:35-36: pred = 0.5*w1 + (1-0.5)*w2                      */
};

B200VC_API int b200vc_version(void);
B200VC_API const char* b200vc_last_error(void);
/* Number of SMs of the current device. */
B200VC_API int b200vc_sm_count(void);
/* Number of CTAs (= number of fp64 partials) per sample the reducing entry points should be launched with
 * for `elems_per_sample` elements.  A function of the size only, never of the device, so the partial-sum
 * shapes -- and therefore the fp64 totals -- are identical on every GPU of a GOP-sharded run. */
B200VC_API int b200vc_reduce_blocks(int64_t elems_per_sample);

/* ---------------------------------------------------------------------------------------------- warp
 * Replaces Model.backwarp / flow.backwarp / BidirFlowRef.backwarp / FlowGuidedB.warp (see enum above).
 *   img  [N,C,H,W] (batch stride img_bs), flow [N,2,H,W] contiguous (ch0 = x, ch1 = y, pixels),
 *   out  [N,C,H,W] (batch stride out_bs; lets the caller write straight into a concat buffer).
 *   tab_x[W], tab_y[H]: the base grid exactly as the reference builds it with torch.linspace on the CPU
 *   (LHBDC/model/m.py:113-116; ICIP2024/src/model/m.py:264-265); ignored (may be NULL) for WARP_FLEX.
 */
B200VC_API int b200vc_warp_f32(const float* img, int64_t img_bs, const float* flow, const float* tab_x,
                    const float* tab_y, float* out, int64_t out_bs, int N, int C, int H, int W,
                    int variant, int arith, void* stream);

/* Fused LHBDC motion compensation: flow glue (LHBDC/model/m.py:55-59: chunk, + prior, crop, x4 bilinear
 * upsample) + both backward warps (m.py:61) + channel concat (m.py:63, torch.cat([fw, bw])).
 *   x_before, x_after [N,3,H,W] contiguous; flow_hat [N,4,h4,w4] (mv_compressor x_hat, padded quarter res);
 *   flow_ab, flow_ba [N,2,h4,w4]; the crop [:hh,:ww] with 4*hh == H, 4*ww == W;
 *   out [N,6,H,W] = cat(fw, bw); flows_out (nullable) [N,4,H,W] = cat(flow_cb_hat, flow_ca_hat).
 */
B200VC_API int b200vc_warp2_lhbdc_f32(const float* x_before, const float* x_after, const float* flow_hat,
                           const float* flow_ab, const float* flow_ba, const float* tab_x,
                           const float* tab_y, float* out, float* flows_out, int N, int H, int W, int h4,
                           int w4, int arith, void* stream);

/* Flex-Rate form of the fused motion compensation (Flex-Rate-Hier-Bidir-Video-Compression/b_model/b_model.py:34-45
 * `process`, :58-66 in `forward`): flow glue + both FLEX warps (zeros padding, half-pixel grid, :99-112) + the 16-channel
 * concat cat(ft0, ft1, x0, x1, warp(x0, ft0), warp(x1, ft1)) in one pass.
 *   mode 0 (linear motion): fa = Flow_0_1, fb = Flow_1_0 (each [N,2,H,W] with batch stride fa_bs / fb_bs: the two halves
 *     of the flow predictor's 4-channel output), ft0 = a0*fa + b0*fb, ft1 = a1*fa + b1*fb with the reference's
 *     coefficients a0 = -(1-t)t, b0 = t*t, a1 = (1-t)(1-t), b1 = -t(1-t) (python floats cast to fp32 by the caller);
 *   mode 1 (refinement): fa = cat(mv_before, mv_after), fb = flow_hat[:, 0:4] (each [N,4,H,W] with batch stride):
 *     ft0 = fa[0:2] + fb[0:2], ft1 = fa[2:4] + fb[2:4]; a0..b1 ignored.
 *   x0, x1 [N,3,H,W] contiguous; out16 [N,16,H,W] contiguous.
 */
B200VC_API int b200vc_warp2_flex_f32(const float* x0, const float* x1, const float* fa, int64_t fa_bs, const float* fb,
                                     int64_t fb_bs, int mode, float a0, float b0, float a1, float b1, float* out16,
                                     int N, int H, int W, void* stream);

/* Search form (ICIP2024/src/opt_helpers.py:23-51: prediction_flowonly + clamp + MSE per candidate down-ratio;
 * OJSP2025/video_model.py:621-666): both warps + 0.5/0.5 blend + clamp + squared error against x_cur in one pass.
 *   x1, x2, x_cur [N,3,H,W]; flow1, flow2 [N,2,H,W]; pred (nullable) [N,3,H,W];
 *   partials double[N * b200vc_warp2_half_sse_blocks(H, W)]: per-CTA SSE, reduce with b200vc_sum_partials_f64.
 */
B200VC_API int b200vc_warp2_half_sse_blocks(int H, int W);
B200VC_API int b200vc_warp2_half_sse_f32(const float* x1, const float* x2, const float* flow1, const float* flow2,
                                         const float* x_cur, const float* tab_x, const float* tab_y, float* pred,
                                         double* partials, double* totals, int32_t* counters, int N, int H, int W,
                                         int variant, void* stream);

/* Single-reference search form (OJSP2025/video_model.py:621-666: x_hat = self.warp(ref_frame, est_mv);
 * PSNR(x, x_hat) per candidate down-sampling ratio): warp + squared error against x_cur, no clamp.
 *   img, x_cur [N,3,H,W]; flow [N,2,H,W]; pred (nullable) [N,3,H,W] receives the warped frame;
 *   partials double[N * b200vc_warp2_half_sse_blocks(H, W)] as above.
 */
B200VC_API int b200vc_warp_sse_f32(const float* img, const float* flow, const float* x_cur, const float* tab_x,
                                   const float* tab_y, float* pred, double* partials, double* totals,
                                   int32_t* counters, int N, int H, int W, int variant, void* stream);

/* ------------------------------------------------------------------------------ SPyNet glue
 * Replaces the non-convolutional part of Network.forward (LHBDC/model/flow.py:78-101).
 *
 * spynet_pyramid: Preprocess (flow.py:39-44: channel order reversed, (x - mean) / std) when `preprocess` != 0, then
 *   `n_levels` (0..5) avg_pool2d(2, 2, count_include_pad=False) poolings (flow.py:83-88) in one pass.
 *   frame [N,3,H,W] (batch stride frame_bs); levels = HOST array of n_levels + 1 device pointers, levels[l] =
 *   [N,3,H>>l,W>>l] contiguous (floor division per level, as avg_pool2d); levels[0] is written only when
 *   `preprocess` != 0 (otherwise the frame itself is level 0 and levels[0] may be NULL).
 * spynet_level: one pyramid level's conv input (flow.py:93-98)
 *   feat [N,8,H,W] = cat(first, backwarp(second, up), up),
 *   up = interpolate(flow_prev [N,2,hp,wp], scale_factor=2, bilinear, align_corners=True) * 2.0, replicate-padded by
 *   one row / column when H == 2*hp + 1 / W == 2*wp + 1.  flow_prev == NULL means the all-zero initial flow.
 *   tab_x / tab_y as for warp_f32 (WARP_LHBDC). */
B200VC_API int b200vc_spynet_pyramid_f32(const float* frame, int64_t frame_bs, float* const* levels, int N, int H,
                                         int W, int n_levels, int preprocess, void* stream);
B200VC_API int b200vc_spynet_level_f32(const float* first, int64_t first_bs, const float* second, int64_t second_bs,
                                       const float* flow_prev, const float* tab_x, const float* tab_y, float* feat,
                                       int N, int H, int W, int hp, int wp, void* stream);

/* ---------------------------------------------------------------- deformable convolution
 * Replaces torchvision.ops.deform_conv2d (modulated, v2) behind DeformConv2d at ICIP2023/src/model/m.py:29-34 and
 * ICIP2024/src/model/helpers.py:40,57.  Same tensor layouts as torchvision:
 *   input  [N,Cin,H,W]; weight [Cout,Cin/groups,kh,kw]; bias [Cout] or NULL;
 *   offset [N, offset_groups*kh*kw*2, Ho, Wo]  (per offset group and kernel point: dy, dx);
 *   mask   [N, offset_groups*kh*kw, Ho, Wo] or NULL (v1);
 *   out    [N,Cout,Ho,Wo], Ho = (H + 2*pad_h - (dil_h*(kh-1)+1)) / stride_h + 1;
 *   workspace (nullable) float[N*Cin*H*W], 16-byte aligned: scratch for the group-channels-last copy of the input
 *   that the fast path samples from (used when Cin/groups is 4, 8, 12 or 16 and Cout/groups is 4, 6, 8, 12 or 16;
 *   without it, or for other shapes, the NCHW gather kernel runs).
 * No im2col matrix is materialised. */
B200VC_API int b200vc_deform_conv2d_f32(const float* input, const float* offset, const float* mask,
                                        const float* weight, const float* bias, float* out, float* workspace, int N,
                                        int Cin, int H, int W, int Cout, int kh, int kw, int stride_h, int stride_w, int pad_h,
                                        int pad_w, int dil_h, int dil_w, int groups, int offset_groups, void* stream);

/* ------------------------------------------------------- checkerboard context glue
 * ICIP2024/src/model/compression_bottlenecks.py:237-268 (and :479-510, ICIP2023/src/model/elic.py).
 * round_checker: y_hat = ste_round(y) = (round(y) - y) + y over the whole latent [N,C,H,W]; y_half = y_hat with the
 *   anchor positions ((h + w) even: [0::2,0::2] and [1::2,1::2]) set to zero.  Either output may be NULL.
 * checker_mask: dst = src with the positions of parity `zero_parity` ((h + w) & 1) set to zero; src / dst are
 *   [N,C,H,W] blocks with batch strides (dst may be a channel slice of a concat buffer, or src itself).
 *   The reference zeroes [0::2,1::2] and [1::2,0::2] of the context convolution's output: zero_parity = 1. */
B200VC_API int b200vc_round_checker_f32(const float* y, float* y_hat, float* y_half, int N, int C, int H, int W,
                                        void* stream);
B200VC_API int b200vc_checker_mask_f32(const float* src, int64_t src_bs, float* dst, int64_t dst_bs, int N, int C, int H,
                                       int W, int zero_parity, void* stream);

/* ------------------------------------------------------------------------------------ blend / residual
 * Replaces LHBDC/model/m.py:63-67, Flex-Rate.../b_model/b_model.py:68-73, ICIP2024/src/opt_helpers.py:35-45.
 *   a, b: the two warped references [N,3,H,W] (batch strides a_bs, b_bs: may be halves of the concat
 *   buffer); mask [N,1|2,H,W] (NULL for BLEND_HALF); x_cur [N,3,H,W]; pred, res (each nullable) [N,3,H,W].
 *   sse_partials (nullable, double[N*n_blocks]) receives per-CTA sums of (clamp(pred,0,1) - x_cur)^2 -- the
 *   search form only needs that scalar; reduce with b200vc_sum_partials_f64, or pass sse_totals / counters (see
 *   "fused finish" below).  n_blocks = CTAs per sample.
 */
B200VC_API int b200vc_blend_residual_f32(int mode, const float* mask, const float* a, int64_t a_bs, const float* b,
                              int64_t b_bs, const float* x_cur, float* pred, float* res,
                              double* sse_partials, int n_blocks, double* sse_totals, int32_t* counters, int N,
                              int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------ GDN / IGDN
 * Replaces compressai.layers.GDN.forward (instantiated at LHBDC/model/layers.py:49-53,84-88,124-128,159-163).
 * gdn_prepare applies CompressAI's NonNegativeParametrizer to the stored parameters once per weight
 * version:  beta_eff = max(beta, beta_bound)^2 - pedestal ; gamma_eff = max(gamma, gamma_bound)^2 - pedestal.
 * params_out layout (floats): [0,C) beta_eff | [C, C+C*C) gamma_eff row-major [i][j] | [.., +C*C) its transpose |
 *   [.., +2*C*C) the tf32 hi / lo split of gamma_eff (row-major), the tensor-core kernel's A operand.
 * b200vc_gdn_params_floats(C) gives the size.
 */
B200VC_API int64_t b200vc_gdn_params_floats(int C);
B200VC_API int b200vc_gdn_prepare_f32(const float* beta, const float* gamma, float beta_bound, float gamma_bound,
                           float pedestal, float* params_out, int C, void* stream);
/*   x, out [N,C,HW]; addend (nullable) [N,C,HW] is added to the result (the residual-block skip,
 *   compressai ResidualBlockWithStride/ResidualBlockUpsample `out += identity`); out must not alias x; out == addend is allowed and
 *   is the fast path (in-place residual add: the tensor-core kernel accumulates with a TMA reduce-add).
 *   out_i = x_i * rsqrt(beta_i + sum_j gamma_ij x_j^2)    (inverse == 1: * sqrt; inverse == 2, diagnostics:
 *   out_i = the norm beta_i + sum_j gamma_ij x_j^2 itself).
 *   impl: 0 = auto (2 when C in {128, 192} and HW % 4 == 0, else 1), 1 = CUDA-core fp32 kernel (C in {64, 128, 192};
 *   bit-identical to the reference's fp32 chain), 2 = tcgen05 3xTF32 kernel (C == 128: gamma resident in TMEM;
 *   C == 192: output channels split over CTA pairs; <= 2.5e-6 relative: split products plus raw MUFU rsqrt / sqrt on
 *   the always-positive norm).  Other C return B200VC_EUNSUPPORTED.
 */
B200VC_API int b200vc_gdn_f32(const float* x, const float* params, const float* addend, float* out, int N, int C,
                   int64_t HW, int inverse, int impl, void* stream);

/* ----------------------------------------------------------------------- Gaussian conditional (Q1, Q2, Q4, Q5)
 * Replaces GaussianConditional.forward / quantize("symbols") / build_indexes and the torch.log(lik).sum()
 * bit sums (LHBDC/model/layers.py:102-103; LHBDC/model/m.py:73-91; Flex-Rate.../b_model/layers.py:145-146).
 *   y [N,C,HW] contiguous; scales, means [N,C,HW] with batch stride sm_bs (they are the two channel chunks
 *   of h_s's [N,2C,H,W] output: sm_bs = 2*C*HW);
 *   inv_gain (nullable) [C]: y_hat is multiplied per channel on the way out (Flex inv_gain_unit, layers.py:146);
 *   y_hat, lik (nullable) [N,C,HW]; symbols, indexes (nullable int32) [N,C,HW];
 *   scale_table [n_table] (needed iff indexes != NULL);
 *   bits_partials (nullable) double[N * blocks_per_sample]: partial sums of -log2(lik).  Slot (n, b) covers a
 *   contiguous 1/blocks_per_sample of sample n (whole 1024-element chunks); its value depends on those elements
 *   only, never on the launch geometry (r2: a persistent grid walks the slots).  Any blocks_per_sample >= 1 works;
 *   b200vc_reduce_blocks(C*HW / 4) (4096 elements per slot) is what the Python wrapper passes.
 *   Fused finish (every kernel that writes per-CTA partials has this pair): totals (nullable) double[N] + counters
 *   int32[N], counters ZERO on entry.  The CTA that finishes last for a sample sums that sample's partials in the
 *   same fixed order as b200vc_sum_partials_f64 (bit-identical result, independent of scheduling), writes totals[n]
 *   and leaves the counter zero again -- the separate reduction launch (the reference's `torch.log(lik).sum()`
 *   second pass, LHBDC/model/m.py:73-91) disappears.  totals == NULL keeps the two-launch form.
 */
B200VC_API int b200vc_gauss_cond_f32(const float* y, const float* scales, const float* means, int64_t sm_bs,
                          const float* inv_gain, float* y_hat, float* lik, int32_t* symbols,
                          int32_t* indexes, const float* scale_table, int n_table, float scale_bound,
                          float lik_bound, double* bits_partials, int blocks_per_sample, double* bits_totals,
                          int32_t* counters, int N, int C, int64_t HW, void* stream);

/* ------------------------------------------------------------------------ factorised prior (Q3, Q4, Q5)
 * Replaces EntropyBottleneck.forward (eval) (LHBDC/model/layers.py:97-98 call sites).
 * eb_prepare packs, per channel, softplus(_matrix{0..4}), _bias{0..4}, tanh(_factor{0..3}) and the median
 * into 59 floats (layout in DESIGN.md); inputs are the raw CompressAI parameters, filters (3,3,3,3).
 */
#define B200VC_EB_PARAMS_PER_CHANNEL 59
B200VC_API int b200vc_eb_prepare_f32(const float* const* matrices /*[host] 5 device ptrs*/,
                          const float* const* biases /*[host] 5*/, const float* const* factors /*[host] 4*/,
                          const float* quantiles /*[C,1,3]*/, float* packed /*[C,59]*/, int C, void* stream);
/*   z [N,C,HW]; gain / inv_gain (nullable) [C] (Flex hyper_gain_unit / hyper_inv_gain_unit);
 *   z_hat, lik (nullable) [N,C,HW]; symbols (nullable int32): round(z*gain - median);
 *   bits_partials / bits_totals / counters as above.
 */
B200VC_API int b200vc_entropy_bottleneck_f32(const float* z, const float* packed, const float* gain,
                                  const float* inv_gain, float* z_hat, float* lik, int32_t* symbols,
                                  float lik_bound, double* bits_partials, int blocks_per_sample,
                                  double* bits_totals, int32_t* counters, int N, int C, int64_t HW, void* stream);

/* out[s] = sum_{k < n_per} partials[s*n_per + k] in fixed order (deterministic: 1-GPU and N-GPU totals agree). */
B200VC_API int b200vc_sum_partials_f64(const double* partials, int n_per, int n_out, double* out, void* stream);

/* -------------------------------------------------------------------------------------------- metrics
 * Replaces the D2H + numpy PSNR of LHBDC/test/testing.py:176-182: sum over the unpadded crop [:h,:w] of
 * (round(clip(a)*255) - round(clip(b)*255))^2 PER SAMPLE, as partials[N][n_blocks] per-CTA doubles (exact integers;
 * reduce with b200vc_sum_partials_f64(partials, n_blocks, N, out), or pass totals / counters: fused finish).
 */
B200VC_API int b200vc_sse_u8_f32(const float* a, const float* b, double* partials, int n_blocks, double* totals,
                      int32_t* counters, int N, int C, int H, int W, int h, int w, void* stream);

/* ------------------------------------------------------------------------------------- rANS
 * Replaces the CPU coder behind `.compress()` / `.decompress()` (compressai.ans BufferedRansEncoder.encode_with_indexes
 * / RansDecoder.decode_with_indexes; LHBDC/model/layers.py:93-117, LHBDC/encode_B.py:96,104).  The model side is
 * CompressAI's (16-bit quantised CDF rows, per-symbol row index, offset, tail-bin escape); the container is b200vc's
 * own ("b2r1": independent streams of stream_len symbols, one thread each) -- see DESIGN.md.
 *   symbols, indexes int32 [n_symbols]; cdf int32 [rows, cdf_stride]; cdf_len, offset int32 [rows].
 *   encode: scratch uint16 [n_streams, b200vc_rans_scratch_words(stream_len)], sizes_words int32 [n_streams] (out).
 *   compact: offsets_words int64 [n_streams] = exclusive prefix sum of sizes_words; out uint16 [sum(sizes)].
 *   decode: payload / offsets_words / sizes_words as produced by encode + compact (they come from an untrusted
 *           container: no stream reads past its own words); status int32 [1] (may be NULL), zeroed by the caller,
 *           becomes 1 when a stream over-ran, stopped short or did not return to the initial coder state.
 */
B200VC_API int b200vc_rans_scratch_words(int stream_len);
B200VC_API int b200vc_rans_encode(const int32_t* symbols, const int32_t* indexes, const int32_t* cdf,
                                  const int32_t* cdf_len, const int32_t* offset, int cdf_stride, int64_t n_symbols,
                                  int stream_len, uint16_t* scratch, int32_t* sizes_words, void* stream);
B200VC_API int b200vc_rans_compact(const uint16_t* scratch, int stream_len, const int32_t* sizes_words,
                                   const int64_t* offsets_words, int n_streams, uint16_t* out, void* stream);
B200VC_API int b200vc_rans_decode(const uint16_t* payload, const int64_t* offsets_words, const int32_t* sizes_words,
                                  const int32_t* indexes, const int32_t* cdf, const int32_t* cdf_len,
                                  const int32_t* offset, int cdf_stride, int64_t n_symbols, int stream_len,
                                  int32_t* symbols_out, int32_t* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200VC_H_ */
