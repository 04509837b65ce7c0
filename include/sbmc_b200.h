/*
 * sbmc_b200.h -- C ABI of the B200-native SBMC kernel-splatting library
 * (libsbmc_b200.so, built from sbmc_b200/csrc/ with nvcc for sm_100a).
 *
 * These entry points are what a binding for the reference's native-op module
 * `sbmc.halide_ops` would call.  The reference binds six functions
 * (reference setup.py:65-84, pybind `m.def` synthesised at
 * halide_pytorch/halide_pytorch/extension.py:168-173) whose arguments are the
 * Halide Input<>s then Output<>s in declaration order:
 *
 *   scatter2gather_{cpu,cuda}_float32(weights, output)
 *       src/scatter2gather.cpp:61-62, called at sbmc/functions.py:56-59,67-70
 *   kernel_weighting_{cpu,cuda}_float32(data, weights, output, sum_w)
 *       src/kernel_weighting.cpp:130-133, called at sbmc/functions.py:95-98
 *   kernel_weighting_grad_{cpu,cuda}_float32(data, weights, sum_w, d_output,
 *                                            d_sum_w, d_data, d_weights)
 *       src/kernel_weighting.cpp:195-202, called at sbmc/functions.py:109-114
 *
 * Conventions (all functions):
 *   - plain C, no torch / Halide types; fp32, contiguous, torch index order:
 *       data[n][c][y][x]   weights[n][dy][dx][y][x]   sum_w[n][y][x]
 *   - the caller owns and allocates every buffer; outputs are fully
 *     overwritten (they may be uninitialised on entry, as in
 *     sbmc/functions.py:53-54,91-94,105-108); nothing is retained.
 *   - `stream` is a cudaStream_t (NULL = legacy default stream); device
 *     entry points are asynchronous with respect to the host.
 *   - return 0 on success, a negative SBMC_E* code otherwise; never throws.
 *     sbmc_b200_last_error() gives a thread-local message for the last failure.
 *   - kernel sizes: any kh, kw >= 1 (centre (k-1)/2, floor, as in
 *     src/kernel_weighting.cpp:53-54), any c >= 1, 64-bit element counts.
 *   - zero is read outside the image for data, weights and d_output
 *     (BoundaryConditions::constant_exterior, src/kernel_weighting.cpp:35-39,
 *     78-85; src/scatter2gather.cpp:34-35).
 */
#ifndef SBMC_B200_H_
#define SBMC_B200_H_

#include <stdint.h>

#if defined(__GNUC__)
#define SBMC_API __attribute__((visibility("default")))
#else
#define SBMC_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define SBMC_OK 0
#define SBMC_EINVAL (-1)   /* bad argument (null pointer, negative size, ...) */
#define SBMC_ECUDA (-2)    /* a CUDA runtime / driver call failed             */
#define SBMC_ENODEV (-3)   /* no sm_100 device / driver available             */
#define SBMC_EALIGN (-4)   /* pointer not 4-byte aligned                      */
#define SBMC_EUNSUPPORTED (-5) /* no fused kernel for this shape: the caller
                                  composes the individual operators instead   */

/* Library version (major*10000 + minor*100 + patch). */
SBMC_API int sbmc_b200_version(void);

/* Message for the last error on the calling thread ("" if none). */
SBMC_API const char *sbmc_b200_last_error(void);

/* 0: pick the tuned sm_100a kernels when the shape allows (default);
 * 1: always run the shape-generic kernels (used by the parity tests to check
 *    both implementations against the oracle).  Returns the previous value. */
SBMC_API int sbmc_b200_force_generic(int flag);

/* Which implementation the last call on this thread used:
 * 0 = none yet, 1 = tuned (TMA / vectorised) kernels, 2 = generic kernels. */
SBMC_API int sbmc_b200_last_path(void);

/* Number of kernels this library has launched since load (all threads). */
SBMC_API int64_t sbmc_b200_launch_count(void);

/* Per-kernel device timing for benchmarks.  While enabled, each kernel the
 * library launches is bracketed by CUDA events on its stream.
 * sbmc_b200_timing_collect waits for them and writes, per kernel kind, the
 * summed device time in ms and the number of launches since the last collect
 * (arrays of SBMC_NUM_KERNEL_KINDS entries; either may be NULL). */
#define SBMC_KERNEL_KW_FWD 0       /* KernelWeighting forward                */
#define SBMC_KERNEL_KW_DWEIGHTS 1  /* backward, d_weights (write stream)     */
#define SBMC_KERNEL_KW_DDATA 2     /* backward, d_data (reads the weights)   */
#define SBMC_KERNEL_S2G 3          /* Scatter2Gather                         */
#define SBMC_KERNEL_OTHER 4        /* halo adds of the host pipeline         */
#define SBMC_KERNEL_SPLAT_FWD 5    /* fused ProgressiveKernelApply forward   */
#define SBMC_KERNEL_SPLAT_BWD 6    /* fused ProgressiveKernelApply backward  */
#define SBMC_KERNEL_CONV1X1 7      /* fused 3-layer 1x1 ConvChain (tcgen05)  */
#define SBMC_KERNEL_TILES 8        /* tile reader: LZ4 inflate + assembly    */
#define SBMC_KERNEL_OPTIM 9        /* fused gradient clipping + Adam         */
#define SBMC_KERNEL_CONV3X3 10     /* 3x3 implicit-GEMM convolution (tcgen05) */
#define SBMC_NUM_KERNEL_KINDS 11
SBMC_API int sbmc_b200_timing_enable(int flag);
SBMC_API int sbmc_b200_timing_collect(double *ms_by_kind, int64_t *launches_by_kind);

/* ---- device-pointer entry points (replace the *_cuda_float32 ops) -------- */

/* gather[n][dy][dx][y][x] = scatter[n][kh-1-dy][kw-1-dx][y+dy-c0h][x+dx-c0w]
 * (0 if the source pixel is outside the image).  Bit-exact copy.
 * Replaces scatter2gather_cuda_float32 (src/scatter2gather.cpp:28-52). */
SBMC_API int sbmc_scatter2gather_f32(const float *scatter, float *gather, int64_t n,
                            int kh, int kw, int64_t h, int64_t w, void *stream);

/* output[n][c][y][x] = sum_{dy,dx} weights[n][dy][dx][y][x]
 *                                  * data[n][c][y+dy-c0h][x+dx-c0w]
 * sum_w[n][y][x]     = sum_{dy,dx} weights[n][dy][dx][y][x]   (all taps)
 * Replaces kernel_weighting_cuda_float32 (src/kernel_weighting.cpp:27-64). */
SBMC_API int sbmc_kernel_weighting_fwd_f32(const float *data, const float *weights,
                                  float *output, float *sum_w, int64_t n, int c,
                                  int64_t h, int64_t w, int kh, int kw,
                                  void *stream);

/* d_data[n][c][y][x]      = sum_{ry,rx} weights[n][kh-1-ry][kw-1-rx][y+ry-c0h][x+rx-c0w]
 *                                       * d_output[n][c][y+ry-c0h][x+rx-c0w]
 * d_weights[n][dy][dx][y][x] = d_sum_w[n][y][x]
 *                              + sum_c data[n][c][y+dy-c0h][x+dx-c0w] * d_output[n][c][y][x]
 * `sum_w` is accepted for signature parity and never read (it is unused by the
 * reference pipeline too, src/kernel_weighting.cpp:67-124); it may be NULL.
 * Replaces kernel_weighting_grad_cuda_float32. */
SBMC_API int sbmc_kernel_weighting_bwd_f32(const float *data, const float *weights,
                                  const float *sum_w, const float *d_output,
                                  const float *d_sum_w, float *d_data,
                                  float *d_weights, int64_t n, int c, int64_t h,
                                  int64_t w, int kh, int kw, void *stream);

/* ---- fused progressive splat (sbmc/modules.py:364-473, inference) -------- *
 * One ProgressiveKernelApply update in a single pass over the kernel logits:
 *   G = splat ? Scatter2Gather(kernels) : kernels        (never materialised)
 *   new_max = max(max_taps G, max_w);  scaler = exp(max_w - new_max)
 *   sum_r = sum_r * scaler + sum_taps exp(G - new_max) * data(tap)
 *   sum_w = sum_w * scaler + sum_taps exp(G - new_max);   max_w = new_max
 * kernels [n][kh*kw][h][w] are the raw logits (not modified), data [n][c][h][w];
 * sum_r [n][c][h][w], sum_w / max_w [n][h][w] are updated in place.  first != 0
 * is the initialisation step (sum_r = sum_w = max_w = None in the reference):
 * the state buffers are written without being read. */
SBMC_API int sbmc_progressive_splat_fwd_f32(const float *kernels, const float *data,
                                   float *sum_r, float *sum_w, float *max_w,
                                   int64_t n, int c, int64_t h, int64_t w, int kh,
                                   int kw, int splat, int first, void *stream);

/* Backward of one splat-mode update in scatter space (one read of the logits,
 * one write of their gradient).  planes [n][c+3][h][w] packs, per target pixel:
 * [0,c) dL/dsum_r', c: dL/dsum_w', c+1: the updated running max m', c+2: the
 * gradient routed to the tap maximum (0 where the max came from earlier samples).
 *   d_kernels[n][t][p] = e (g_w[q] + sum_c g_r[c][q] data[c][p]) + [kernels[t][p] == m'[q]] T[q]
 *   d_data[n][c][p]    = sum_t e g_r[c][q],   e = exp(kernels[t][p] - m'[q]),  q = p + off(t)
 * Returns SBMC_EUNSUPPORTED when no fused kernel exists for the shape (odd
 * kernels, c == 3, w % 4 == 0, 16-byte aligned pointers are required). */
SBMC_API int sbmc_progressive_splat_bwd_f32(const float *planes, const float *kernels,
                                   const float *data, float *d_kernels, float *d_data,
                                   int64_t n, int c, int64_t h, int64_t w, int kh,
                                   int kw, void *stream);

/* ---- fused per-sample 1x1 ConvChain on the tensor cores (inference) ------- *
 * y = W3 act(W2 act(W1 [xa ; xb] + b1) + b2) + b3 for every pixel: the 3-layer
 * 1x1 chains of sbmc/models.py:86-102 (hidden width 128).  xa [n][ca][hw] and
 * xb [n][cb][hw] (or [n][cb] broadcast over the pixels when b_broadcast != 0;
 * cb may be 0) are fp32 with `*_img_stride` elements between images; y is fp32
 * [n][cout][hw] with y_img_stride elements between images.  Weights are bf16,
 * row-major [out][in] with the input dimension zero-padded to k1p (128 or 256)
 * for w1 [128][k1p], w2 [128][128], w3 [n3p][128] (n3p = cout rounded up to a
 * multiple of 16, extra rows zero); biases fp32 (b3 has n3p entries).
 * act: 0 = ReLU, 1 = LeakyReLU(0.01) on the two hidden layers; the output is
 * linear.  bf16 operands, fp32 accumulation and fp32 output. */
SBMC_API int sbmc_conv1x1_chain_f32(const float *xa, int ca, int64_t a_img_stride,
                           const float *xb, int cb, int64_t b_img_stride,
                           int b_broadcast, const void *w1, const float *b1,
                           const void *w2, const float *b2, const void *w3,
                           const float *b3, int k1p, int cout, int n3p, int act,
                           float *y, int64_t y_img_stride, int64_t n_img, int64_t hw,
                           void *stream);

/* Same chain on bf16 channels-innermost activations (the inference pipeline's
 * layout): xa, xb are bf16 [n][hw][128] (xb may be NULL; `*_img_stride` elements
 * between images) and are pulled by TMA straight into the tensor-core operand
 * layout.  w1 is bf16 [128][128 or 256] (xa's channels first), b1 may differ per
 * image (b1_img_stride = 128) or be shared (0) -- this is how broadcast global
 * features enter.  Output: out_nhwc_bf16 != 0 -> bf16 [n][hw][128] (cout must be
 * 128), else fp32 [n][cout][hw]. */
SBMC_API int sbmc_conv1x1_chain_nhwc_bf16(const void *xa, int64_t a_img_stride, const void *xb,
                                 int64_t b_img_stride, const void *w1, const float *b1,
                                 int64_t b1_img_stride, const void *w2, const float *b2,
                                 const void *w3, const float *b3, int cout, int n3p,
                                 int act, void *y, int64_t y_img_stride,
                                 int out_nhwc_bf16, int64_t n_img, int64_t hw,
                                 void *stream);

/* The same chain for ALL SAMPLES of every pixel in one launch, software-pipelined
 * (csrc/chain_v3.cu; reference: the per-sample loops of sbmc/models.py:143-181
 * (embedding_XX + mean over spp) and :195-199 (kernel_regressor)).
 * feats: bf16 [n][spp_total][hw][128] (f_img_stride / f_smp_stride elements between
 * images / samples); prop: bf16 [n][hw][128] shared by the samples of a pixel, or
 * NULL (then w1 is [128][128], else [128][256] with the feature channels first).
 * Samples [sample0, sample0 + nsamples) are processed.
 * regress == 0 (embedding; cout = n3p = 128): out bf16 [n][spp_total][hw][128]
 *   (out_img_stride / out_smp_stride elements), and, when mean != NULL, mean over
 *   the nsamples samples of the chain output (taken on the fp32 accumulators) as
 *   bf16 (mean_f32 == 0) or fp32 [n][hw][128] -- models.py:181 `features.mean(1)`.
 * regress != 0: out fp32 [n][spp_total][cout][hw] logits (fp32 accumulate + store),
 *   mean must be NULL.  w3 bf16 [n3p][128], b3 fp32 [n3p], n3p multiple of 16. */
SBMC_API int sbmc_chain_samples_nhwc_bf16(
    const void *feats, int64_t f_img_stride, int64_t f_smp_stride, int64_t spp_total,
    const void *prop, int64_t p_img_stride, const void *w1, const float *b1,
    int64_t b1_img_stride, const void *w2, const float *b2, const void *w3, const float *b3,
    int cout, int n3p, int act, int regress, void *out, int64_t out_img_stride,
    int64_t out_smp_stride, void *mean, int64_t mean_img_stride, int mean_f32, int64_t n_img,
    int64_t sample0, int64_t nsamples, int64_t hw, void *stream);

/* 3x3 convolution, stride 1, zero padding 1 (the U-net convolutions of
 * sbmc/modules.py:248-320) as an implicit GEMM on tcgen05 (csrc/conv3x3.cu):
 * x bf16 [n][h][w][cin] (channels innermost), w9 bf16 [9][cout][cin] with tap
 * index 3 * dy + dx (out[y][x] += w9[3 dy + dx] . x[y + dy - 1][x + dx - 1]),
 * bias fp32 [cout]; y bf16 [n][h][w][cout] = act(conv + bias); act: 0 none,
 * 1 ReLU, 2 LeakyReLU(0.01).  cin multiple of 64, cout multiple of 128. */
SBMC_API int sbmc_conv3x3_nhwc_bf16(const void *x, const void *w9, const float *bias, void *y,
                                    int64_t n, int h, int w, int cin, int cout, int act,
                                    void *stream);
/* flag != 0 (the default): cout = 128 runs on the CTA-pair kernel
 * (tcgen05.mma.cta_group::2, 4-8 % faster); 0 selects the single-CTA kernel.  Returns
 * the previous value. */
SBMC_API int sbmc_b200_conv3x3_pair(int flag);
/* flag != 0 (the default): images at most 85 pixels wide run in the kernel's "linear" mode
 * (tiles of 256 consecutive pixels of the row-major image with one shared zero column per
 * row instead of 128-pixel row segments -- the coarse U-net levels of a training crop); 0
 * selects the row-segment tiling for every width.  Same results.  Returns the previous value. */
SBMC_API int sbmc_b200_conv3x3_linear(int flag);

/* U-net decoder glue (sbmc/modules.py:314-319): out = cat([bilinear_upsample(low,
 * size=(h, w), align_corners=False), skip], channels) in one pass on bf16
 * channels-innermost tensors: low [n][hl][wl][cu], skip [n][h][w][cs],
 * out [n][h][w][cu+cs]; cu and cs multiples of 8, 16-byte aligned pointers. */
SBMC_API int sbmc_upsample_concat_nhwc_bf16(const void *low, const void *skip, void *out,
                                   int64_t n, int hl, int wl, int h, int w, int cu,
                                   int cs, void *stream);

/* One 1x1-convolution layer as a tcgen05 GEMM (csrc/linear.cu): x bf16 [pixels][cin],
 * w bf16 [cout][cin], bias fp32 [cout] or NULL; y [pixels][cout] = act(x . w^T + bias) as
 * bf16 (out_f32 == 0) or fp32; act: 0 none, 1 ReLU, 2 LeakyReLU(0.01).  cin multiple of
 * 64, cout multiple of 128, 32-byte aligned pointers.  The layer-by-layer form of the
 * per-sample ConvChains (sbmc/modules.py:34-125) that the mixed-precision training path
 * needs (every layer output is kept for the backward pass). */
SBMC_API int sbmc_linear_nhwc_bf16(const void *x, const void *w, const float *bias, void *y,
                                   int64_t pixels, int cin, int cout, int act, int out_f32,
                                   void *stream);

/* General form of the layer above for the mixed-precision training pipeline
 * (sbmc_b200/train_pipeline.py; reference: the train branch of sbmc/models.py:171-209):
 *  - two sources: rows of x (bf16 [rows][cin_a], one row per SAMPLE, rows ordered image,
 *    sample, pixel) are concatenated with rows of xb (bf16 [rows / spp][cin_b], one row per
 *    PIXEL, shared by the spp samples of the pixel) -- the reference's
 *    `th.cat([features, propagated.unsqueeze(1).repeat(...)], 2)` (models.py:175-177,193)
 *    without materialising it; w is [cout][cin_a + cin_b]; needs hw % 256 == 0; cin_b = 0:
 *    single source;
 *  - mask != NULL (bf16 [rows][cout]): the result is multiplied by the derivative of
 *    ReLU (mask_act 1) / LeakyReLU(0.01) (mask_act 2) read off the sign of mask -- the
 *    data-gradient calls pass the saved activations of the previous layer;
 *  - out_mode 0: bf16 rows, 1: fp32 rows, 2: fp32 channel planes
 *    y[b * out_img_stride + s * out_smp_stride + c * hw + p] for c < cout_valid (the
 *    kernel regressor's logits in the layout the splat reads, models.py:195-199). */
SBMC_API int sbmc_linear2_nhwc_bf16(const void *x, int cin_a, const void *xb, int cin_b,
                                    int64_t hw, int64_t spp, const void *w, const float *bias,
                                    const void *mask, int mask_act, void *y, int out_mode,
                                    int64_t out_img_stride, int64_t out_smp_stride,
                                    int cout_valid, int64_t rows, int cout, int act,
                                    void *stream);

/* Weight and bias gradient of such a layer (csrc/wgrad.cu): dw[co][ci] (fp32, leading
 * dimension ldw) = sum_r dy[r][co] x[r][ci] for co < cout_valid, ci < cin_valid, and
 * db[co] = sum_r dy[r][co] (db may be NULL).  dy bf16 [rows][cout], x bf16 rows of
 * x_row_pitch elements; cout, cin multiples of 128.  Split-K tcgen05 GEMM with MN-major
 * operands: nsplit CTAs per 128 x 128 block, partial sums in `workspace`
 * (nsplit * cout * (cin + 1) floats), reduced in a fixed order (deterministic).  The
 * reference obtains these from cuDNN through autograd (sbmc/interfaces.py:78-106). */
SBMC_API int sbmc_wgrad_nhwc_bf16(const void *dy, const void *x, int64_t x_row_pitch,
                                  int64_t rows, int cout, int cin, int nsplit,
                                  float *workspace, float *dw, int64_t ldw, int cout_valid,
                                  int cin_valid, float *db, void *stream);

/* Weight gradient of the 3x3 convolutions (csrc/wgrad.cu): dw9[3 dy + dx][co][ci] (fp32) =
 * sum_{n,y,x} dp[n][y][x][co] x[n][y + dy - 1][x + dx - 1][ci] (zero outside the image);
 * and db[co] = sum dp[n][y][x][co] (db may be NULL); dp bf16 [n][h][w][cout], x bf16
 * [n][h][w][cin]; cout, cin multiples of 128; workspace: nsplit * cout * (9 * cin + 1) floats.
 * Deterministic split-K tcgen05 GEMM; the reference gets these gradients from cuDNN through
 * autograd (sbmc/modules.py:248-320). */
SBMC_API int sbmc_wgrad3x3_nhwc_bf16(const void *dp, const void *x, int64_t n, int h, int w,
                                     int cout, int cin, int nsplit, float *workspace,
                                     float *dw9, float *db, void *stream);

/* Weight normalization of all convolutions of a model in one launch (csrc/weight_bank.cu;
 * the reference wraps every convolution in nn.utils.weight_norm, sbmc/modules.py:84-87,
 * 176-179).  entries int64 [ne][16] = {v, g, F, D, 1/|v|, dW, dv, dg pointers, cout, cin, T,
 * cout_pad, cin_pad, 0, 0, 0}; blocks int64 [nblocks][2] = {entry, first output channel} (8
 * channels per block).  backward == 0: v, g -> bf16 operands F [T][cout_pad][cin_pad] and D
 * (data-gradient layout) + 1/|v|; backward != 0: dW fp32 [T][cout][cin] -> dv, dg. */
SBMC_API int sbmc_weight_bank_run(const int64_t *entries, const int64_t *blocks, int64_t nblocks,
                                  int backward, void *stream);

/* conv3x3 with the activation-derivative mask of sbmc_linear2_nhwc_bf16 in its epilogue
 * (mask bf16 [n][h][w][cout] or NULL): the data-gradient convolutions of the U-net. */
SBMC_API int sbmc_conv3x3_masked_nhwc_bf16(const void *x, const void *w9, const float *bias,
                                           const void *mask, int mask_act, void *y, int64_t n,
                                           int h, int w, int cin, int cout, int act,
                                           void *stream);

/* Memory-bound passes of the training pipeline on bf16 channels-innermost tensors
 * (csrc/train_ops.cu).  in [n_img][spp][hw][c] -> out [n_img][hw][c] = scale * sum over the
 * samples (bf16, or fp32 when out_f32 != 0): `features.mean(1)` (models.py:181). */
SBMC_API int sbmc_spp_reduce_nhwc_bf16(const void *in, void *out, int out_f32, int64_t n_img,
                                       int spp, int64_t hw, int c, float scale, void *stream);
/* out[b][s][p][:] = a[b][s][p][:] + scale * r[b][p][:] (a may be NULL or equal out). */
SBMC_API int sbmc_bcast_add_nhwc_bf16(const void *a, const void *r, void *out, int64_t n_img,
                                      int spp, int64_t hw, int c, float scale, void *stream);
/* Backward of MaxPool2d(2, 2) on x [n][h][w][c] (gradient to the first maximum of every
 * window, as torch) + the skip-connection gradient dskip (rows of skip_pitch elements, may
 * be NULL), times act'(x) (act as above, from the sign of x): modules.py:296-319. */
SBMC_API int sbmc_maxpool2x2_bwd_nhwc_bf16(const void *x, const void *dpool, const void *dskip,
                                           int64_t skip_pitch, void *out, int64_t n, int h,
                                           int w, int c, int act, void *stream);
/* Transpose of the decoder's bilinear upsampling (align_corners = False): dup rows of
 * `pitch` elements [n][h][w] -> out [n][hl][wl][c], times act'(coarse) when act != 0. */
SBMC_API int sbmc_upsample_bwd_nhwc_bf16(const void *dup, int64_t pitch, const void *coarse,
                                         void *out, int64_t n, int hl, int wl, int h, int w,
                                         int c, int act, void *stream);
/* out = g * act'(y), elementwise on bf16 (elems multiple of 8). */
SBMC_API int sbmc_dact_bf16(const void *y, const void *g, void *out, int64_t elems, int act,
                            void *stream);
/* out[c] = sum_r x[r][c] in fp32 (bias gradients); x bf16 rows of `pitch` elements;
 * workspace: nblk * c floats; deterministic. */
SBMC_API int sbmc_colsum_bf16(const void *x, int64_t pitch, int64_t rows, int c,
                              float *workspace, int nblk, float *out, void *stream);

/* 2 x 2 / stride 2 max pooling (the U-net's `downsample`, sbmc/modules.py:296-299:
 * nn.MaxPool2d(2, 2), floor mode) on bf16 channels-innermost x [n][h][w][c] ->
 * y [n][h/2][w/2][c]; c multiple of 8, 16-byte aligned pointers. */
SBMC_API int sbmc_maxpool2x2_nhwc_bf16(const void *x, void *y, int64_t n, int h, int w, int c,
                                       void *stream);

/* y = act(y + bias[channel]) in place on bf16 channels-innermost y [pixels][c]
 * (the bias + activation after every U-net convolution, modules.py:176-181);
 * act: 0 none, 1 ReLU, 2 LeakyReLU(0.01); c multiple of 8. */
SBMC_API int sbmc_bias_act_nhwc_bf16(void *y, const float *bias, int64_t pixels, int c, int act,
                            void *stream);

/* fp32 channel planes x [n][c][hw] -> bf16 channels-innermost y [n][hw][cpad],
 * zero-padded to cpad (multiple of 8) channels; `*_img_stride` in elements. */
SBMC_API int sbmc_nchw_to_nhwc_bf16(const float *x, int64_t x_img_stride, void *y,
                           int64_t y_img_stride, int64_t n, int c, int64_t hw, int cpad,
                           void *stream);

/* ---- sample-buffer reader (the callers' input format, sbmc/datasets.py) --- *
 * A .bin tile is a header followed by 1 + sample_count chunks, each an int32
 * byte count and an LZ4 frame (writer: pbrt_patches/sbmc_pbrt.diff:6140-6158).
 * The host parses the header and the chunk sizes (sbmc_b200/datasets.py), ships
 * the COMPRESSED bytes to the device, and these two entry points replace
 * `lz4.frame.decompress` (datasets.py:570-579) and the numpy assembly of
 * `TilesDataset._read_data` / `_preprocess_standard` /
 * `FullImagesDataset.__getitem__` (datasets.py:581-739, 744-778, 920-957). */

/* Inflates `nframes` LZ4 frames on the device, one warp per frame.
 * frame_table (device) is int64 [nframes][4] = {source byte offset in src,
 * source byte count, destination byte offset in dst, expected inflated size};
 * status (device, int32 [nframes]) receives 0 or an SBMC_LZ4_* code per frame
 * (the call itself only fails for bad arguments / launch errors).  Header,
 * block and content checksums (xxHash32) are verified when the frame has them. */
#define SBMC_LZ4_OK 0
#define SBMC_LZ4_BAD_MAGIC 1
#define SBMC_LZ4_BAD_HEADER 2
#define SBMC_LZ4_TRUNCATED 3
#define SBMC_LZ4_OVERFLOW 4
#define SBMC_LZ4_BAD_OFFSET 5
#define SBMC_LZ4_SIZE_MISMATCH 6
#define SBMC_LZ4_BLOCK_TOO_LARGE 7
#define SBMC_LZ4_BAD_CHECKSUM 8
SBMC_API int sbmc_lz4_frames_inflate(const void *src, const int64_t *frame_table,
                                     int64_t nframes, void *dst, int32_t *status,
                                     void *stream);

/* Builds the model's input tensors from inflated tiles.  raw (device) holds,
 * per tile, one image frame ([pixel_features][ts][ts] fp32: channel means then
 * variances) and spp sample frames `sample_stride_bytes` apart ([27 + 6*depth]
 * fp32 planes then [depth] int16 bounce-type planes, each [ts][ts]);
 * tile_table (device) is int64 [ntiles][4] = {image frame byte offset, first
 * sample frame byte offset, block_x, block_y}.  Outputs are planar fp32, the
 * tile pasted at rows block_y.., columns block_x.. of an h x w image:
 *   features [spp][nf][h][w]   (channel selection by SBMC_TILE_* flags, bounce
 *                               types expanded to 5 flag planes per vertex;
 *                               SBMC_TILE_LOG_RADIANCE applies the sbmc-mode
 *                               log(1 + max(.,0)) / 10 compression in place)
 *   radiance [spp][3][h][w]    diffuse + specular, raw
 *   low_spp  [3][h][w]         mean of radiance over the samples
 *   image_data / image_data_var [pixel_features/2][h][w], target_image [3][h][w]
 * Pixels no tile covers are left untouched (zero them first).  spp == 0 or
 * pixel_features == 0 skips the corresponding outputs (pointers may be NULL).
 * `row0` is the image row the outputs start at: they hold rows row0 .. row0+h-1
 * and tile rows outside that range are skipped, so a rank of a row-sharded job
 * assembles just its band (0 and h = image height for the whole image). */
#define SBMC_TILE_COORDS 1
#define SBMC_TILE_GBUFFER 2
#define SBMC_TILE_P 4
#define SBMC_TILE_LD 8
#define SBMC_TILE_BT 16
#define SBMC_TILE_LOG_RADIANCE 32
#define SBMC_TILE_ALIGNED 64   /* frame offsets % 16 == 0 and block_x % 4 == 0 */
SBMC_API int sbmc_tile_assemble_f32(const void *raw, const int64_t *tile_table, int64_t ntiles,
                                    int64_t sample_stride_bytes, int ts, int spp,
                                    int sample_features, int pixel_features, int path_depth,
                                    int flags, float *features, float *radiance,
                                    float *low_spp, float *image_data, float *image_data_var,
                                    float *target_image, int64_t h, int64_t w, int64_t row0,
                                    void *stream);

/* ---- fused optimizer step of the training interface ---------------------- *
 * sbmc/interfaces.py:78-106 clips the gradient norm at 1000
 * (torch.nn.utils.clip_grad_norm_) and steps Adam over every parameter tensor
 * of the model; these two entry points do both over ALL tensors in three
 * launches.  Device tables built by the caller: tensors int64 [ntensors][5] =
 * {param, grad, exp_avg, exp_avg_sq (fp32 device pointers), numel}; chunks int64
 * [nchunks][2] = {tensor index, first element}, one chunk per
 * SBMC_MT_CHUNK_ELEMS elements of a tensor. */
#define SBMC_MT_CHUNK_ELEMS 65536

/* norm_and_coef[0] = ||all gradients||_2, norm_and_coef[1] = min(1, max_norm /
 * (norm + 1e-6)) (device, 2 floats); partial: device scratch of nchunks floats.
 * Deterministic (no atomics). */
SBMC_API int sbmc_multi_tensor_grad_norm_f32(const int64_t *tensors, const int64_t *chunks,
                                             int64_t nchunks, float *partial, float max_norm,
                                             float *norm_and_coef, void *stream);

/* torch.optim.Adam's update (no weight decay, no amsgrad) on every element:
 *   g *= *clip_coef (written back when != 1; clip_coef may be NULL = no clipping)
 *   m += (g - m)(1 - beta1);  v = v beta2 + (1 - beta2) g g
 *   p -= lr / bias_correction1 * m / (sqrt(v) / bias_correction2_sqrt + eps)
 * with bias_correction1 = 1 - beta1^step, bias_correction2_sqrt = sqrt(1 - beta2^step).
 * Hyper-parameters are doubles: the fp32 constants of the update (1 - beta2, lr /
 * bias_correction1, ...) are derived in double like torch derives them. */
SBMC_API int sbmc_multi_tensor_adam_f32(const int64_t *tensors, const int64_t *chunks,
                                        int64_t nchunks, const float *clip_coef, double lr,
                                        double beta1, double beta2, double eps,
                                        double bias_correction1, double bias_correction2_sqrt,
                                        void *stream);

/* The same update for CUDA-graph capture: the step count is a device float (*step = steps
 * taken so far; incremented by one after the update), so a replayed graph advances the
 * bias corrections.  Used by FusedAdam(capturable=True). */
SBMC_API int sbmc_multi_tensor_adam_devstep_f32(const int64_t *tensors, const int64_t *chunks,
                                                int64_t nchunks, const float *clip_coef,
                                                double lr, double beta1, double beta2,
                                                double eps, float *step, void *stream);

/* ---- row-band entry points (H-sharding across GPUs, host streaming) ------ *
 * A band is `h` consecutive image rows.  weights / output / sum_w / d_output /
 * d_sum_w / d_weights cover exactly the band.  `data_ext` ([n][c][halo_top +
 * h + halo_bot][w]) carries the band plus real neighbour rows above / below
 * (from the adjacent band, or zeros at the image border); rows outside
 * data_ext read as zero.  `d_data_ext` has the same extended shape: the band's
 * samples scatter gradient into their neighbours' rows, which the caller adds
 * to the adjacent bands.  halo_* = 0 gives the plain ops.
 */
SBMC_API int sbmc_kernel_weighting_fwd_band_f32(const float *data_ext,
                                       const float *weights, float *output,
                                       float *sum_w, int64_t n, int c, int64_t h,
                                       int64_t w, int kh, int kw, int halo_top,
                                       int halo_bot, void *stream);

SBMC_API int sbmc_kernel_weighting_bwd_band_f32(const float *data_ext,
                                       const float *weights,
                                       const float *d_output,
                                       const float *d_sum_w, float *d_data_ext,
                                       float *d_weights, int64_t n, int c,
                                       int64_t h, int64_t w, int kh, int kw,
                                       int halo_top, int halo_bot, void *stream);

/* ---- host-pointer entry points (replace the *_cpu_float32 ops) ----------- *
 * Same semantics with HOST buffers (pageable or pinned): the library streams
 * row bands through the GPU `device` (H2D, kernel, D2H overlapped on several
 * streams) and returns when the host outputs are complete.  There is no CPU
 * compute path in this library.
 */
SBMC_API int sbmc_scatter2gather_host_f32(const float *scatter, float *gather, int64_t n,
                                 int kh, int kw, int64_t h, int64_t w,
                                 int device);

SBMC_API int sbmc_kernel_weighting_fwd_host_f32(const float *data, const float *weights,
                                       float *output, float *sum_w, int64_t n,
                                       int c, int64_t h, int64_t w, int kh,
                                       int kw, int device);

SBMC_API int sbmc_kernel_weighting_bwd_host_f32(const float *data, const float *weights,
                                       const float *sum_w, const float *d_output,
                                       const float *d_sum_w, float *d_data,
                                       float *d_weights, int64_t n, int c,
                                       int64_t h, int64_t w, int kh, int kw,
                                       int device);

/* Release the cached device staging buffers of the host entry points. */
SBMC_API int sbmc_b200_host_release(void);

#ifdef __cplusplus
}
#endif
#endif /* SBMC_B200_H_ */
