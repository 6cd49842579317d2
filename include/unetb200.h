/* libunetb200 -- C ABI of the B200-native U-Net hot path.
 *
 * The reference (usnistgov/semantic-segmentation-unet) has no FFI of its own: UNet/model.py calls TensorFlow/Keras
 * directly.  Each entry point below therefore replaces the *library kernel* TensorFlow would dispatch for one op of
 * the reference graph; the citation after each declaration is the reference call site (file:line under
 * /root/reference) whose arithmetic it implements.  A maintainer binds them with ctypes -- see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 (UB_OK) or a negative UB_ERR_*; ub_last_error() holds a thread-local message;
 *     nothing throws or aborts.
 *   - all pointers are caller-owned DEVICE memory (except where noted), valid until `stream` reaches the op.
 *     Launches are asynchronous on `stream`; no hidden synchronisation or allocation.
 *   - activations are NHWC.  dtype: UB_BF16 (product path) or UB_F32 (fp32 check mode).
 *   - conv weights (fp32 master and bf16 shadow) are packed [Cout][tap = 3*dy+dx][Cin]; deconv weights
 *     [(2*a+b)*Cout + co][Cin]; "dgrad packs" are the transposes produced by ub_transpose_pack.
 *   - "partial" buffers hold per-block partial sums: UB_STATS_ROWS rows; the entry point zero-fills them first.
 */
#ifndef UNETB200_H_
#define UNETB200_H_

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UB_VERSION 100
#define UB_OK 0
#define UB_ERR_INVALID_ARG (-1)
#define UB_ERR_UNSUPPORTED_SHAPE (-2)
#define UB_ERR_CUDA (-3)
#define UB_ERR_NCCL (-4)

#define UB_BF16 0
#define UB_F32 1

#define UB_STATS_ROWS 592   /* 148 SMs x 4: rows of every "partial" buffer */
#define UB_MAX_CLASSES 8    /* number_classes of the register-resident head kernels (weights of every class in registers) */
#define UB_MAX_CLASSES_ANY 255  /* number_classes of the head entry points: above UB_MAX_CLASSES the class-per-lane kernels of
                                   csrc/head_generic.cu run; 255 because labels and masks are uint8 (UNet/build_lmdb.py:151) */
#define UB_MAX_CHANNELS 16  /* number_channels of the first-layer kernels (1..4 templated, 5..16 with run-time channel loops) */
#define UB_ZSCORE_BLOCKS 256
#define UB_BORDER_CHUNKS 64  /* ub_border_sums: scratch = UB_BORDER_CHUNKS * 8 * C floats */

const char* ub_last_error(void);
int ub_version(void);
int ub_device_sm_count(void);
/* HOST utility (no device work): CRC-32C (Castagnoli) of n bytes at `data`, continuing from `crc` (0 to start), as used by
 * the TensorBundle checkpoint files `tf.train.Checkpoint.write` produces (UNet/train.py:96, :181-184). Returns the CRC. */
long long ub_host_crc32c(const void* data, long long n, long long crc);

/* ---- tensor-core implicit GEMMs (bf16 storage, fp32 accumulate in TMEM) -------------------------------------- */

/* Conv2D(3x3, same, relu) forward over concat(x0, x1) -- UNet/model.py:30-35 (+ :57 concat).  stats (nullable):
 * partial[UB_STATS_ROWS][2][Cout] sum / sum-of-squares of the activated output for the BatchNorm of model.py:36. */
int ub_conv3x3_fwd(const void* x0, int C0, const void* x1, int C1, const void* w, const float* bias, void* out,
                   float* stats, int N, int H, int W, int Cout, int relu, cudaStream_t stream);
/* The same forward (bias_cases != 0: `bias` is the [9][Cout] border-case table of ub_conv3x3_fwd_cases) followed IN THE SAME LAUNCH by
 * the finalisation of the BatchNormalization that follows the conv (UNet/model.py:36): the last CTA to finish reduces the partial rows
 * in a fixed order (fp64) and writes mean / rstd (biased variance, eps) and the momentum update of the moving statistics (unbiased
 * variance; nullable) -- no separate ub_bn_finalize launch.  counter: one zero-initialised device word, left zero. */
int ub_conv3x3_fwd_bn(const void* x0, int C0, const void* x1, int C1, const void* w, const float* bias, int bias_cases, void* out,
                      float* stats, int N, int H, int W, int Cout, int relu, float* mean, float* rstd, float* moving_mean,
                      float* moving_var, float momentum, float eps, unsigned int* counter, cudaStream_t stream);
/* Gradient w.r.t. the conv input (tape.gradient, UNet/model.py:219).  w_t = ub_transpose_pack(w, flip=1).
 * Channels [0,C0) go to dx0 and [C0,C0+C1) to dx1 (gradient of the concat; C1 == 0 or C1 == C0). */
int ub_conv3x3_dgrad(const void* dz, int Cout, const void* w_t, void* dx0, int C0, void* dx1, int C1, int N, int H,
                     int W, cudaStream_t stream);
/* Same, fused with the backward-BatchNorm reduction of the tensor whose gradient it writes (the LAST output: dx1 of a
 * concat dgrad, else dx0): a = that tensor's saved pre-normalisation activation, mean/rstd its batch statistics;
 * partial[UB_STATS_ROWS][2][C] = {sum dy, rstd * sum dy * (a - mean)} -- what ub_bn_bwd_reduce computes in a separate pass. */
int ub_conv3x3_dgrad_bnred(const void* dz, int Cout, const void* w_t, void* dx0, int C0, void* dx1, int C1, int N, int H, int W,
                           const void* a, const float* mean, const float* rstd, float* partial, cudaStream_t stream);
/* Gradient w.r.t. the conv kernel: dw fp32 [Cout][9][C0+C1] -- UNet/model.py:219. */
long long ub_conv3x3_wgrad_workspace_bytes(int C0, int C1, int Cout, int N, int H, int W);
int ub_conv3x3_wgrad(const void* x0, int C0, const void* x1, int C1, const void* dz, int Cout, float* dw, void* workspace,
                     long long workspace_bytes, int N, int H, int W, cudaStream_t stream);
/* Conv2DTranspose(2x2, stride 2, same, no activation) -- UNet/model.py:41-46.  x [N,h,w,Cin] -> out [N,2h,2w,Cout];
 * stats: partial[UB_STATS_ROWS][2][4*Cout] (finalise with groups = 4). */
int ub_deconv2x2_fwd(const void* x, int Cin, const void* w, const float* bias, void* out, float* stats, int N, int h,
                     int w_in, int Cout, cudaStream_t stream);
/* ... with the BatchNormalization of UNet/model.py:47 finalised by the last CTA (see ub_conv3x3_fwd_bn). */
int ub_deconv2x2_fwd_bn(const void* x, int Cin, const void* w, const float* bias, void* out, float* stats, int N, int h, int w_in,
                        int Cout, float* mean, float* rstd, float* moving_mean, float* moving_var, float momentum, float eps,
                        unsigned int* counter, cudaStream_t stream);
int ub_deconv2x2_dgrad(const void* dz, int Cout, const void* w_t, void* dx, int Cin, int N, int h, int w_in,
                       cudaStream_t stream);
/* ... fused with the backward reduction of the BatchNormalization in front of the up-sampling (the conv block whose output the
 * transposed convolution reads, UNet/model.py:97-131): partial[UB_STATS_ROWS][2][Cin] = {sum dy, rstd * sum dy * (a - mean)} over the
 * dx it writes, as ub_conv3x3_dgrad_bnred does -- replaces ub_bn_bwd_reduce for that layer. */
int ub_deconv2x2_dgrad_bnred(const void* dz, int Cout, const void* w_t, void* dx, int Cin, int N, int h, int w_in, const void* a,
                             const float* mean, const float* rstd, float* partial, cudaStream_t stream);
long long ub_deconv2x2_wgrad_workspace_bytes(int Cin, int Cout, int N, int h, int w_in);
int ub_deconv2x2_wgrad(const void* x, int Cin, const void* dz, int Cout, float* dw, void* workspace,
                       long long workspace_bytes, int N, int h, int w_in, cudaStream_t stream);

/* ---- first layer (Cin <= 4) and class head --------------------------------------------------------------------- */

/* Conv2D(3x3, same, relu) on the fp32 NCHW network input -- UNet/model.py:88.  w fp32 [64][9][Cin]. */
int ub_conv_first_fwd(const float* x_nchw, const float* w, const float* bias, void* out, float* partial, int N, int H, int W,
                      int Cin, int dtype, cudaStream_t stream);
/* partial scratch: UB_STATS_ROWS * Cin * 9 * 64 floats */
int ub_conv_first_wgrad(const float* x_nchw, const void* dz, float* dw, float* partial, int N, int H, int W, int Cin, int dtype,
                        cudaStream_t stream);
/* gradient w.r.t. the fp32 NCHW network input (tape.gradient(loss, img) of UNet.estimate_radius, UNet/model.py:186) */
int ub_conv_first_dgrad(const void* dz, const float* w, float* dx_nchw, int N, int H, int W, int Cin, int dtype, cudaStream_t stream);
/* 1x1 Conv2D(relu) 64 -> K classes -- UNet/model.py:136.  a_out fp32 [P][K]; partial [UB_STATS_ROWS][2][K]. */
int ub_head_fwd(const void* x, const float* w, const float* b, float* a_out, float* partial, long long P, int K, int dtype,
                cudaStream_t stream);
/* BatchNorm -> Softmax -> CategoricalCrossentropy + CategoricalAccuracy -- UNet/model.py:136-142, :211-215, :225-226.
 * labels: uint8 class index per pixel (nullable: inference); class_w nullable (all ones = reference behaviour);
 * label_smoothing eps (UNet/model.py:65, :77; the reference passes 0): target = onehot * (1 - eps) + eps / K;
 * dlogits = (softmax - target) * class_w[label] * inv_denom;
 * partial[UB_STATS_ROWS][2] = {inv_denom * sum CE, acc_scale * #correct} (sum the rows with ub_reduce_rows). */
int ub_head_loss(const float* a, const float* mean, const float* rstd, const float* gamma, const float* beta,
                 const unsigned char* labels, const float* class_w, float inv_denom, float acc_scale, float label_smoothing,
                 float* softmax_out, float* dlogits, float* partial, long long P, int K, cudaStream_t stream);
/* one-hot int32 [P][K] (the reference's label tensor, UNet/imagereader.py:302-312) -> uint8 index */
int ub_onehot_to_index(const int* onehot, unsigned char* idx, long long P, int K, cudaStream_t stream);
int ub_head_bwd_reduce(const float* dy, const float* a, const float* mean, const float* rstd, float* partial, long long P, int K,
                       cudaStream_t stream);
/* partial: UB_STATS_ROWS rows of {dW[K][64], db[K]} */
int ub_head_bwd_apply(const float* dy, const float* a, const void* x, const float* w, const float* mean, const float* rstd,
                      const float* gamma, const float* dbeta, const float* dgamma, void* dx, float* partial, long long P, int K,
                      int dtype, cudaStream_t stream);
/* same, and accumulates the BatchNorm-backward sums of the 64-channel tensor dx differentiates (saved activation red_a, batch
 * red_mean / red_rstd): red_partial[UB_STATS_ROWS][2][64], summed by ub_reduce_rows -- replaces ub_bn_bwd_reduce for that layer */
int ub_head_bwd_apply_bnred(const float* dy, const float* a, const void* x, const float* w, const float* mean, const float* rstd,
                            const float* gamma, const float* dbeta, const float* dgamma, void* dx, float* partial, long long P, int K,
                            int dtype, const void* red_a, const float* red_mean, const float* red_rstd, float* red_partial,
                            cudaStream_t stream);
/* Inference epilogue: argmax of the head written into each tile's zone of the uint8 mask -- UNet/inference.py:105-129.
 * x: [ntiles][h][w][64]; geo: device int[ntiles][6] = {cy0, cy1, cx0, cx1, dst_y, dst_x} (crop box inside the tile, destination
 * of its top-left corner in the mask); scale/shift: BatchNorm moving statistics of the head folded (gamma*rstd, beta-mean*scale);
 * softmax_out (nullable) [ntiles][h][w][K] serves the model-call contract of inference.py:105. */
int ub_head_argmax(const void* x, const float* w, const float* b, const float* scale, const float* shift, int K, int ntiles, int h,
                   int w_tile, const int* geo, unsigned char* mask, long long mask_ld, float* softmax_out, int dtype,
                   cudaStream_t stream);

/* ---- inference forms (training=False, UNet/model.py:240, inference.py:105): BatchNorm moving statistics folded into the
 * producer's epilogue, y = act(conv + b) * scale + shift; plain max-pool */
int ub_conv3x3_fwd_affine(const void* x0, int C0, const void* x1, int C1, const void* w, const float* bias, const float* scale,
                          const float* shift, void* out, int N, int H, int W, int Cout, int relu, cudaStream_t stream);
int ub_deconv2x2_fwd_affine(const void* x, int Cin, const void* w, const float* bias, const float* scale, const float* shift, void* out,
                            int N, int h, int w_in, int Cout, cudaStream_t stream);
int ub_conv_first_fwd_affine(const float* x_nchw, const float* w, const float* bias, const float* scale, const float* shift, void* out,
                             int N, int H, int W, int Cin, int dtype, cudaStream_t stream);
/* The same layer for N equal-sized inference tiles read IN PLACE from the resident normalised image (the tile loop of
 * UNet/inference.py:56-105 without one copy per tile): img fp32 [Cin][>= img_h rows][row_pitch], plane_stride floats between
 * channels; tile n = rows [origin_yx[2n], +H) x columns [origin_yx[2n+1], +W).  Source coordinates >= img_h / img_w (the
 * unpadded extent) are mirrored without repeating the edge -- np.pad(mode='reflect') to a multiple of 16, inference.py:46 --
 * and positions outside the tile are zero ('same' padding of the tile).  W % 4 == 0. */
int ub_conv_first_fwd_affine_tiles(const float* img, const int* origin_yx, int img_h, int img_w, long long row_pitch, long long plane_stride,
                                   const float* w, const float* bias, const float* scale, const float* shift, void* out, int N, int H, int W,
                                   int Cin, int dtype, cudaStream_t stream);
int ub_maxpool2x2_fwd(const void* y, void* pooled, int N, int H, int W, int C, int dtype, cudaStream_t stream);
/* scale = gamma / sqrt(moving_var + eps), shift = beta - moving_mean * scale */
int ub_bn_fold(const float* gamma, const float* beta, const float* moving_mean, const float* moving_var, float* scale, float* shift,
               int n, float eps, cudaStream_t stream);

/* ---- BatchNormalization(axis=1), eps 1e-3, momentum 0.99 -- UNet/model.py:36, :47 ------------------------------ */
int ub_bn_stats(const void* a, float* partial, long long M, int C, int dtype, cudaStream_t stream);
int ub_bn_finalize(const float* partial, int ncols, int groups, long long count, float* mean, float* rstd, float* moving_mean,
                   float* moving_var, float momentum, float eps, cudaStream_t stream);
int ub_bn_apply(const void* a, void* y, const float* mean, const float* rstd, const float* gamma, const float* beta,
                const unsigned char* drop_mask, long long M, int C, int dtype, cudaStream_t stream);
/* BN apply (+Dropout, model.py:62) fused with MaxPool2D(2) (model.py:52): writes y, pooled and the argmax slot. */
int ub_bn_apply_pool(const void* a, void* y, void* pooled, unsigned char* idx, const float* mean, const float* rstd,
                     const float* gamma, const float* beta, const unsigned char* drop_mask, int N, int H, int W, int C, int dtype,
                     cudaStream_t stream);
int ub_bn_bwd_reduce(const void* dy, const void* a, const float* mean, const float* rstd, float* partial, long long M, int C,
                     int dtype, cudaStream_t stream);
int ub_bn_bwd_apply(const void* dy, const void* a, const float* mean, const float* rstd, const float* gamma, const float* dbeta,
                    const float* dgamma, void* dz, float* partial, long long M, int C, int relu, int dtype, cudaStream_t stream);
/* out[c] = scale * sum_r partial[r * row_stride + c], c < ncols (fp64 accumulation, fixed order) */
int ub_reduce_rows(const float* partial, int rows, int row_stride, int ncols, float* out, float scale, cudaStream_t stream);
/* inference: rstd = 1/sqrt(moving_var + eps) (BN with training=False, UNet/model.py:240, inference.py:105) */
int ub_bn_inference_rstd(const float* moving_var, float* rstd, int n, float eps, cudaStream_t stream);

/* ---- pool / dropout backward ----------------------------------------------------------------------------------- */
/* dy = maxpool_bwd(dpool, idx) + dskip (skip fan-out, model.py:91/:132 ...), optional dropout backward */
int ub_maxpool2x2_bwd_add(const void* dpool, const unsigned char* idx, const void* dskip, const unsigned char* drop_mask, void* dy,
                          int N, int H, int W, int C, int dtype, cudaStream_t stream);
/* same, and accumulates the BatchNorm-backward sums of the tensor dy differentiates (saved activation a, batch mean / rstd):
 * red_partial[UB_STATS_ROWS][2][C] = per-block (sum dy, rstd * sum dy (a - mean)), to be summed by ub_reduce_rows -- replaces
 * the ub_bn_bwd_reduce pass over dy and a */
int ub_maxpool2x2_bwd_add_bnred(const void* dpool, const unsigned char* idx, const void* dskip, const unsigned char* drop_mask, void* dy,
                                int N, int H, int W, int C, const void* a, const float* mean, const float* rstd, float* red_partial,
                                int dtype, cudaStream_t stream);
int ub_dropout_bwd(const void* in, const unsigned char* mask, void* out, long long n, int dtype, cudaStream_t stream);
/* Philox4x32-10 Bernoulli(0.5) keep mask, one byte per element */
int ub_dropout_mask(unsigned char* mask, long long n, unsigned long long seed, unsigned long long offset, cudaStream_t stream);

/* ---- optimizer / packing / input ------------------------------------------------------------------------------- */
/* Keras Adam (UNet/model.py:79, :223): theta -= lr_t * m / (sqrt(v) + eps); optional bf16 shadow copy of theta */
int ub_adam(float* param, const float* grad, float* m, float* v, void* bf16_shadow, long long n, float lr_t, float beta1,
            float beta2, float eps, float grad_scale, cudaStream_t stream);
/* same update with lr_t read from device memory at execution time (the step can then be replayed from a CUDA graph) */
int ub_adam_dev(float* param, const float* grad, float* m, float* v, void* bf16_shadow, long long n, const float* lr_t_dev, float beta1,
                float beta2, float eps, float grad_scale, cudaStream_t stream);
/* dst[c][t'][r] = src(r,t,c) (t' = T-1-t when flip): builds the dgrad weight packs.
 * src_layout 0: src [R][T][C] (conv [Cout][tap][Cin]); 1: src [T][R][C] (deconv [(2a+b)][Cout][Cin]) */
int ub_transpose_pack(const float* src, void* dst, int R, int T, int C, int flip, int src_layout, int dst_dtype,
                      cudaStream_t stream);
/* every pack of the model in one launch.  jobs_dev: device array of njobs records
 *   { const float* src; void* dst; int R, T, C, flip, src_layout, tiles_c, tiles_r, tile_begin; }   (48 bytes, natural alignment)
 * with tiles_c = ceil(C/32), tiles_r = ceil(R/32), tile_begin = exclusive prefix sum of tiles_c*tiles_r*T; total_tiles = the sum */
int ub_transpose_pack_multi(const void* jobs_dev, int njobs, int total_tiles, int dst_dtype, cudaStream_t stream);
int ub_cast_bf16(const float* src, void* dst, long long n, cudaStream_t stream);
/* zscore_normalize (UNet/imagereader.py:33-66) per plane; src_dtype 0 = u8, 1 = u16, 2 = f32;
 * scratch: planes * UB_ZSCORE_BLOCKS * 2 doubles */
int ub_zscore(const void* src, int src_dtype, float* dst, double* scratch, int planes, long long plane, cudaStream_t stream);
/* the same normalisation in two halves, for an image whose rows are split across ranks (unetb200.inference.segment_banded; first
 * GPU run pending): ub_zscore_sums -> sums = double[planes][2] (sum, sum of squares) over `plane` elements of each plane
 * (planes `plane_stride` elements apart; scratch as for ub_zscore), to be SUM-all-reduced; ub_zscore_apply_sums normalises
 * contiguous planes with the statistics sums / count */
int ub_zscore_sums(const void* src, int src_dtype, double* sums, double* scratch, int planes, long long plane, long long plane_stride,
                   cudaStream_t stream);
int ub_zscore_apply_sums(const void* src, int src_dtype, float* dst, const double* sums, double count, int planes, long long plane,
                         cudaStream_t stream);

/* ---- training-time augmentation on raw-pixel batches (UNet/augment.py:19-174, called from UNet/imagereader.py:283-294) ---
 * Planes are NCHW as the reader ships them. dtype codes here: 0 = u8, 1 = u16, 2 = f32.
 * ub_aug_warp: dst[n,c,y,x] = bilinear sample of src[n,c] at (m0 x + m1 y + m2, m3 x + m4 y + m5), mats = double[N][6], the
 *   INVERSE map skimage.transform.warp(order=1, mode='reflect') is given (augment.py:160-167); mirror boundary without edge
 *   repeat; fp64 coordinates.  dst_dtype 2 = f32; 0 = uint8 class index rounded half-to-even (the mask, augment.py:155).
 * ub_aug_minmax: partial = float[N][64][2] per-image (min, max) partials over all channels.
 * ub_aug_noise: x += range_n * (factors[n][0] * z + factors[n][1]), z ~ N(0,1) Philox(seed, offset), range_n = max - min
 *   from the partials (noise: augment.py:118-127; intensity shift: :141-153).
 * ub_aug_blur_axis: one axis (0 = H, 1 = W) of scipy.ndimage.gaussian_filter(mode='reflect') (augment.py:130-139): per-image
 *   symmetric taps weights = double[N][33] (w[0] = centre), radius = int[N] (0 = copy).
 * ub_aug_chanmix: the same filter along the channel axis as an N x C x C matrix (double[N][C][C]), in place. */
int ub_aug_warp(const void* src, int src_dtype, void* dst, int dst_dtype, const double* mats, int N, int C, int H, int W,
                cudaStream_t stream);
int ub_aug_minmax(const float* x, float* partial, int N, long long per_image, cudaStream_t stream);
int ub_aug_noise(float* x, const float* minmax_partial, const float* factors, int N, long long per_image, unsigned long long seed,
                 unsigned long long offset, cudaStream_t stream);
int ub_aug_blur_axis(const float* src, float* dst, const double* weights, const int* radius, int axis, int N, int C, int H, int W,
                     cudaStream_t stream);
int ub_aug_chanmix(float* x, const double* mix, int N, int C, long long plane, cudaStream_t stream);

/* ---- BatchNorm folded into the consumer convolution (training forward; NOT YET ENABLED in the step: compiled, parity cases written,
 * first GPU run pending) ------------------------------------------------------------------------------------------------------
 * The producer's BatchNorm (UNet/model.py:36) is y = s a + t per channel; the consumer conv reads the pre-BatchNorm activation a
 * with W' = W s[ci] and a bias that depends on the pixel's border case, so y is never written.
 * ub_fold_conv3_weights: w fp32 [Cout][9][C0+C1] -> w_out bf16 (same layout) scaled per input channel; a source with mean == NULL
 *   is taken as already normalised (s = 1, t = 0); bias9 = float[9][Cout] (case = row case * 3 + column case; 0 first, 1 interior,
 *   2 last); scale_out / shift_out = float[C0+C1] (s, t) for ub_wgrad_fold_fix.
 * ub_conv3x3_fwd_cases: ub_conv3x3_fwd with that bias table (H, W >= 2).
 * ub_border_sums: sdz = float[9][C], sum of dz over the output pixels whose tap neighbour is inside the image; total = float[C]
 *   sum of dz over all pixels (the bias gradient); scratch = UB_BORDER_CHUNKS * 8 * C floats.
 * ub_wgrad_fold_fix: dw [Cout][9][Cin] (weight gradient computed with x = a) <- scale[ci] * dw + shift[ci] * sdz[tap][co]. */
int ub_fold_conv3_weights(const float* w, int Cout, int C0, const float* mean0, const float* rstd0, const float* gamma0, const float* beta0,
                          int C1, const float* mean1, const float* rstd1, const float* gamma1, const float* beta1, const float* bias,
                          void* w_out, float* bias9, float* scale_out, float* shift_out, cudaStream_t stream);
int ub_conv3x3_fwd_cases(const void* x0, int C0, const void* x1, int C1, const void* w, const float* bias9, void* out, float* stats, int N,
                         int H, int W, int Cout, int relu, cudaStream_t stream);
int ub_border_sums(const void* dz, const float* total, float* sdz, float* scratch, int N, int H, int W, int C, int dtype, cudaStream_t stream);
int ub_wgrad_fold_fix(float* dw, const float* scale, const float* shift, const float* sdz, int Cout, int Cin, cudaStream_t stream);
/* The backward sums of the BatchNormalization that FEEDS a convolution (channels [c_begin, c_begin + c_count) of its Cin inputs), without
 * a pass over that BatchNorm's gradient tensor: with dy = dgrad(dz, w),
 *   dbeta[c] = sum dy = sum_{co,tap} w * sdz[tap][co];  dgamma[c] = sum dy * xhat = rstd[c] * sum_{co,tap} w * (dw_a - mean[c] * sdz[tap][co])
 * w [Cout][taps][Cin] = the weights the dgrad used (bf16 shadow, or fp32 for the 1x1 head), dw_a [Cout][taps][Cin] = the weight gradient
 * computed on the pre-BatchNorm activation (BEFORE ub_wgrad_fold_fix), sdz [taps][Cout] = ub_border_sums (taps == 1: the bias gradient).
 * Replaces ub_bn_bwd_reduce / the fused dgrad reduction for every BatchNorm whose only consumer is a folded convolution. */
int ub_bn_bwd_sums_wgrad(const void* w, int w_dtype, const float* dw_a, const float* sdz, int Cout, int taps, int Cin, int c_begin, int c_count,
                         const float* mean, const float* rstd, float* dbeta, float* dgamma, cudaStream_t stream);
/* the 1x1 head on a folded input (no padding, one bias): w fp32 [K][64] -> w_out = w s[c], bias_out[k] = b[k] + sum_c w[k][c] t[c];
 * its weight gradient computed with x = a is fixed by dW[k][c] = s[c] dW[k][c] + t[c] db[k] */
int ub_fold_head_weights(const float* w, const float* bias, const float* mean, const float* rstd, const float* gamma, const float* beta,
                         float* w_out, float* bias_out, float* scale_out, float* shift_out, int K, cudaStream_t stream);
int ub_head_wgrad_fold_fix(float* dw, const float* db, const float* scale, const float* shift, int K, cudaStream_t stream);
/* ub_bn_apply_pool without the y output: pooled = maxpool2x2(dropout(BN(a))) and the argmax slots only */
int ub_bn_pool(const void* a, void* pooled, unsigned char* idx, const float* mean, const float* rstd, const float* gamma,
               const float* beta, const unsigned char* drop_mask, int N, int H, int W, int C, int dtype, cudaStream_t stream);

/* ---- fp32 check mode (CUDA cores, fp32 storage) ----------------------------------------------------------------- */
int ub_check_conv3x3(const float* x0, int C0, const float* x1, int C1, const float* w, const float* bias, float* out0, int Co0,
                     float* out1, int Co1, int N, int H, int W, int relu, cudaStream_t stream);
int ub_check_conv3x3_wgrad(const float* x0, int C0, const float* x1, int C1, const float* dz, int Cout, float* dw, int N, int H,
                           int W, cudaStream_t stream);
int ub_check_deconv2x2_fwd(const float* x, const float* w, const float* bias, float* out, int N, int h, int w_in, int Cin, int Cout,
                           cudaStream_t stream);
int ub_check_deconv2x2_dgrad(const float* dz, const float* w, float* dx, int N, int h, int w_in, int Cin, int Cout,
                             cudaStream_t stream);
int ub_check_deconv2x2_wgrad(const float* x, const float* dz, float* dw, int N, int h, int w_in, int Cin, int Cout,
                             cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UNETB200_H_ */
