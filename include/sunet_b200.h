/* sunet_b200.h — C ABI of libsunet_b200.so
 *
 * The drop-in boundary of the B200-native SelectiveUNet hot path.  The reference
 * (yellofi/SelectiveNet_for_semantic_segmentation_binary) has no FFI layer of its own: every
 * op below replaces a stock-PyTorch dispatch made from the reference's Python, cited per entry
 * point as file:line into /root/reference.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error (sunet_last_error() explains);
 *     nothing throws, exits or allocates device memory: the caller owns every buffer,
 *     including workspaces;
 *   - all work is enqueued on the stream passed in (a cudaStream_t / CUstream); no implicit
 *     synchronisation, so every call is CUDA-graph capturable;
 *   - activations are NHWC bf16 ("pix_stride" = elements between consecutive pixels, which
 *     lets a tensor be a channel slice of a wider buffer); parameters, statistics, logits and
 *     gradients of parameters are fp32 in the reference's own layouts;
 *   - functions are re-entrant and may be called from any host thread (PyTorch's autograd
 *     worker threads included); one process drives one GPU.
 */
#ifndef SUNET_B200_H_
#define SUNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* sunet_stream_t;

#define SUNET_ABI_VERSION 2
/* workspace every reduce-type call may use (bytes); callers pass one buffer of at least this size */
#define SUNET_WORKSPACE_BYTES (8u << 20)

int sunet_abi_version(void);
const char* sunet_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long sunet_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * G1: tensor-core implicit GEMM  D[pixel, n] = sum_{tap,c} A_tap[pixel, c] * W[n, tap*C + c] (+bias)
 * tcgen05.mma / TMEM / TMA.  Replaces nn.Conv2d(k3,p1) forward + backward-data and
 * nn.ConvTranspose2d(k2,s2) forward + backward-data (model.py:11, :44-45, :51-52, :57-58;
 * backward via train.py:208).
 * ---------------------------------------------------------------------------------------- */
enum { SUNET_A_CONV3X3 = 0, SUNET_A_PLAIN = 1, SUNET_A_GATHER2X2 = 2 };
enum { SUNET_D_NHWC = 0, SUNET_D_SCATTER2X2 = 1 };

typedef struct sunet_conv_gemm_args {
  int batch, height, width;    /* the pixel grid that is the GEMM M dimension                    */
  int a_mode;                  /* CONV3X3: 9 taps, zero pad 1; PLAIN: 1 tap;                     */
                               /* GATHER2X2: src0 is [batch][2*height][2*width][..], 4 taps (a,b) */
  const void* src0;            /* NHWC bf16                                                       */
  int src0_channels, src0_pix_stride;
  const void* src1;            /* optional second source, concatenated after src0 along C (NULL)  */
  int src1_channels, src1_pix_stride;
  const void* weights;         /* bf16 [n_total][k_total], k = tap*(C0+C1) + c                    */
  int n_total, k_total;
  const float* bias;           /* optional fp32 [n_total]                                         */
  void* dst;                   /* NHWC bf16; SCATTER2X2: [batch][2*height][2*width][n_total/4],   */
  int dst_pix_stride;          /*   column n = (a*2+b)*C' + co goes to pixel (2y+a, 2x+b), chan co */
  int d_mode;
  float* stats;                /* optional fp32 [sunet_conv_gemm_stat_rows(args)][n_total][2]:    */
                               /*   per-CTA partial (sum, sum of squares) of the bf16 outputs     */
  /* Optional fused BatchNorm-backward reduction (CBR_2D backward, model.py:9-15 via train.py:208): when this
   * launch is the backward-data conv whose output dA is the gradient w.r.t. relu(bn(y)), pass y (NHWC bf16,
   * n_total channels) and that BN's per-channel scale/shift/mean/invstd.  `stats` then receives per-CTA partial
   * (sum g, sum g*xhat) with g = dA * (scale*y+shift > 0), xhat = (y-mean)*invstd — exactly the rows
   * sunet_bn_bwd_apply() folds.  dst is stored unmasked.  Only where sunet_conv_gemm_bnb_supported() != 0. */
  const void* bnb_y;
  int bnb_y_pix_stride;
  const float* bnb_scale;
  const float* bnb_shift;
  const float* bnb_mean;
  const float* bnb_invstd;
  /* Optional inference epilogue: dst = relu(acc * ep_scale[n] + ep_shift[n]) (BatchNorm in eval mode folded into
   * the conv: model.py:11-13 under net.eval(), eval.py:191-206).  fp32 [n_total]; no statistics, no bias. */
  const float* ep_scale;
  const float* ep_shift;
  int bnb_col0;                /* the reduction covers output columns [bnb_col0, n_total) (multiple of 64; y and the  */
                               /* bnb_* vectors then have n_total - bnb_col0 channels); lower columns keep (sum, sq)   */
  /* Optional TRAINING prologue (CBR_2D forward, model.py:11-13 with BatchNorm in training mode): the source(s) flagged
   * in pro_mask (bit 0: src0, bit 1: src1) hold the RAW conv output y of the producer block; every landed operand tile
   * is transformed in place to relu(y * pro_scale[c] + pro_shift[c]) rounded to bf16 — bit-identical to what
   * sunet_bn_relu_pool would have materialised — so that pass and the activation tensor disappear.  fp32 vectors over
   * the concatenated input channels (src0 first).  Only where sunet_conv_gemm_pro_supported() != 0. */
  const float* pro_scale;
  const float* pro_shift;
  int pro_mask;
} sunet_conv_gemm_args;

int sunet_conv_gemm(const sunet_conv_gemm_args* args, sunet_stream_t stream);
/* rows of `stats` the call described by args will write (args->stats itself is ignored) */
int sunet_conv_gemm_stat_rows(const sunet_conv_gemm_args* args);
/* 1 if the kernel variant this shape dispatches to implements the bnb_* epilogue, else 0 */
int sunet_conv_gemm_bnb_supported(const sunet_conv_gemm_args* args);
/* 1 if the kernel variant this shape dispatches to implements the pro_* prologue, else 0 */
int sunet_conv_gemm_pro_supported(const sunet_conv_gemm_args* args);

/* ------------------------------------------------------------------------------------------
 * G2: weight-gradient GEMM over pixels, split-K with fp32 partials.
 *   P[split][tap][m][n] = sum_p A[p, m] * B[p (+) tap, n]
 * Replaces the backward-weight half of conv2d / conv_transpose2d (train.py:208).
 * ---------------------------------------------------------------------------------------- */
typedef struct sunet_wgrad_gemm_args {
  int batch, height, width;    /* pixel grid of A (the reduction dimension)                       */
  const void* a;               /* NHWC bf16, never shifted                                         */
  int a_channels, a_pix_stride;
  int b_mode;                  /* SUNET_A_CONV3X3 (9 shifted taps), _PLAIN (1), _GATHER2X2 (4)     */
  const void* b0;
  int b0_channels, b0_pix_stride;
  const void* b1;              /* optional second B source (concat), NULL otherwise                */
  int b1_channels, b1_pix_stride;
  float* partials;             /* fp32 [splits][taps][a_channels][b0+b1 channels]                  */
  size_t partials_bytes;
  /* Optional training prologue on the B operand (the layer input): b0 holds the raw conv output y of the producer
   * block and every landed tile is transformed to relu(y * b_pro_scale[c] + b_pro_shift[c]) (bf16) in place, as in
   * sunet_conv_gemm's pro_*.  Only where sunet_wgrad_gemm_pro_supported() != 0 (conv3x3, one B source, >= 256 A
   * channels: the CTA-pair kernel with the shifted-window B box). */
  const float* b_pro_scale;
  const float* b_pro_shift;
} sunet_wgrad_gemm_args;

int sunet_wgrad_gemm(const sunet_wgrad_gemm_args* args, sunet_stream_t stream);
int sunet_wgrad_gemm_splits(const sunet_wgrad_gemm_args* args); /* how many splits it will write */
int sunet_wgrad_gemm_pro_supported(const sunet_wgrad_gemm_args* args);

/* sum the split-K partials into the reference's parameter-gradient layout.
 *   layout 0: conv3x3  grad[co][ci][3][3]   from P[s][r*3+q][co][ci]
 *   layout 1: convT    grad[ci][co][2][2]   from P[s][a*2+b][ci][co]
 *   layout 2: first conv (im2col'ed input)  grad[co][cin][3][3] from P[s][0][co][(r*3+q)*cin+ci]
 *   layout 3: first conv, paired-pixel form (a_channels 128, b_channels 64): the two diagonal 64x32 blocks of
 *             P[s][0][128][64] summed (see sunet_pack_input_im2col32) */
int sunet_wgrad_reduce(const float* partials, int splits, int taps, int a_channels, int b_channels, int layout,
                       int real_cin, float* grad, sunet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Layout / packing (fp32 reference layouts -> bf16 kernel layouts)
 * ---------------------------------------------------------------------------------------- */
/* x: fp32 NCHW [B][cin][H][W] (cin*9 <= 64) -> bf16 [B][H][W][64], channel tap*cin+ci = x[.., y+r-1, x+q-1] */
int sunet_pack_input_im2col(const float* x, void* out, int batch, int cin, int height, int width,
                            sunet_stream_t stream);
/* Paired-pixel form of the first layer (cin = 2 or 3, even width): bf16 [B][H][W][32], channel tap*cin+ci, rest 0.
 * Viewed as [B][H][W/2][64], a row holds two adjacent pixels; with the weights of
 * sunet_pack_conv1_pair_weights ([128][64]: rows 0..63 = [w | 0], rows 64..127 = [0 | w]) one PLAIN
 * sunet_conv_gemm over the (B, H, W/2) grid with n_total = 128 writes the NHWC [B][H][W][64] output viewed as
 * [B][H][W/2][128] — half the im2col traffic of the 64-wide form and no new GEMM kernel.  The matching weight
 * gradient is a PLAIN sunet_wgrad_gemm with A = dY viewed [..][W/2][128], B = this tensor, folded by
 * sunet_wgrad_reduce(layout 3). */
int sunet_pack_input_im2col32(const float* x, void* out, int batch, int cin, int height, int width,
                              sunet_stream_t stream);
int sunet_pack_conv1_pair_weights(const float* w, void* wf, int cout, int cin, sunet_stream_t stream);
/* Device-side input pipeline (utils/data_utils.py:94-126 Normalization + RandomFlip, :159-168 ToTensor, :216-219 the
 * /255 of PatchDataset): img uint8 [B][H][W][3] as decoded, lut[256] = float32((b/255) - mean)/std computed by the
 * caller in numpy (bit-identical values), flip[B] (bit 0 left-right, bit 1 up-down; NULL = none) -> the same
 * [B][H][W][32] bf16 tensor sunet_pack_input_im2col32 would produce from the transformed float32 NCHW tensor.
 * sunet_pack_label_u8: label uint8 [B][H][W] -> float32 {0,1} = (label/255).astype(uint8), same flips. */
int sunet_pack_input_u8_im2col32(const void* img, const float* lut, const void* flip, void* out, int batch, int height,
                                 int width, sunet_stream_t stream);
int sunet_pack_label_u8(const void* label, const void* flip, float* out, int batch, int height, int width,
                        sunet_stream_t stream);
/* Conv2d weight [co][ci][3][3] -> wf [co][9*ci] (k = tap*ci_total + ci) and, if wd != NULL,
 * the dgrad operand wd [ci][9*co] (k = flipped_tap*co_total + co) */
int sunet_pack_conv3x3_weights(const float* w, void* wf, void* wd, int cout, int cin, sunet_stream_t stream);
/* first conv: [co][cin][3][3] -> [co][64] with k = tap*cin + ci, zero padded */
int sunet_pack_conv1_weights(const float* w, void* wf, int cout, int cin, sunet_stream_t stream);
/* ConvTranspose2d weight [ci][co][2][2] -> wf [4*co][ci] (row (a*2+b)*co_total+co), wd [ci][4*co];
 * bias [co] -> bias4 [4*co] */
int sunet_pack_convT_weights(const float* w, const float* bias, void* wf, void* wd, float* bias4, int cin, int cout,
                             sunet_stream_t stream);

/* every pack of one forward pass in ONE launch: a DEVICE table of jobs.
 *   kind 0: conv3x3  (a = cout, b = cin; wf, wd as sunet_pack_conv3x3_weights)
 *   kind 1: first conv (a = cout, b = cin; wf as sunet_pack_conv1_weights)
 *   kind 3: first conv, paired-pixel form (a = 64, b = cin; wf as sunet_pack_conv1_pair_weights)
 *   kind 2: ConvTranspose2d (a = cin, b = cout; wf, wd, bias, bias4 as sunet_pack_convT_weights) */
typedef struct sunet_pack_job {
  int kind, a, b, tile_start;  /* first block of this job in the flat grid: prefix sum of (a/32)*(b/32) for kinds 0
                                  and 2, 1 for kinds 1 and 3; total_tiles = the sum over all jobs */
  const float* w;
  const float* bias;
  void* wf;
  void* wd;
  float* bias4;
} sunet_pack_job;
int sunet_pack_weights_table(const sunet_pack_job* jobs_dev, int n_jobs, int total_tiles, sunet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * BatchNorm2d + ReLU (+ MaxPool2d(2)), model.py:12-13,31,35,39
 * ---------------------------------------------------------------------------------------- */
/* training: fold per-CTA partials into batch statistics, produce the per-channel affine
 * (scale, shift) applied to the bias-free conv output, and update the running statistics
 * (momentum 0.1, unbiased variance; the conv bias only shifts running_mean). */
int sunet_bn_finalize(const float* stats, int rows, int channels, long long count, const float* gamma,
                      const float* beta, const float* conv_bias, float* running_mean, float* running_var,
                      long long* num_batches_tracked, float momentum, float eps, float* scale, float* shift,
                      float* mean, float* invstd, sunet_stream_t stream);
/* eval: affine from running statistics */
int sunet_bn_eval_affine(const float* gamma, const float* beta, const float* conv_bias, const float* running_mean,
                         const float* running_var, float eps, float* scale, float* shift, int channels,
                         sunet_stream_t stream);
/* a = relu(y*scale + shift) -> bf16; pooled (optional) = 2x2/stride-2 max of a */
int sunet_bn_relu_pool(const void* y, int y_pix_stride, const float* scale, const float* shift, void* a,
                       int a_pix_stride, void* pooled, int pooled_pix_stride, int batch, int height, int width,
                       int channels, sunet_stream_t stream);
/* sunet_bn_relu_pool that also stores ywin[b][y/2][x/2][c] = the conv output y of the window's FIRST maximum of
 * the fp32 activation (the element MaxPool2d's backward routes the gradient to; model.py:31).  The backward-data
 * conv that produces the pooled gradient can then take ywin as its bnb_y and reduce the pool-routed part of this
 * block's BatchNorm backward in its epilogue. */
int sunet_bn_relu_pool_ywin(const void* y, int y_pix_stride, const float* scale, const float* shift, void* a,
                            int a_pix_stride, void* pooled, int pooled_pix_stride, void* ywin, int ywin_pix_stride,
                            int batch, int height, int width, int channels, sunet_stream_t stream);
/* pooled = 2x2 / stride-2 max of a (model.py:31,35,39 when the activation was produced by the ep_* epilogue) */
int sunet_maxpool2x2(const void* a, int a_pix_stride, void* pooled, int pooled_pix_stride, int batch, int height,
                     int width, int channels, sunet_stream_t stream);
/* backward of [BN(train) -> ReLU -> (skip + MaxPool)]:
 *   g = (dA [+ dPool routed to the first maximum of each 2x2 window]) * (a > 0)
 *   dgamma = sum g*xhat, dbeta = sum g, dy = scale*(g - dbeta/n - xhat*dgamma/n)  -> bf16
 * dA may be NULL when only the pooled branch carries gradient. */
int sunet_bn_relu_pool_bwd(const void* dA, int dA_pix_stride, const void* dPool, int dPool_pix_stride, const void* y,
                           int y_pix_stride, const float* scale, const float* shift, const float* mean,
                           const float* invstd, const float* gamma, float* dgamma, float* dbeta, void* dy,
                           int dy_pix_stride, int batch, int height, int width, int channels, void* workspace,
                           size_t workspace_bytes, sunet_stream_t stream);
/* Second half of sunet_bn_relu_pool_bwd for a non-pooled block whose reduction was fused into the producer of
 * dA (sunet_conv_gemm with bnb_y, or sunet_heads_bwd with bnb_y): folds `partial_rows` rows of
 * [channels][2] = (sum g, sum g*xhat) into dgamma / dbeta and writes dy.  dA may be stored unmasked. */
int sunet_bn_bwd_apply(const void* dA, int dA_pix_stride, const void* y, int y_pix_stride, const float* scale,
                       const float* shift, const float* mean, const float* invstd, const float* partials,
                       int partial_rows, float* dgamma, float* dbeta, void* dy, int dy_pix_stride, int batch,
                       int height, int width, int channels, void* workspace, size_t workspace_bytes,
                       sunet_stream_t stream);
/* Pooled block whose reduction rows were written by the two producers of its gradient (skip part: the decoder
 * backward-data conv with bnb_col0 = C over [d_up | d_skip]; pooled part: the encoder backward-data conv with
 * bnb_y = ywin).  Row r, channel c of source i sits at partials_i[(r*stride_i + col_i + c)*2 + {0,1}]. */
int sunet_bn_pool_bwd_apply(const void* dA, int dA_pix_stride, const void* dPool, int dPool_pix_stride, const void* y,
                            int y_pix_stride, const float* scale, const float* shift, const float* mean,
                            const float* invstd, const float* partials0, int rows0, int stride0, int col0,
                            const float* partials1, int rows1, int stride1, int col1, float* dgamma, float* dbeta,
                            void* dy, int dy_pix_stride, int batch, int height, int width, int channels,
                            void* workspace, size_t workspace_bytes, sunet_stream_t stream);
/* out[c] = sum_rows stats[row][col_offset + c][0]  (column sums from the G1 epilogue; ConvT bias grad) */
int sunet_colsum_finalize(const float* stats, int rows, int n_total, int col_offset, int channels, float* out,
                          sunet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Heads: three 1x1 convs 64 -> 1 (model.py:62,65-66,96,99-101)
 * ---------------------------------------------------------------------------------------- */
/* logits[h][p] = b_h + sum_c a[p][c] * w_h[c];  nheads = 1 or 3; planar fp32 == the [N,H,W] tensors */
int sunet_heads_fwd(const void* a, int a_pix_stride, const float* w0, const float* b0, const float* w1,
                    const float* b1, const float* w2, const float* b2, int nheads, float* logits, long long pixels,
                    sunet_stream_t stream);
/* last block fused: a = relu(y*scale + shift) (64 channels) and the heads' logits from the same registers
 * (model.py:12-13 of decoder_layer_1_1 + :96-101).  a may be NULL when the backward pass recomputes the
 * activation from y (sunet_heads_bwd_bn): the tensor is then never written. */
int sunet_bn_relu_heads(const void* y, int y_pix_stride, const float* scale, const float* shift, void* a,
                        int a_pix_stride, const float* w0, const float* b0, const float* w1, const float* b1,
                        const float* w2, const float* b2, int nheads, float* logits, long long pixels,
                        sunet_stream_t stream);
/* sunet_heads_bwd with the activation recomputed from y = the last block's conv output (nothing but y is read),
 * plus that block's BatchNorm-backward reduction: bn_partials[sunet_heads_bwd_bn_rows(pixels)][64][2] receives
 * per-block (sum g, sum g*xhat), g = dA where relu(bn(y)) > 0 — the rows sunet_bn_bwd_apply() folds.
 * addend (optional, NHWC bf16, may alias dA): gradient of the same activation from heads handled by an earlier call
 * (model.py:106-191 `UNet` has 2 x 3 head channels); it is added before rounding and enters the reduction rows. */
int sunet_heads_bwd_bn_rows(long long pixels);
int sunet_heads_bwd_bn(const float* dlogits, const void* y, int y_pix_stride, const float* scale, const float* shift,
                       const float* mean, const float* invstd, const float* w0, const float* w1, const float* w2,
                       int nheads, void* dA, int dA_pix_stride, float* dw0, float* db0, float* dw1, float* db1,
                       float* dw2, float* db2, float* bn_partials, const void* addend, int addend_pix_stride,
                       long long pixels, void* workspace, size_t workspace_bytes, sunet_stream_t stream);
/* dA[p][c] = sum_h dl[h][p]*w_h[c] (bf16);  dw_h[c] = sum_p dl[h][p]*a[p][c];  db_h = sum_p dl[h][p] */
int sunet_heads_bwd(const float* dlogits, const void* a, int a_pix_stride, const float* w0, const float* w1,
                    const float* w2, int nheads, void* dA, int dA_pix_stride, float* dw0, float* db0, float* dw1,
                    float* db1, float* dw2, float* db2, long long pixels, void* workspace, size_t workspace_bytes,
                    sunet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Losses: BCEWithLogitsLoss (train.py:78,195) and calc_selective_risk_image_b
 * (selective_loss.py:58-85), in two phases so that a data-parallel job can all-reduce the
 * three sums in between (the reference computes the loss on the gathered global batch).
 * ---------------------------------------------------------------------------------------- */
/* sums[0] = sum sigmoid(sel), sums[1] = sum bce(out,t)*sigmoid(sel), sums[2] = sum bce(aux,t)   (fp64)
 * any of out/sel/aux may be NULL (its sums stay 0).  pixels_out (optional, device) receives (double)pixels, so a
 * data-parallel caller can all-reduce [S, R, A, pixels] as one buffer and uneven shards need no host bookkeeping. */
int sunet_loss_sums(const float* out, const float* sel, const float* aux, const float* target, long long pixels,
                    double* sums, double* pixels_out, void* workspace, size_t workspace_bytes, sunet_stream_t stream);
/* results[0] = selective loss (risk + lamb*max(0, cov_target - c)^2), [1] = coverage c,
 * [2] = aux BCE mean, [3] = total;  P = global pixel count: global_pixels, or *global_pixels_dev when that device
 * pointer is not NULL (the all-reduced count; keeps a captured CUDA graph valid for any shard split) */
int sunet_loss_finalize(const double* sums, long long global_pixels, const double* global_pixels_dev, float lamb,
                        float target_coverage, float* results, sunet_stream_t stream);
/* per-pixel gradients given the (global) sums; g_sel / g_aux = upstream grads of the two losses
 * (device scalars, NULL = 1.0).  Outputs may be NULL.  global_pixels_dev as above. */
int sunet_loss_bwd(const float* out, const float* sel, const float* aux, const float* target, long long pixels,
                   const double* sums, long long global_pixels, const double* global_pixels_dev, float lamb,
                   float target_coverage, const float* g_sel, const float* g_aux, float* d_out, float* d_sel,
                   float* d_aux, sunet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Metrics: thresholding + Evaluator confusion matrix (train.py:211-239, eval.py:228-251,
 * utils/compute_metric.py:10-26).  Exact integer counts.
 *   counts[0..3] = confusion matrix rows=label cols=pred (only selected pixels if sel != NULL
 *                  and masked != 0), counts[4] = selected pixels, counts[5] = pixels seen
 * pred = out >= thr_out, selected = sel >= thr_sel: thresholds in logit space, bisected on the
 * host against numpy's float64 / float32 sigmoid so the masks are bit-identical.
 * label_dtype: 0 uint8, 1 float32 (truncated to uint8 like .astype('uint8')), 2 int64.
 * counts are ACCUMULATED (caller zeroes them at reset()). */
int sunet_metric_hist(const float* out, const float* sel, const void* label, int label_dtype, long long pixels,
                      float thr_out, float thr_sel, int masked, unsigned long long* counts, sunet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Evaluation ensemble (eval.py:116-123 model list, :209-222 mean of re-scaled outputs; scaling lambdas
 * eval.py:162-179).  mean[p] = (((s(map_0[p]) + s(map_1[p])) + ...) / n_models in float32, the order and rounding of
 * numpy's np.mean(np.asarray(outputs), axis=0) — bit-identical for scale NONE / CLIP / MINMAX (SIGMOID uses CUDA's
 * expf).  `maps` and `minmax` are HOST arrays of n_models DEVICE pointers (each map fp32 [pixels], 16-byte aligned;
 * minmax[m] -> device (min, max) of map m over the whole batch tensor as written by sunet_minmax_f32, MINMAX only).
 * ---------------------------------------------------------------------------------------- */
enum { SUNET_SCALE_NONE = 0, SUNET_SCALE_CLIP = 1, SUNET_SCALE_MINMAX = 2, SUNET_SCALE_SIGMOID = 3 };
int sunet_minmax_f32(const float* x, long long n, float* out_min_max, void* workspace, size_t workspace_bytes,
                     sunet_stream_t stream);
int sunet_ensemble_mean(const float* const* maps, const float* const* minmax, int n_models, long long pixels,
                        int scale, float* mean, sunet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * fp32 CHECK MODE (north_star: logits / loss within 1e-4 "in an fp32 check mode"): slow SIMT fp32 twins of every op
 * of the path, selected with SUNET_CHECK_FP32=1.  Activations NHWC fp32; parameters in the reference's own layouts
 * (Conv2d [co][ci][3][3], ConvTranspose2d [ci][co][2][2]); fp64 accumulation for every reduction over pixels.
 * Same reference citations as the fast entry points above (model.py:9-15,31,44-45,83,96-101; train.py:208).
 * ---------------------------------------------------------------------------------------- */
/* y[b,h,w,co] = bias[co] + sum_{r,s,ci} [x0 | x1][b,h+r-1,w+s-1,ci] * w[co][ci][r][s]  (x1 / bias may be NULL) */
int sunet_f32_conv3x3_fwd(const float* x0, int c0, const float* x1, int c1, const float* w, const float* bias, float* y,
                          int batch, int height, int width, int cout, sunet_stream_t stream);
/* input gradient; channels [0,c0) go to dx0, [c0,c0+c1) to dx1 (the [d_up | d_skip] halves of a decoder block) */
int sunet_f32_conv3x3_dgrad(const float* dy, const float* w, float* dx0, int c0, float* dx1, int c1, int batch,
                            int height, int width, int cout, sunet_stream_t stream);
int sunet_f32_conv3x3_wgrad(const float* dy, const float* x0, int c0, const float* x1, int c1, float* dw, int batch,
                            int height, int width, int cout, sunet_stream_t stream);
/* ConvTranspose2d(k2,s2): x [b][h][w][cin] -> y [b][2h][2w][cout]; (height, width) are the INPUT grid */
int sunet_f32_convT_fwd(const float* x, const float* w, const float* bias, float* y, int batch, int height, int width,
                        int cin, int cout, sunet_stream_t stream);
int sunet_f32_convT_dgrad(const float* dy, const float* w, float* dx, int batch, int height, int width, int cin,
                          int cout, sunet_stream_t stream);
int sunet_f32_convT_wgrad(const float* dy, const float* x, float* dw, float* dbias, int batch, int height, int width,
                          int cin, int cout, sunet_stream_t stream);
/* BatchNorm2d(train) statistics of the bias-free conv output y [pixels][channels] + running-stat update -> the
 * per-channel scale / shift / mean / invstd (same outputs as sunet_bn_finalize) */
int sunet_f32_bn_stats(const float* y, long long pixels, int channels, const float* gamma, const float* beta,
                       const float* conv_bias, float* running_mean, float* running_var, long long* num_batches_tracked,
                       float momentum, float eps, float* scale, float* shift, float* mean, float* invstd,
                       sunet_stream_t stream);
/* a = relu(y*scale + shift); pooled (optional) = MaxPool2d(2)(a) */
int sunet_f32_bn_relu_pool(const float* y, const float* scale, const float* shift, float* a, float* pooled, int batch,
                           int height, int width, int channels, sunet_stream_t stream);
/* backward of BN(train) + ReLU (+ skip / MaxPool fan-in): g = (dA + dPool routed to the first maximum) * (a > 0),
 * dgamma, dbeta, dy = scale*(g - sum g/n - xhat*sum(g*xhat)/n).  dA or dPool may be NULL.  workspace >= 16*channels B */
int sunet_f32_bn_relu_pool_bwd(const float* dA, const float* dPool, const float* y, const float* a, const float* scale,
                               const float* mean, const float* invstd, float* dgamma, float* dbeta, float* dy, int batch,
                               int height, int width, int channels, void* workspace, size_t workspace_bytes,
                               sunet_stream_t stream);
/* heads: w [nheads][channels], b [nheads], logits [nheads][pixels]; backward adds into dA when accumulate != 0 */
int sunet_f32_heads_fwd(const float* a, const float* w, const float* b, int nheads, float* logits, long long pixels,
                        int channels, sunet_stream_t stream);
int sunet_f32_heads_bwd(const float* dlogits, const float* a, const float* w, int nheads, float* dA, int accumulate,
                        float* dw, float* db, long long pixels, int channels, sunet_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer: Adam over a table of tensors in one launch (train.py:88-92,209)
 * ---------------------------------------------------------------------------------------- */
typedef struct sunet_adam_tensor {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  long long numel;
} sunet_adam_tensor;
/* table: DEVICE array of n_tensors entries.  step is 1-based.  lr_dev / step_dev (device scalars,
 * may be NULL) override lr / step so a CUDA graph of the whole step can be replayed unchanged. */
int sunet_adam_step(const sunet_adam_tensor* table, int n_tensors, long long max_numel, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int step, const float* lr_dev, const int* step_dev,
                    sunet_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SUNET_B200_H_ */
