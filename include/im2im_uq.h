/*
 * im2im_uq.h - C ABI of libim2im_uq.so: the B200 (sm_100a) hot path of aangelopoulos/im2im-uq.
 *
 * The reference is pure Python/PyTorch, so its "FFI" for this path is the set of torch calls made by the functions
 * cited at each entry point (paths relative to the reference root).  A maintainer binds these symbols with ctypes
 * (INTEGRATION.md shows the stub); im2im_uq_b200/_lib.py is exactly that binding.
 *
 * Conventions
 *   - plain C symbols, POD arguments, no torch types; every pointer named d_* is a DEVICE pointer
 *   - the caller allocates every buffer; the library never frees or keeps caller memory
 *   - all work is enqueued on the cudaStream_t passed as `stream` (void* so this header needs no CUDA include);
 *     no hidden cudaDeviceSynchronize, no default-stream use
 *   - return 0 on success, a negative IM2IM_E* code on failure; im2im_last_error() gives a thread-local message
 *   - re-entrant per (device, stream)
 */
#ifndef IM2IM_UQ_H_
#define IM2IM_UQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IM2IM_ABI_VERSION 1

#define IM2IM_OK 0
#define IM2IM_EINVAL (-22)     /* bad argument (null pointer, negative size, unsorted lambda grid ...) */
#define IM2IM_ERANGE (-34)     /* size outside what the kernel supports (n_lambdas, pixels per image) */
#define IM2IM_ECUDA (-5)       /* a CUDA runtime call or launch failed; see im2im_last_error() */
#define IM2IM_ENOTSUP (-95)    /* head kind / option not implemented */

/* head kinds (core/models/add_uncertainty.py:56-85 `uncertainty_type`): what the score planes mean.  After the outer
 * clamp of add_uncertainty.py:35-36 every head has the form
 *     upper = max(fl(fl(lam*w_up) + pred), pred+1e-6),   lower = min(fl(pred - fl(lam*w_lo)), pred-1e-6)
 * and only the widths differ:
 *   QUANTILES     d_lower/d_pred/d_upper = (lower, prediction, upper); w_up = max(upper, pred+1e-6) - pred,
 *                 w_lo = pred - min(lower, pred-1e-6).  uncertainty_type "quantiles", "quantiles_l1", "inn"
 *                 (finallayers/quantile_layer.py:34-44, quantile_l1_layer.py:34-44, inn_layer.py:30-40)
 *   RESIDUAL      d_pred/d_upper = (prediction, |residual| magnitude), d_lower ignored (may be NULL); w_up = w_lo = d_upper.
 *                 "residual_magnitude", "residual_magnitude_l1" (residual_magnitude_layer.py:28-36)
 *   GAUSSIAN      d_pred/d_upper = (mean, variance), d_lower ignored; w_up = w_lo = sqrt(variance).
 *                 "gaussian" (gaussian_layer.py:26-34)
 *   SOFTMAX_SETS  d_lower/d_pred/d_upper = (lower quantile, argmax prediction, upper quantile) as produced by
 *                 im2im_softmax_sets; w_up = relu(upper - pred), w_lo = relu(pred - lower).  "softmax"
 *                 (softmax_layer.py:27-53)
 */
#define IM2IM_HEAD_QUANTILES 0
#define IM2IM_HEAD_RESIDUAL 1
#define IM2IM_HEAD_GAUSSIAN 2
#define IM2IM_HEAD_SOFTMAX_SETS 3

/* flags for im2im_rcps_miss_counts */
#define IM2IM_RCPS_ZERO_OUTPUTS 1u  /* memset d_counts and d_totals on `stream` before the pass */
#define IM2IM_RCPS_FORCE_GENERIC 2u /* use the scalar-load kernel even when the bulk-copy fast path applies (tests) */

#define IM2IM_RCPS_MAX_LAMBDAS 8192

int im2im_abi_version(void);
const char* im2im_last_error(void);

/*
 * RCPS per-pixel miss counts for EVERY lambda of a sorted grid in ONE pass over the scores.
 *
 * Replaces, for all lambda steps at once, the body of the reference's sweep
 *     core/calibration/calibrate_model.py:134-136  (loop over lambdas -> get_rcps_losses_from_outputs :21-29)
 *       -> core/models/add_uncertainty.py:33-38             nested_sets_from_output (+/-1e-6 clamp)
 *       -> core/models/finallayers/quantile_layer.py:34-44  quantile_regression_nested_sets_from_output
 *       -> core/calibration/calibrate_model.py:76-80        fraction_missed_loss
 * and the dense table loop of core/scripts/eval.py:115-125 (get_loss_table).
 *
 * counts[i, j] = number of pixels of image i with  lower(lam_j) > y  or  upper(lam_j) < y, evaluated with the
 * reference's exact fp32 operation order (no FMA contraction).  The reference's per-image loss is exactly
 * float(counts[i,j]) / float(px) (im2im_rcps_loss_table).
 *
 *   d_lower/d_pred/d_upper/d_label : fp32 planes; image i of a plane starts at base + i*stride_* (in elements)
 *                                    and holds `px` = C*H*W contiguous values.  For the reference's
 *                                    (N,3,C,H,W) output tensor: d_pred = d_lower + px, d_upper = d_lower + 2*px,
 *                                    strides 3*px; labels stride px.
 *   d_lambdas  : DEVICE fp32[n_lambdas], finite, sorted ascending (host-computed, e.g. lambdas - dlambda)
 *   d_counts   : DEVICE int32[n_images, n_lambdas], row-major.  Must be zero on entry unless
 *                IM2IM_RCPS_ZERO_OUTPUTS is set (images that straddle two thread blocks are accumulated atomically)
 *   d_totals   : DEVICE uint64[n_lambdas]; column sums of counts are ADDED to it (so a caller can feed the
 *                calibration set in chunks); may be NULL
 * Limits: 1 <= n_lambdas <= IM2IM_RCPS_MAX_LAMBDAS, px < 2^24 (beyond that the reference's fp32 mean is no longer
 * an exact integer ratio).  The bulk-copy (TMA) fast path needs 16-byte aligned planes and px % 4 == 0; anything
 * else runs the scalar-load kernel with identical results.
 */
int im2im_rcps_miss_counts(const float* d_lower, const float* d_pred, const float* d_upper, const float* d_label,
                           int64_t n_images, int64_t px, int64_t stride_lower, int64_t stride_pred,
                           int64_t stride_upper, int64_t stride_label, const float* d_lambdas, int32_t n_lambdas,
                           int32_t head_kind, int32_t* d_counts, unsigned long long* d_totals, uint32_t flags,
                           void* stream);

/*
 * counts -> the reference's fp32 loss table: table[i,j] = float(counts[i,j]) / float(px) for j >= first_visited_col,
 * 0 for j < first_visited_col (columns the early-stopped sweep never wrote, calibrate_model.py:133,136,144).
 */
int im2im_rcps_loss_table(const int32_t* d_counts, int64_t n_images, int32_t n_lambdas, int64_t px,
                          int32_t first_visited_col, float* d_table, void* stream);

/* Same as im2im_rcps_loss_table but the first visited column is read from DEVICE memory (result[3] of im2im_rcps_decide),
 * so the table can be produced without a host round trip. */
int im2im_rcps_loss_table_dev(const int32_t* d_counts, int64_t n_images, int32_t n_lambdas, int64_t px,
                              const int32_t* d_first_visited_col, float* d_table, void* stream);

/*
 * Device-side screening of the reference's stopping rule (core/calibration/calibrate_model.py:137-140) from the exact
 * per-lambda totals: R_j = totals[j] / n_images_times_px; the reference's fp32 Rhat lies within +/- gamma*R_j; the rule
 * `Rhat >= alpha or HB_mu_plus(Rhat) > alpha` is certainly true above max(alpha32, r_hi+slack) and certainly false below
 * min(alpha32, r_lo-slack), where (r_lo, r_hi) brackets the level set of HB_mu_plus (host: bounds.hb_stop_bracket;
 * pass +inf when no muhat exceeds alpha).  The grid is scanned from the top like the reference.
 * d_result: DEVICE int32[4] = {stop index or -1, decided (1) / host must replay (0), first unsure column,
 *                              first visited column for im2im_rcps_loss_table_dev}.
 */
int im2im_rcps_decide(const unsigned long long* d_totals, int32_t n_lambdas, double n_images_times_px, double gamma,
                      double alpha32, double r_lo, double r_hi, double slack, int32_t* d_result, void* stream);

/*
 * Multi-GPU form of im2im_rcps_decide: the all-reduce of the per-lambda totals fused with the decision, over NVLink peer
 * memory instead of a NCCL call (the reduction the reference does implicitly by holding the whole calibration set in one
 * process, calibrate_model.py:136-138).  One-shot push all-reduce: this rank stores d_local_totals into its slot of every
 * peer's mailbox, publishes a release flag per peer, waits for all peers' flags, sums its own mailbox, decides.
 *   d_peer_mailboxes : DEVICE array [world] of pointers; entry r = rank r's mailbox uint64[2][world][n_lambdas] as mapped
 *                      into THIS process (symmetric / peer memory; zero-initialised is not required)
 *   d_peer_flags     : DEVICE array [world] of pointers; entry r = rank r's flag array uint32[world], zero before first use
 *   d_epoch          : DEVICE uint32, local, zero before first use; advanced by the call (CUDA-graph replays need no new
 *                      arguments).  Every rank must make the same sequence of calls.
 *   d_totals_out     : DEVICE uint64[n_lambdas], the reduced totals (for the host replay of guard-band columns)
 *   d_result         : as im2im_rcps_decide; {-2, -2, -2, 0} when a peer did not arrive within ~15 s (the kernel gives up
 *                      instead of hanging the GPU)
 */
int im2im_rcps_decide_p2p(const unsigned long long* d_local_totals, unsigned long long* const* d_peer_mailboxes,
                          unsigned* const* d_peer_flags, unsigned* d_epoch, int32_t rank, int32_t world, int32_t n_lambdas,
                          double n_images_times_px, double gamma, double alpha32, double r_lo, double r_hi, double slack,
                          unsigned long long* d_totals_out, int32_t* d_result, void* stream);

/*
 * The WHOLE device side of one calibration (core/calibration/calibrate_model.py:130-145) in ONE kernel launch:
 * im2im_rcps_miss_counts + the exchange of the per-lambda totals between GPUs + im2im_rcps_decide +
 * im2im_rcps_loss_table_dev, with no memset of the outputs and no device->host copy.
 *   - every row of d_counts is written in full by the thread block that owns the image's first tile (an image that
 *     straddles two blocks travels through a per-block partial row in the workspace), so d_counts need not be zeroed;
 *   - block totals are reduced with uint64 atomics into the workspace; the LAST block to finish (ticket) sums them, and
 *     when world > 1 all-reduces them with the one-shot push protocol of im2im_rcps_decide_p2p over NVLink peer memory;
 *   - it screens the stopping rule exactly like im2im_rcps_decide, writes d_result AND h_result_mapped (mapped pinned
 *     host memory: int32[8] = {result[0..3], epoch tag, -, -, -}; the tag is stored last, after a system fence, so a host
 *     thread spinning on h_result_mapped[4] (im2im_host_wait_flag) sees a complete result without a stream synchronize);
 *   - the other blocks then convert the counts rows they own into d_table (0 left of the first visited column).
 *   d_table, d_totals_out, h_result_mapped may be NULL.
 *   d_workspace : DEVICE, im2im_rcps_fused_workspace_bytes(n_lambdas) bytes, ZERO-FILLED once by the caller before the
 *                 first call and then left alone (the kernel restores its invariants; the first uint32 is the launch
 *                 epoch, advanced by every call so that CUDA-graph replays need no new arguments).  With world > 1 every
 *                 rank must make the same sequence of calls on a workspace of the same age.
 *   peer tables : as im2im_rcps_decide_p2p (ignored when world == 1).  peer_timeout_s > 0 bounds the wait for the peers'
 *                 flags (result = {-2,-2,-2,0}: the ranks are out of step, tear the peer buffers down and rebuild);
 *                 0 waits for ever.
 * Launched cooperatively (all blocks resident).  Returns IM2IM_ENOTSUP - before launching anything - when the bulk-copy
 * fast path does not apply (unaligned planes, px % 4 != 0, lambda grid too long for the staging ring): callers then
 * use the separate entry points above.
 */
size_t im2im_rcps_fused_workspace_bytes(int32_t n_lambdas);
/* IM2IM_OK when im2im_rcps_calibrate_fused would take these planes, IM2IM_ENOTSUP otherwise; launches nothing (ranks of a
 * multi-GPU job agree on the path with it before the first collective call). */
int im2im_rcps_calibrate_fused_check(const float* d_lower, const float* d_pred, const float* d_upper, const float* d_label,
                                     int64_t n_images, int64_t px, int64_t stride_lower, int64_t stride_pred,
                                     int64_t stride_upper, int64_t stride_label, int32_t n_lambdas, int32_t head_kind);
int im2im_rcps_calibrate_fused(const float* d_lower, const float* d_pred, const float* d_upper, const float* d_label,
                               int64_t n_images, int64_t px, int64_t stride_lower, int64_t stride_pred,
                               int64_t stride_upper, int64_t stride_label, const float* d_lambdas, int32_t n_lambdas,
                               int32_t head_kind, int32_t* d_counts, float* d_table, unsigned long long* d_totals_out,
                               double n_images_times_px, double gamma, double alpha32, double r_lo, double r_hi,
                               double slack, void* d_workspace, size_t workspace_bytes,
                               unsigned long long* const* d_peer_mailboxes, unsigned* const* d_peer_flags, int32_t rank,
                               int32_t world, double peer_timeout_s, int32_t* d_result, int32_t* h_result_mapped,
                               void* stream);

/* HOST helper for the fused step: spin (pause loop) until *h_flag - expected >= 0 or spin_us microseconds have passed.
 * Returns 0 when the flag arrived, 1 on timeout (the caller falls back to synchronising the stream). */
int im2im_host_wait_flag(const volatile int32_t* h_flag, int32_t expected, int64_t spin_us);

/*
 * Interval endpoints at one lambda: ModelWithUncertainty.nested_sets_from_output
 * (core/models/add_uncertainty.py:33-38 over core/models/finallayers/quantile_layer.py:34-44).
 * Writes lower/upper as dense (n_images, px) fp32; the prediction plane is returned by the caller as a view.
 * The reference clamps its `output` argument in place (quantile_layer.py:39-40: lower plane = min(lower, pred-1e-6),
 * upper plane = max(upper, pred+1e-6)); with write_back_clamp != 0 the same side effect is applied to
 * d_lower / d_upper, otherwise the inputs are left untouched.
 */
int im2im_quantile_nested_sets(float* d_lower, const float* d_pred, float* d_upper, int64_t n_images, int64_t px,
                               int64_t stride_lower, int64_t stride_pred, int64_t stride_upper, float lam,
                               int32_t write_back_clamp, float* d_lower_out, float* d_upper_out, void* stream);

/*
 * The same for any head kind (planes as described at IM2IM_HEAD_*): the head's own set function
 *   gaussian_layer.py:26-34, residual_magnitude_layer.py:28-36, residual_magnitude_l1_layer.py:28-36,
 *   quantile_l1_layer.py:34-44, inn_layer.py:30-40, softmax_layer.py:50-51
 * followed by the +/-1e-6 clamp of ModelWithUncertainty.nested_sets_from_output (add_uncertainty.py:35-36).
 * write_back_clamp only applies to IM2IM_HEAD_QUANTILES (the only set functions that mutate their argument).
 */
int im2im_nested_sets(int32_t head_kind, float* d_lower, const float* d_pred, float* d_upper, int64_t n_images,
                      int64_t px, int64_t stride_lower, int64_t stride_pred, int64_t stride_upper, float lam,
                      int32_t write_back_clamp, float* d_lower_out, float* d_upper_out, void* stream);

/*
 * Softmax head, lambda-independent half of softmax_nested_sets_from_output (softmax_layer.py:34-48): per pixel
 * softmax over the n_classes logits, cumulative sum, lower/upper quantile = #{cumsum <= 0.05 / 0.95}/n_classes,
 * prediction = argmax/n_classes, the +/- 1/n_classes separation of :45-46 and the clamp to [0,1].
 *   d_logits : DEVICE fp32; class k of pixel j of image i at d_logits[i*stride_image + k*stride_class + j]
 *              (the head's (B, K, 1, H, W) tensor: stride_class = inner = H*W, stride_image = K*inner)
 *   d_sets   : DEVICE fp32 [n_images, 3, inner] = (lower quantile, prediction, upper quantile), the planes of
 *              IM2IM_HEAD_SOFTMAX_SETS.  The reference recomputes this for every lambda step; here it is computed once.
 * Parity: the outputs are multiples of 1/n_classes; they equal the reference's except where a cumulative sum lies
 * within rounding distance of 0.05 / 0.95 (exp and the summation order differ between torch's CPU, torch's CUDA and
 * this kernel by an ulp).  1 <= n_classes <= 64.
 */
int im2im_softmax_sets(const float* d_logits, int64_t n_images, int32_t n_classes, int64_t inner, int64_t stride_image,
                       int64_t stride_class, float* d_sets, void* stream);

/*
 * Per-pixel miss map at one lambda, summed over images: map[k] = #{i : pixel k of image i is missed}
 * (get_rcps_metrics_from_outputs, core/calibration/calibrate_model.py:47,55 "spatial_miscoverage" numerator).
 * d_map: DEVICE int32[px], accumulated into (zero it first, or set IM2IM_RCPS_ZERO_OUTPUTS).
 */
int im2im_rcps_miss_map(const float* d_lower, const float* d_pred, const float* d_upper, const float* d_label,
                        int64_t n_images, int64_t px, int64_t stride_lower, int64_t stride_pred,
                        int64_t stride_upper, int64_t stride_label, float lam, int32_t head_kind, int32_t* d_map,
                        uint32_t flags, void* stream);

/*
 * fraction_missed_loss on ALREADY COMPUTED interval endpoints (core/calibration/calibrate_model.py:76-80):
 * counts[i] = #{k : lower_edge[i,k] > label[i,k] or upper_edge[i,k] < label[i,k]}; the loss is counts[i]/px.
 * d_counts: DEVICE int32[n_images], overwritten.  n_images <= 65535 per call.
 */
int im2im_fraction_missed_counts(const float* d_lower_edge, const float* d_upper_edge, const float* d_label,
                                 int64_t n_images, int64_t px, int64_t stride_lower, int64_t stride_upper,
                                 int64_t stride_label, int32_t* d_counts, void* stream);

/*
 * 3x3 (pad 1) or 1x1 convolution, NHWC bf16, as an implicit GEMM on tcgen05 tensor cores (fp32 accumulation in TMEM).
 * Replaces the library convolution + (folded) BatchNorm + ReLU of the reference's trunk for inference:
 *   core/models/trunks/unet_parts.py:16-21 (DoubleConv), :90 (OutConv 1x1); the channel concatenation of
 *   unet_parts.py:68 `torch.cat([x2, x1], dim=1)` is expressed by passing the two tensors (x1 = skip, x2 = upsampled).
 *   d_x1 / d_x2 : DEVICE bf16 [B,H,W,c_in1] / [B,H,W,c_in2] (c_in2 = 0 and d_x2 = NULL for a single input); channel
 *                 counts must be multiples of 64
 *   d_weight    : DEVICE bf16 [c_out, taps, c_in1+c_in2] (taps = 9: ky*3+kx row-major; 1 for 1x1)
 *   d_bias      : DEVICE fp32 [c_out] or NULL;  relu != 0 applies max(x, 0)
 *   outputs     : DEVICE NHWC [B,H,W,c_out] as bf16 and/or fp32 (either may be NULL, not both); c_out % 32 == 0
 * Which kernel runs is decided per layer shape: 3x3 layers whose map tiles into 8 x 16 pixel tiles
 * (W % 8 == 0, H % 16 == 0) and whose weights fit (resident, or streamed up to 256 input channels) run on the halo kernel as
 * clusters of two CTAs (tcgen05 cta_group::2); everything else on the persistent kernel.  The result does not depend on the
 * CTA pairing (bit-identical, tests/test_conv_gpu.py); INTEGRATION.md lists the IM2IM_* switches that force a route.
 */
int im2im_conv_igemm_bf16(const void* d_x1, int32_t c_in1, const void* d_x2, int32_t c_in2, const void* d_weight,
                          const float* d_bias, int32_t B, int32_t H, int32_t W, int32_t c_out, int32_t taps,
                          int32_t relu, void* d_out_bf16, float* d_out_f32, void* stream);

/* im2im_conv_igemm_bf16 that ALSO writes maxpool2x2 of its (bias + ReLU'd) output - Down = MaxPool2d(2) -> DoubleConv
 * (unet_parts.py:27-36) without the separate pooling pass over the skip tensor: d_pool_out bf16 NHWC [B, H/2, W/2, c_out].
 * *h_pooled (HOST int) = 1 when the pooled tensor was written (layer on the halo kernel, H and W even); 0: only the
 * convolution ran and the caller pools (im2im_maxpool2x2_bf16).  Bit-identical to the two-kernel sequence. */
int im2im_conv_igemm_bf16_pool(const void* d_x1, int32_t c_in1, const void* d_x2, int32_t c_in2, const void* d_weight,
                               const float* d_bias, int32_t B, int32_t H, int32_t W, int32_t c_out, int32_t taps,
                               int32_t relu, void* d_out_bf16, void* d_pool_out, int32_t* h_pooled, void* stream);
/*
 * im2im_conv_igemm_bf16 (no bias, no ReLU, bf16 output) with per-channel statistics accumulated in the convolution's
 * epilogue, from the bf16 values it stores - one pass over the activations less per BatchNorm layer of the training step
 * (core/models/trunks/unet_parts.py:17,20 in train mode; autograd of the same in core/scripts/train.py:160):
 *   stat_mode 1 (forward)   d_stat_sums[c] += z, d_stat_sums[c_out + c] += z*z      = im2im_channel_stats_bf16 of the output
 *   stat_mode 2 (backward)  the convolution is a data gradient whose output dy feeds BatchNorm+ReLU backward of the layer
 *                           with saved pre-normalisation output d_bn_z [B,H,W,c_out] and (gamma, beta, mean, rstd):
 *                           g = dy * (bn_z*gamma*rstd + beta - mean*gamma*rstd > 0) is stored INSTEAD of dy, and
 *                           d_stat_sums[c] += g, d_stat_sums[c_out + c] += rstd*(sum(g*bn_z) - mean*sum(g))
 *                           (the reduction half of im2im_bn_relu_bwd_bf16; finish with im2im_bn_relu_bwd_apply_bf16).
 * d_stat_sums is ACCUMULATED into (zero it first).  *h_fused (HOST int) = 1 when the statistics were produced (the layer
 * ran on the halo kernel, or - stat_mode 1 only - on the persistent kernel with >= 256 input channels); 0 means only the convolution ran - stored values are then plain dy / z and the caller runs the
 * separate reduction.
 */
int im2im_conv_igemm_bf16_stats(const void* d_x1, int32_t c_in1, const void* d_x2, int32_t c_in2, const void* d_weight,
                                int32_t B, int32_t H, int32_t W, int32_t c_out, int32_t taps, void* d_out_bf16,
                                int32_t stat_mode, float* d_stat_sums, const void* d_bn_z, const float* d_bn_gamma,
                                const float* d_bn_beta, const float* d_bn_mean, const float* d_bn_rstd, int32_t* h_fused,
                                void* stream);

/*
 * The same convolution in the REFERENCE'S PRECISION: the reference's modules are fp32 (unet_parts.py:16-21, no autocast
 * anywhere) and torch's cuDNN convolutions run them as TF32 on the GPU; here fp32 NHWC activations and fp32 weights are
 * staged by TMA as they are and multiplied with tcgen05 kind::tf32 (fp32 accumulation), output fp32 NHWC with every value
 * rounded (to nearest) onto the TF32 grid, so that the next layer's MMA reads exactly what was stored.
 *   d_x1 / d_x2 : DEVICE fp32 [B,H,W,c_in1] / [B,H,W,c_in2]; channel counts multiples of 32
 *   d_weight    : DEVICE fp32 [c_out, taps, c_in1+c_in2];  d_bias fp32 [c_out] or NULL;  d_out fp32 NHWC, c_out % 32 == 0
 * About half the bf16 path's throughput (the TF32 tensor rate); it is the parity mode of the inference engine.
 */
int im2im_conv_igemm_tf32(const float* d_x1, int32_t c_in1, const float* d_x2, int32_t c_in2, const float* d_weight,
                          const float* d_bias, int32_t B, int32_t H, int32_t W, int32_t c_out, int32_t taps,
                          int32_t relu, float* d_out, void* stream);

/*
 * Weight gradient of the convolution above (autograd of nn.Conv2d inside core/models/trunks/unet_parts.py:16-21, as
 * driven by core/scripts/train.py:160 `loss.backward()`):  dW[co, tap, ci] += sum_p dZ[p, co] * X[p + shift(tap), ci].
 *   d_x  : DEVICE bf16 NHWC [B,H,W,c_in]  (the conv's input);   d_dz : DEVICE bf16 NHWC [B,H,W,c_out] (grad of its output)
 *   d_dw : DEVICE fp32 [c_out, taps, c_in], ACCUMULATED into (zero it first); c_in, c_out multiples of 64
 * tcgen05 GEMM with K = pixels, operands consumed MN-major straight from the NHWC tensors; split-K with fp32 atomics.
 */
int im2im_conv_wgrad_bf16(const void* d_x, const void* d_dz, int32_t B, int32_t H, int32_t W, int32_t c_in,
                          int32_t c_out, int32_t taps, float* d_dw, void* stream);

/*
 * The CUDA-core passes around the tcgen05 convolutions of the UNet forward (all NHWC bf16 unless noted):
 *   im2im_conv_first_bf16          first 3x3 conv of DoubleConv(n_channels_in, 64) (core/models/trunks/unet.py:20):
 *                                  x fp32 NCHW [B,c_in<=8,H,W], weight fp32 [c_out,c_in,3,3] (+bias, ReLU) -> bf16 NHWC
 *   im2im_maxpool2x2_bf16          nn.MaxPool2d(2) (core/models/trunks/unet_parts.py:34)
 *   im2im_upsample2x_bilinear_bf16 nn.Upsample(x2, bilinear, align_corners=True) + F.pad to (H_out, W_out)
 *                                  (unet_parts.py:50,63-64; pad split floor/ceil like the reference)
 *   im2im_head_conv3x3_f32         QuantileRegressionLayer (core/models/finallayers/quantile_layer.py:15-20): the three
 *                                  3x3 convs stacked as n_out = 3*C_out output planes; x bf16 NHWC with c_mid channels
 *                                  used out of a row stride of c_stride (>= c_mid) elements; d_tap_bias (optional,
 *                                  [n_out][9]) is added once per in-range tap - the bias of a 1x1 conv folded into
 *                                  d_weight (inference: OutConv composed with the head, unet.py:31 + quantile_layer.py:20),
 *                                  weight fp32 [n_out,c_mid,3,3] -> fp32 [B,n_out,H,W] == the (B,3,C_out,H,W) tensor
 */
/* fp32 nn.Conv2d weight [c_out,c_in,k,k] -> the bf16 operands of im2im_conv_igemm_bf16: d_out_fwd [c_out,taps,c_in] and
 * (optional) d_out_bwd [c_in,taps,c_out] with reversed taps, the weights of the data-gradient convolution. */
int im2im_pack_conv_weights(const float* d_weight, int32_t c_out, int32_t c_in, int32_t taps, void* d_out_fwd,
                            void* d_out_bwd, void* stream);
int im2im_conv_first_bf16(const float* d_x, const float* d_weight, const float* d_bias, int32_t B, int32_t c_in,
                          int32_t H, int32_t W, int32_t c_out, int32_t relu, void* d_out, void* stream);
int im2im_maxpool2x2_bf16(const void* d_x, int32_t B, int32_t H, int32_t W, int32_t C, void* d_out, void* stream);
int im2im_upsample2x_bilinear_bf16(const void* d_x, int32_t B, int32_t h, int32_t w, int32_t C, int32_t H_out,
                                   int32_t W_out, void* d_out, void* stream);
int im2im_head_conv3x3_f32(const void* d_x, const float* d_weight, const float* d_bias, const float* d_tap_bias,
                           int32_t B, int32_t H, int32_t W, int32_t c_mid, int32_t c_stride, int32_t n_out, float* d_out,
                           void* stream);
/* The same with the head's output activation fused: planes >= act_from_plane get act_kind (0 none, 1 relu, 2 abs) -
 * GaussianRegressionLayer.forward (gaussian_layer.py:17-19: relu on the variance conv) and ResidualMagnitude(L1)Layer
 * .forward (residual_magnitude_layer.py:17-19: abs on the magnitude conv).  n_out in {2, 3, 4, 6, 9}. */
int im2im_head_conv3x3_act_f32(const void* d_x, const float* d_weight, const float* d_bias, const float* d_tap_bias,
                               int32_t B, int32_t H, int32_t W, int32_t c_mid, int32_t c_stride, int32_t n_out,
                               int32_t act_kind, int32_t act_from_plane, float* d_out, void* stream);
/* fp32 NHWC (tf32 mode) forms of the four CUDA-core kernels above: same arithmetic, activations stored as fp32 rounded onto
 * the TF32 grid.  unet.py:20 first conv, unet_parts.py:34 max pool, :50,:63 bilinear x2 + pad, quantile_layer.py:15-20 head. */
int im2im_conv_first_nhwc_f32(const float* d_x, const float* d_weight, const float* d_bias, int32_t B, int32_t c_in,
                              int32_t H, int32_t W, int32_t c_out, int32_t relu, float* d_out, void* stream);
int im2im_maxpool2x2_nhwc_f32(const float* d_x, int32_t B, int32_t H, int32_t W, int32_t C, float* d_out, void* stream);
int im2im_upsample2x_bilinear_nhwc_f32(const float* d_x, int32_t B, int32_t h, int32_t w, int32_t C, int32_t H_out,
                                       int32_t W_out, float* d_out, void* stream);
int im2im_head_conv3x3_act_nhwc_f32(const float* d_x, const float* d_weight, const float* d_bias, const float* d_tap_bias,
                                    int32_t B, int32_t H, int32_t W, int32_t c_mid, int32_t c_stride, int32_t n_out,
                                    int32_t act_kind, int32_t act_from_plane, float* d_out, void* stream);

/* The same head on TENSOR CORES (tcgen05 halo kernel): x bf16 NHWC with exactly 64 channels per pixel (the reference's 32
 * feature channels zero-padded), d_weight bf16 [64, 9, 64] = the stacked head convolutions packed like
 * im2im_pack_conv_weights with rows >= n_real and input channels >= c_mid zero, d_bias fp32 [>= n_real] or NULL; output
 * fp32 planes [B, n_real, H, W], n_real <= 32.  Needs W % 8 == 0 and H % 16 == 0 (IM2IM_ENOTSUP otherwise: use the
 * CUDA-core entry points above).  The head's weights enter as bf16 here (fp32 above); accumulation is fp32. */
int im2im_head_conv3x3_tc_f32(const void* d_x, const void* d_weight, const float* d_bias, int32_t B, int32_t H, int32_t W,
                              int32_t n_real, int32_t act_kind, int32_t act_from_plane, float* d_out, void* stream);
/* im2im_head_conv3x3_tc_f32 with the trunk's 1x1 OutConv (unet.py:45, unet_parts.py:87-93) FOLDED into the head's weights, so
 * that the head reads the 64-channel feature map directly and the (B, H, W, 32) OutConv output is never written or read:
 * d_weight[o][tap][c] = sum_m head[o][m][tap] * out[m][c] (bf16, packed as above), d_bias[o] = head bias + sum over the nine
 * taps of t[o][tap], t[o][tap] = sum_m head[o][m][tap] * out_bias[m].  A border pixel subtracts the tap terms that fall into the
 * zero padding (where the reference's OutConv output, bias included, is padding, not bias): d_tap_bias is fp32 [9][n_real],
 * d_tap_bias[cls][o] = sum of t[o][tap] over the out-of-range taps of border class cls = 3 * (0 inside | 1 first row | 2 last
 * row) + (0 inside | 1 first column | 2 last column); n_real <= 7.  The result equals head(OutConv(x)) up to fp32 summation
 * order and one bf16 rounding less. */
int im2im_head_conv3x3_tc_folded_f32(const void* d_x, const void* d_weight, const float* d_bias, const float* d_tap_bias,
                                     int32_t B, int32_t H, int32_t W, int32_t n_real, int32_t act_kind, int32_t act_from_plane,
                                     float* d_out, void* stream);
/* Head -> calibration histogram without the head tensor (streaming calibrate_model: replaces writing outputs[counter:...]
 * at calibrate_model.py:121-123 and re-reading them at :134-136 for a one-channel quantile head, quantile_layer.py:19-21).
 * The same convolution as above with n_real == 3 (lower, prediction, upper); in its epilogue every pixel is ranked against
 * the ascending lambda grid from the fp32 values im2im_head_conv3x3_tc_f32 would store - bit for bit the rank
 * im2im_rcps_miss_counts gives that pixel - and booked into d_hist[b][k], k = 1..n_lambdas (u32 [B][n_lambdas + 1],
 * ACCUMULATED into: zero it once; im2im_rcps_counts_from_hist leaves it zero again).  d_labels fp32 [B, 1, H, W].
 * d_out_or_null: also write the planes (tests), or NULL.  d_tap_bias_or_null: as im2im_head_conv3x3_tc_folded_f32.
 * IM2IM_ENOTSUP for other plane counts. */
int im2im_head_conv3x3_tc_hist(const void* d_x, const void* d_weight, const float* d_bias, const float* d_tap_bias_or_null,
                               int32_t B, int32_t H, int32_t W, int32_t n_real, int32_t act_kind, int32_t act_from_plane,
                               float* d_out_or_null, const float* d_labels, const float* d_lambdas_sorted, int32_t n_lambdas,
                               uint32_t* d_hist, void* stream);
/* d_hist (above) -> d_counts int32 [n_images, n_lambdas] exactly as im2im_rcps_miss_counts writes them (counts[i][j] =
 * #pixels of image i missed at lambda_j), d_totals_or_null u64 [n_lambdas] += column sums; zeroes d_hist. */
int im2im_rcps_counts_from_hist(uint32_t* d_hist, int64_t n_images, int32_t n_lambdas, int32_t* d_counts,
                                unsigned long long* d_totals_or_null, void* stream);
/* fp32 planes [B, n_planes, H, W] -> bf16 NHWC [B, H, W, 64] (channels >= n_planes zero): the head's output gradient as an
 * operand of im2im_conv_wgrad_bf16 / im2im_conv_igemm_bf16 (head weight / data gradient on tensor cores). */
int im2im_planar_to_nhwc64_bf16(const float* d_src, int32_t n_planes, int32_t B, int32_t H, int32_t W, void* d_dst,
                                void* stream);
/* The same for n_planes <= 8 into a buffer whose channels 8..63 the caller keeps zero across calls: only the first 16 bytes
 * of each 128-byte pixel row are written (8x fewer bytes per training step). */
int im2im_planar_to_nhwc64_first8_bf16(const float* d_src, int32_t n_planes, int32_t B, int32_t H, int32_t W, void* d_dst,
                                       void* stream);

/*
 * Training-side passes of the UNet path (autograd of core/scripts/train.py:152-162 through the modules of
 * core/models/trunks/unet_parts.py and core/models/finallayers/quantile_layer.py).  NHWC bf16 activations, fp32 params.
 *
 *   im2im_channel_stats_bf16     sums[c] += sum_p z[p,c], sums[C+c] += sum_p z[p,c]^2            (BatchNorm batch statistics)
 *   im2im_bn_finalize            sums -> scale/shift for y = relu(z*scale+shift), saved mean/rstd, running-stat update
 *                                (nn.BatchNorm2d train mode, unet_parts.py:17,20; conv_bias only enters running_mean)
 *   im2im_bn_apply_relu_bf16     y = relu(z*scale + shift)                                       (unet_parts.py:17-18)
 *   im2im_bn_relu_bwd_bf16       ReLU + BatchNorm backward from dy and the saved pre-normalisation z:
 *                                sums[c] = dbeta, sums[C+c] = dgamma, dz written
 *   im2im_maxpool2x2_bwd_bf16    gradient of nn.MaxPool2d(2) (first maximal element, optional accumulate)
 *   im2im_upsample2x_bilinear_bwd_bf16  gradient of Upsample(x2, bilinear, align_corners=True) + F.pad
 *   im2im_quantile_loss_f32      quantile_regression_loss_fn (quantile_layer.py:23-32, pinball.py:12-24): loss_parts[3]
 *                                (double: sum pinball_lo, sum pinball_hi, sum squared error) and d loss / d pred
 *   im2im_adam_step_f32          torch.optim.Adam step on a flat fp32 buffer (train.py:120,162); grad_scale multiplies
 *                                the gradient first (1/world_size after the NCCL all-reduce)
 *   im2im_head_bwd               QuantileRegressionLayer backward: dm (bf16, row stride c_stride, zero padded),
 *                                dW [n_out,c_mid,3,3] and db [n_out] accumulated
 *   im2im_conv_first_wgrad       weight gradient of the first 3x3 conv (x fp32 NCHW, dz bf16 NHWC), accumulated
 */
int im2im_channel_stats_bf16(const void* d_z, int64_t n_pix, int32_t C, float* d_sums, void* stream);
int im2im_bn_finalize(const float* d_sums, int64_t count, const float* d_conv_bias, const float* d_gamma,
                      const float* d_beta, float eps, float momentum, int32_t C, float* d_running_mean,
                      float* d_running_var, float* d_scale, float* d_shift, float* d_save_mean, float* d_save_rstd,
                      void* stream);
int im2im_bn_apply_relu_bf16(const void* d_z, const float* d_scale, const float* d_shift, int64_t n_pix, int32_t C,
                             void* d_y, void* stream);
int im2im_bn_relu_bwd_bf16(const void* d_dy, const void* d_z, const float* d_gamma, const float* d_beta,
                           const float* d_mean, const float* d_rstd, int64_t n_pix, int32_t C, float* d_sums, void* d_dz,
                           void* stream);
/* The second half of im2im_bn_relu_bwd_bf16 on its own: dz from sums that are already complete.  premasked != 0: d_g is
 * g = dy * relu_mask as written by im2im_conv_igemm_bf16_stats(stat_mode 2), which also produced d_sums. */
int im2im_bn_relu_bwd_apply_bf16(const void* d_g, const void* d_z, const float* d_gamma, const float* d_beta,
                                 const float* d_mean, const float* d_rstd, const float* d_sums, int64_t n_pix, int32_t C,
                                 int32_t premasked, void* d_dz, void* stream);
/* Skip layers (the block output feeds the skip connection AND the next block's MaxPool2d(2), core/models/trunks/unet.py:35-39,
 * unet_parts.py:34): BatchNorm+ReLU and the pool in one pass - y = relu(z*scale + shift), pooled = maxpool2x2(y) - and,
 * backward, the pool's gradient folded into the BatchNorm backward: dy = d_skip + (pooled gradient routed to the first
 * maximal element of each window, rounded to bf16 as the two-kernel sequence stores it), then exactly
 * im2im_bn_relu_bwd_bf16 on dy.  Replaces im2im_maxpool2x2_bf16 / im2im_maxpool2x2_bwd_bf16(accumulate) + the BatchNorm
 * kernels for those layers (5.5 bytes per activation value less traffic per step).  H and W even (IM2IM_ENOTSUP otherwise). */
int im2im_bn_apply_relu_pool_bf16(const void* d_z, const float* d_scale, const float* d_shift, int32_t B, int32_t H, int32_t W,
                                  int32_t C, void* d_y, void* d_pooled, void* stream);
int im2im_bn_relu_pool_bwd_bf16(const void* d_dskip, const void* d_dpooled, const void* d_z, const float* d_gamma,
                                const float* d_beta, const float* d_mean, const float* d_rstd, int32_t B, int32_t H,
                                int32_t W, int32_t C, float* d_sums, void* d_dz, void* stream);
int im2im_maxpool2x2_bwd_bf16(const void* d_x, const void* d_dy, int32_t B, int32_t H, int32_t W, int32_t C,
                              int32_t accumulate, void* d_dx, void* stream);
int im2im_upsample2x_bilinear_bwd_bf16(const void* d_du, int32_t B, int32_t h, int32_t w, int32_t C, int32_t H_out,
                                       int32_t W_out, void* d_dx, void* stream);
int im2im_quantile_loss_f32(const float* d_pred, const float* d_target, int64_t n_images, int64_t px, float q_lo,
                            float q_hi, float w_lo, float w_hi, float w_mse, float* d_dpred, double* d_loss_parts,
                            void* stream);
/*
 * Training losses of every affine head in one fused pass (value parts + gradient), pred (B, planes, px) fp32:
 *   IM2IM_LOSS_QUANTILES     quantile_layer.py:23-32      parts = (sum pinball_lo, sum pinball_hi, sum sq err); w = (w_lo, w_hi, w_mse)
 *   IM2IM_LOSS_QUANTILES_L1  quantile_l1_layer.py:23-32   parts = (sum pinball_lo, sum pinball_hi, sum |err|);  w = (w_lo, w_hi, w_mse)
 *   IM2IM_LOSS_GAUSSIAN      gaussian_layer.py:20-24      parts = (sum 0.5*(log v + d^2/v), 0, 0), v = max(var, 1e-6); w0 = 1
 *   IM2IM_LOSS_RESIDUAL      residual_magnitude_layer.py:20-26     parts = (sum (p-y)^2, sum (r-|y-p|)^2, 0); w = (1, 1)
 *   IM2IM_LOSS_RESIDUAL_L1   residual_magnitude_l1_layer.py:20-26  parts = (sum |p-y|,  sum (r-|y-p|)^2, 0); w = (1, 1)
 *   IM2IM_LOSS_INN           inn_layer.py:23-28 + losses/inn.py:11-14  parts = (sum (p-y)^2, sum relu(y-u)^2+relu(l-y)^2+beta|u-l|, 0)
 * loss = sum_k w_k * parts[k] / (n_images*px); d_dpred (may be NULL) receives d loss / d pred with the same weights.
 */
#define IM2IM_LOSS_QUANTILES 0
#define IM2IM_LOSS_QUANTILES_L1 1
#define IM2IM_LOSS_GAUSSIAN 2
#define IM2IM_LOSS_RESIDUAL 3
#define IM2IM_LOSS_RESIDUAL_L1 4
#define IM2IM_LOSS_INN 5
int im2im_head_loss_f32(int32_t loss_kind, const float* d_pred, const float* d_target, int64_t n_images, int64_t px,
                        float q_lo, float q_hi, float w0, float w1, float w2, float beta, float* d_dpred,
                        double* d_loss_parts, void* stream);
int im2im_adam_step_f32(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n, float lr,
                        float beta1, float beta2, float eps, int32_t step, float grad_scale, void* stream);
/* im2im_adam_step_f32 with the step counter in DEVICE memory, so the whole training step can be replayed from a CUDA
 * graph: d_state is float[3] = {step, 1-beta1^step, sqrt(1-beta2^step)}; the call first advances it (step += 1), then
 * applies the update.  Zero-initialise d_state before the first step. */
int im2im_adam_step_dev_f32(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n,
                            float lr, float beta1, float beta2, float eps, float* d_state, float grad_scale,
                            void* stream);
int im2im_head_bwd(const float* d_dout, const void* d_m, const float* d_weight, int32_t B, int32_t H, int32_t W,
                   int32_t c_mid, int32_t c_stride, int32_t n_out, void* d_dm, float* d_dw, float* d_db, void* stream);
int im2im_conv_first_wgrad(const float* d_x, const void* d_dz, int32_t B, int32_t c_in, int32_t H, int32_t W,
                           int32_t c_out, float* d_dw, void* stream);

/* Number of kernel launches this library has enqueued in this process (bench.py's `gpu_launches`). */
unsigned long long im2im_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* IM2IM_UQ_H_ */
