/* libctk -- C ABI of the B200-native CrosstalkPy hot path.
 *
 * The reference (djpbarry/Torch-Unet) has no FFI of its own: its hot path is a chain of
 * PyTorch library calls made by two nn.Modules and two scripts.  Each entry point below
 * replaces the library call(s) named in its comment (file:line into /root/reference) and is
 * what the Python glue in torch-unet_b200/ctk binds with ctypes.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer unless its name ends in _host; sizes are element
 *     counts unless the name says bytes;
 *   - no torch types, no C++ types, nothing thrown across the boundary: every function
 *     returns CTK_OK (0) or a negative ctk_status; ctk_status_string() explains it;
 *   - no allocation inside the library: outputs and workspaces are caller-owned, their
 *     sizes come from the matching *_workspace_bytes() query;
 *   - asynchronous on `stream` (a cudaStream_t passed as void*); re-entrant; one process per GPU;
 *   - activations between layers are NHWC bf16; parameters arrive in the reference's own
 *     state_dict layout (fp32, conv [Cout,Cin,3,3], fc [out,in] with NCHW-flatten columns)
 *     and are re-packed by the ctk_pack_* calls into a derived cache the caller owns.
 */
#ifndef CTK_H_
#define CTK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ctk_status {
  CTK_OK = 0,
  CTK_ERR_BAD_ARG = -1,      /* null pointer, unsupported shape, misaligned buffer           */
  CTK_ERR_WORKSPACE = -2,    /* workspace smaller than *_workspace_bytes() says              */
  CTK_ERR_CUDA = -3,         /* a CUDA runtime/driver call failed (see ctk_last_cuda_error) */
  CTK_ERR_NO_DEVICE = -4,    /* no sm_100 device is current                                  */
  CTK_ERR_UNSUPPORTED = -5   /* valid request this build does not implement                  */
} ctk_status;

int ctk_abi_version(void);
const char* ctk_status_string(int status);
/* Last CUDA error code seen by this thread inside libctk (0 if none). */
int ctk_last_cuda_error(void);
/* 0 if the current device is an sm_100 part the kernels can run on, else CTK_ERR_NO_DEVICE. */
int ctk_device_check(void);
/* Leave `sms` (0..64) streaming multiprocessors free when sizing the grids of the persistent one-CTA-per-SM tensor-core
 * kernels (ctk_conv3x3_tc_*, ctk_conv3x3_wgrad_tc) launched AFTER this call.  Data-parallel training sets it around the
 * backward pass: the NCCL all-reduce of the gradients (there is none in the single-process reference) then
 * finds free SMs at once instead of displacing CTAs of a full grid into a second wave.  Process-wide setting, 0 by default;
 * results stay deterministic for a fixed value (the partition of every reduction depends on the grid). */
int ctk_set_persistent_sm_reserve(int sms);

/* ------------------------------------------------------------------------------------------
 * Pearson r of channel 0 vs channel 1 of each [2,H,W] float32 tile.
 * Replaces: scipy.stats.pearsonr + the np.std()==0 guard, test-cross-talk-model.py:59-64.
 * r_out[i] is float64; NaN when either plane is constant; clipped to [-1, 1].
 * ------------------------------------------------------------------------------------------ */
size_t ctk_pearson_workspace_bytes(int n_tiles);
int ctk_pearson_f32(const float* tiles, int n_tiles, int plane_elems, double* r_out,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused per-tile comparison metrics (the next metrics of the same host loop, test-cross-talk-model.py:59-79):
 *   pearson_out[n]   f64  as ctk_pearson_f32                                                   (:59-64)
 *   rmse_out[n]      f32  np.sqrt(np.mean((img0 - img1) ** 2))                                  (:79)
 *   hist_out[n][2][256] u32  np.histogram(plane.flatten(), bins=256)[0] of both planes -- bit-exact counts (:65-66)
 *   hist_corr_out[n] f64  pearsonr(hist0, hist1), NaN if either histogram is flat               (:67-70)
 * Any output pointer may be NULL (hist_corr_out needs hist_out).  One read of the tiles from HBM plus one re-read
 * (L2 resident) for the histograms.
 * ------------------------------------------------------------------------------------------ */
size_t ctk_tile_metrics_workspace_bytes(int n_tiles);
int ctk_tile_metrics_f32(const float* tiles, int n_tiles, int plane_elems, double* pearson_out, float* rmse_out,
                         double* hist_corr_out, unsigned int* hist_out, void* workspace, size_t workspace_bytes,
                         void* stream);
/* nmi_out[n] f64 = sklearn.metrics.normalized_mutual_info_score(np.digitize(img0, linspace(min, max, 256)),
 * np.digitize(img1, ...)) -- test-cross-talk-model.py:71-74,84.  The 256 x 256 contingency table of every tile is built in
 * the workspace with the exact float32 digitisation rule of NumPy >= 2; MI and entropies in fp64. */
size_t ctk_tile_nmi_workspace_bytes(int n_tiles);
int ctk_tile_nmi_f32(const float* tiles, int n_tiles, int plane_elems, double* nmi_out, void* workspace,
                     size_t workspace_bytes, void* stream);
/* ssim_out[n] f64 = skimage.metrics.structural_similarity(img0, img1, data_range=max(both) - min(both)) with its defaults
 * (7x7 uniform window, K1 0.01, K2 0.03, sample covariance, float32 arithmetic, mean over the tile cropped by 3 pixels)
 * -- test-cross-talk-model.py:80-82.  tiles: [n][2][H][W] float32, H, W >= 7, H * W even, W <= 588, n <= 65535. */
size_t ctk_tile_ssim_workspace_bytes(int n_tiles);
int ctk_tile_ssim_f32(const float* tiles, int n_tiles, int H, int W, double* ssim_out, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Parameter re-packing (derived cache; redo after every optimizer step / load_state_dict).
 * ------------------------------------------------------------------------------------------ */
/* Eval-mode BatchNorm folded with the conv/linear bias:  y = acc*scale + shift  where
 * scale = gamma/sqrt(rvar+eps), shift = (bias - rmean)*scale + beta.
 * Replaces: nn.BatchNorm2d/1d in eval(), regression_model.py:15,24,37,42; two_branch_regression.py:11,17,23,29,43,48. */
int ctk_fold_bn_eval(const float* bias, const float* gamma, const float* beta, const float* rmean,
                     const float* rvar, float eps, int channels, float* scale, float* shift, void* stream);
/* conv weight [Cout,Cin,3,3] fp32 (nn.Conv2d.weight of regression_model.py:23, two_branch_regression.py:16,22,28)
 * -> [9][Cout][Cin] bf16 (tap-major, Cin contiguous = K-major GEMM B operand). */
int ctk_pack_conv_weight_bf16(const float* w, int cout, int cin, void* w_packed_bf16, void* stream);
/* first-layer weight [Cout,Cin,3,3] fp32 (regression_model.py:14, two_branch_regression.py:10) times per-channel scale
 * -> fp32 [Cout][Cin*9] (BN folded into the taps). */
int ctk_pack_first_weight(const float* w, const float* scale, int cout, int cin, float* w_folded, void* stream);
/* FC1 weight [out, C*HW] fp32 (columns in NCHW-flatten order c*HW+p, nn.Flatten of regression_model.py:35 /
 * two_branch_regression.py:41) -> [out, HW*C] bf16 (columns in NHWC order p*C+c). */
int ctk_pack_fc1_weight_bf16(const float* w, int out_features, int channels, int hw, void* w_packed_bf16, void* stream);

/* ------------------------------------------------------------------------------------------
 * First conv block, eval mode:  Conv2d(cin in {1,2}, cout, 3, 1, 1) -> BN(eval) -> LeakyReLU -> MaxPool2d(2,2).
 * Replaces: regression_model.py:14-17 (cin=2) and two_branch_regression.py:10-13 (cin=1, one plane of x,
 * the split at :88-89 is the c_offset argument).
 * x: [n, c_total, H, W] fp32 NCHW; reads channels [c_offset, c_offset+cin).
 * out: NHWC bf16 [n, H/2, W/2, out_cstride], writes channels [out_coffset, out_coffset+cout).
 * ------------------------------------------------------------------------------------------ */
int ctk_conv_first_eval(const float* x, int n, int c_total, int c_offset, int cin, int H, int W,
                        const float* w_folded, const float* shift, int cout, float slope,
                        void* out_bf16, int out_cstride, int out_coffset, void* stream);
/* Same block, additionally storing what the training backward needs to route gradients through the max-pool and the
 * LeakyReLU without recomputing the convolution: codes_u32 is [n, H/2, W/2, cout/8] uint32, 4 bits per pooled element
 * (channel c of a pixel in bits [4*(c%8), 4*(c%8)+4) of word c/8): bits 0-1 = arg-max position 2*dy+dx inside the 2x2
 * window (aten::max_pool2d_with_indices), bit 2 = 1 if the pre-activation at the arg-max is negative (LeakyReLU branch). */
int ctk_conv_first_pool_codes(const float* x, int n, int c_total, int c_offset, int cin, int H, int W,
                              const float* w_folded, const float* shift, int cout, float slope,
                              void* out_bf16, int out_cstride, int out_coffset, void* codes_u32, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tensor-core conv block, eval mode (implicit GEMM on tcgen05, accumulators in TMEM, operands by TMA):
 *   Conv2d(cin, cout, 3, 1, 1) -> BN(eval, folded) -> LeakyReLU(slope) -> MaxPool2d(2,2)
 * Replaces: regression_model.py:23-26 and two_branch_regression.py:16-19,22-25,28-31.
 * x: NHWC bf16 [n, H, W, cin] (dense), cin % 64 == 0, cout % 128 == 0, H and W even, W % 8 == 0.
 * out: NHWC bf16 [n, H/2, W/2, out_cstride], channels [out_coffset, out_coffset+cout) -- the channel
 * offset is how the two branches write the halves of the concatenated feature map
 * (torch.cat, two_branch_regression.py:96) without a copy.
 * flags: CTK_CONV_* bits below (0 = default).
 * ------------------------------------------------------------------------------------------ */
#define CTK_CONV_NO_POOL 1       /* skip the 2x2 max-pool (out is [n,H,W,out_cstride])                 */
#define CTK_CONV_NO_ACT 2        /* skip LeakyReLU                                                     */
#define CTK_CONV_SINGLE_CTA 4    /* one CTA per tile (cta_group::1, N=128) instead of CTA pairs (cta_group::2) */
int ctk_conv3x3_tc_eval(const void* x_bf16, int n, int H, int W, int cin,
                        const void* w_packed_bf16, int cout, const float* scale, const float* shift, float slope,
                        void* out_bf16, int out_cstride, int out_coffset, int flags, void* stream);

/* ------------------------------------------------------------------------------------------
 * Split-K bf16 GEMM on tcgen05:  partial[s][m][n] = sum_{k in split s} A[m][k] * B[n][k]   (fp32 out)
 * Replaces: the first nn.Linear of each head (aten::addmm), regression_model.py:36 and
 * two_branch_regression.py:42; the bias is added by ctk_head_eval, which also sums the splits.
 * A: [M,K] bf16 row-major, B: [N,K] bf16 row-major, M % 128 == 0 (pad the batch), N % 128 == 0,
 * K % 64 == 0, 1 <= splits <= K / 64.  The K / 64 blocks are dealt out contiguously and as evenly as they go: split s covers
 * blocks [s*q + min(s, r), ...) with q = (K/64) / splits, r = (K/64) % splits, the first r splits one block more.
 * ------------------------------------------------------------------------------------------ */
int ctk_gemm_bf16_splitk(const void* a_bf16, const void* b_bf16, int M, int N, int K, int splits,
                         float* partial, void* stream);

/* ------------------------------------------------------------------------------------------
 * Regression head, eval mode, everything after FC1's matmul:
 *   (+bias1) -> BN1d(eval) -> LeakyReLU -> [Dropout = identity] -> Linear(f1,f2) -> BN1d(eval) -> LeakyReLU
 *   -> Linear(f2,1) [-> Sigmoid -> *0.5 when sigmoid_half != 0]
 * Replaces: regression_model.py:37-46,58-61 and two_branch_regression.py:43-53,100.
 * fc1_partial: [splits][m_stride][f1] fp32 from ctk_gemm_bf16_splitk; scale1/shift1 fold bias1 and BN;
 * w2: [f2][f1] fp32; scale2/shift2 fold bias2 and BN; w3: [f2] fp32; b3: [1]; out: [n] fp32.
 * ------------------------------------------------------------------------------------------ */
int ctk_head_eval(const float* fc1_partial, int splits, int m_stride, int n, int f1, int f2,
                  const float* scale1, const float* shift1, const float* w2, const float* scale2,
                  const float* shift2, const float* w3, const float* b3, float slope, int sigmoid_half,
                  float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * fp32-class inference ("precision = fp32", north_star: score within 1e-5 of the fp32 reference).
 * The tensor cores only multiply bf16, so every fp32 operand is carried as a pair of bf16 numbers, v = hi + lo with
 * hi = bf16(v), lo = bf16(v - hi) (16 significant bits), and every product is three MMAs,
 *     x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo        (dropped x_lo*w_lo ~ 2^-17 relative),
 * accumulated in fp32 in TMEM.  Activations between blocks are NHWC bf16 tensors with 2C channels [hi(C) | lo(C)]
 * (written through the out_hi / out_lo pointers, which share pixel stride and channel offset), conv weights are
 * [9][Cout][3*Cin] = [w_hi | w_hi | w_lo] so that the SAME implicit-GEMM pipeline runs with K = 3*Cin over
 * [x_hi | x_lo | x_hi]; BatchNorm, max-pool and LeakyReLU run on the fp32 accumulators.  FC1 is three split-K GEMMs
 * (feat_hi*W_hi, feat_lo*W_hi, feat_hi*W_lo) whose partial sums ctk_head_eval adds up (splits = 3 * splits).
 * Same reference lines as the bf16 entry points above.
 * ------------------------------------------------------------------------------------------ */
int ctk_pack_conv_weight_split_bf16(const float* w, int cout, int cin, void* w_split_bf16, void* stream);
int ctk_pack_fc1_weight_split_bf16(const float* w, int out_features, int channels, int hw, void* w_hi_bf16,
                                   void* w_lo_bf16, void* stream);
int ctk_conv_first_eval_split(const float* x, int n, int c_total, int c_offset, int cin, int H, int W,
                              const float* w_folded, const float* shift, int cout, float slope,
                              void* out_hi_bf16, void* out_lo_bf16, int out_cstride, int out_coffset, void* stream);
/* x_split: [n,H,W,2*cin] bf16 = [hi | lo]; cin = logical input channels (multiple of 64), cout % 128 == 0. */
int ctk_conv3x3_tc_eval_split(const void* x_split_bf16, int n, int H, int W, int cin, const void* w_split_bf16,
                              int cout, const float* scale, const float* shift, float slope, void* out_hi_bf16,
                              void* out_lo_bf16, int out_cstride, int out_coffset, void* stream);

/* ------------------------------------------------------------------------------------------
 * Input pipeline on the device (the caller side of the path):
 *   raw pixel payloads [n][2][H*W] (float64 as the reference's TIFFs store them, or float32) ->
 *   astype(float32) -> per-plane (img - min) / (max - min), constant planes unchanged -> optional flips -> out [n,2,H,W] f32
 * Replaces: train_model.py:166-167 (imread(...).astype(np.float32)), :211-216 (normalize_image), :225-232 (TF.hflip /
 * TF.vflip on both planes).  flip_flags: NULL or [n] bytes, bit 0 = horizontal flip, bit 1 = vertical flip (the caller
 * draws them: `torch.rand(1) < 0.5` twice per sample in the reference).  Bit-identical to the NumPy float32 arithmetic.
 * ------------------------------------------------------------------------------------------ */
int ctk_prepare_tiles(const void* raw, int raw_is_f64, const unsigned char* flip_flags, int n, int H, int W, float* out,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * MSE loss (mean reduction) and its gradient.  Replaces: torch.nn.MSELoss(), train_model.py:636,421.
 * loss_out: [1] fp32; grad_out (may be NULL): [n] fp32 = 2*(out-target)/n.
 * ------------------------------------------------------------------------------------------ */
int ctk_mse_loss(const float* out, const float* target, int n, float* loss_out, float* grad_out, void* stream);
/* out[i] = in[i] * scalar[0] (scalar on the device): the backward of the MSE loss scales the gradient ctk_mse_loss
 * already produced by the incoming d(loss), which loss.backward() seeds with 1 (train_model.py:422). */
int ctk_scale_by_scalar(const float* in, const float* scalar, int n, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-tensor Adam with coupled L2 (torch.optim.Adam(lr, weight_decay) semantics).
 * Replaces: optim.Adam(...).step(), train_model.py:637,424.
 * The four pointer tables (param, grad, exp_avg, exp_avg_sq: n_tensors device pointers each) and
 * `numel` (n_tensors int64) live in DEVICE memory; block_tensor / block_chunk (n_blocks int32 each, device)
 * map each thread block to (tensor, chunk of CTK_ADAM_CHUNK elements).  grad_scale multiplies the
 * gradient first (1/world_size after a sum-allreduce).
 * ------------------------------------------------------------------------------------------ */
#define CTK_ADAM_CHUNK 65536
int ctk_adam_multi(void* const* params, void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                   const int64_t* numel, const int32_t* block_tensor, const int32_t* block_chunk, int n_blocks,
                   float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                   void* stream);

/* ==========================================================================================
 * Training path (train-mode BatchNorm, backward pass).  Replaces loss.backward() (train_model.py:422, a10 in
 * SURVEY 8a) and the train-mode forward of the layers named at each entry.
 * ========================================================================================== */

/* Raw convolutions for train mode: y = conv(x) WITHOUT bias (train-mode BatchNorm cancels it; it re-enters the
 * running mean in ctk_bn_finalize), stored bf16 NHWC at full resolution, plus stats[c] = sum(y), stats[C+c] = sum(y^2)
 * over all pixels (fp32 accumulators; stats may be NULL for ctk_conv3x3_tc_raw, which is also the dgrad kernel when
 * fed ctk_pack_conv_weight_dgrad_bf16 weights: dX = conv(dY, rot180(W)^T)).
 * Replaces: nn.Conv2d forward / aten::convolution_backward (input gradient), regression_model.py:14,23;
 * two_branch_regression.py:10,16,22,28. */
size_t ctk_conv_first_raw_workspace_bytes(int cout);
int ctk_conv_first_raw(const float* x, int n, int c_total, int c_offset, int cin, int H, int W, const float* w,
                       int cout, void* y_bf16, float* stats, void* workspace, size_t workspace_bytes, void* stream);
/* Statistics are deterministic: each CTA stores one row of partial sums in the workspace (needed only when stats != NULL:
 * ctk_*_raw_workspace_bytes(cout) bytes, 16-byte aligned) and the rows are added in a fixed order -- no float atomics. */
size_t ctk_conv3x3_tc_raw_workspace_bytes(int cout);
int ctk_conv3x3_tc_raw(const void* x_bf16, int n, int H, int W, int cin, const void* w_packed_bf16, int cout,
                       void* y_bf16, float* stats, void* workspace, size_t workspace_bytes, void* stream);
/* conv weight [Cout,Cin,3,3] fp32 -> [9][Cin][Cout] bf16 with the taps rotated by 180 degrees (dgrad operand). */
int ctk_pack_conv_weight_dgrad_bf16(const float* w, int cout, int cin, void* w_packed_bf16, void* stream);
/* Both copies (ctk_pack_conv_weight_bf16 and ..._dgrad_bf16 layouts) of up to 8 conv weights in one launch: what a training
 * step needs of every tensor-core conv layer (nn.Conv2d weights, regression_model.py:23; two_branch_regression.py:16,22,28).
 * w / cout / cin / w_fwd_bf16 / w_dgrad_bf16 are HOST arrays of n_layers entries (device pointers inside). */
int ctk_pack_conv_weights_train(int n_layers, const float* const* w, const int* cout, const int* cin,
                                void* const* w_fwd_bf16, void* const* w_dgrad_bf16, void* stream);

/* Batch statistics -> normalisation constants; updates running_mean/var (momentum, unbiased variance, conv bias added
 * to the mean) and num_batches_tracked like nn.BatchNorm in train().  running_* / num_batches_tracked may be NULL.
 * Replaces: aten::native_batch_norm (training=True), regression_model.py:15,24,37,42. */
int ctk_bn_finalize(const float* sums, double count, const float* bias, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps,
                    int channels, float* scale, float* shift, float* mean, float* invstd, void* stream);
/* Same, from per-channel [mean, biased variance] instead of sums (first block: moments come from the patch Gram matrix). */
int ctk_bn_finalize_moments(const float* moments, double count, const float* bias, const float* gamma, const float* beta,
                            float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                            float eps, int channels, float* scale, float* shift, float* mean, float* invstd,
                            void* stream);

/* First conv block in training without materialising its full-resolution output.  With T = 9*cin taps,
 * gram = [S (T doubles) | G (T*T doubles)], S[t] = sum_p x[p+t], G[t][t'] = sum_p x[p+t] x[p+t'] (zero padded):
 *   ctk_first_patch_gram     gram of the input planes (overwrites gram)
 *   ctk_first_moments        moments[c] = batch mean, moments[cout+c] = biased batch variance of conv(x, w)[.,c] (no bias)
 *   (then ctk_bn_finalize_moments, ctk_pack_first_weight(w, scale) and ctk_conv_first_pool_codes give the pooled output)
 *   ctk_first_wgrad_codes    t1[c][t] = sum_windows g[w,c] * x[argmax(w,c) + t] and sums[c] = sum_windows g[w,c] (= d beta),
 *                            g = dP * f'(z*), arg-max position and sign of z* from the codes ctk_conv_first_pool_codes
 *                            stored in the forward pass; dp is dense bf16 NHWC
 *   ctk_first_wgrad_finalize sums[cout+c] = sum dA*xhat (= d gamma) = invstd_c (w_c . t1_c - mean_c sums[c]), then
 *                            dw[c][t] = scale_c (t1 - m1_c S_t - m2_c invstd_c ((G w_c)[t] - mean_c S_t)),
 *                            m1 = sums[c]/count, m2 = sums[cout+c]/count, count = n*H*W
 * Replaces (train mode): nn.Conv2d + nn.BatchNorm2d statistics and their backward for regression_model.py:14-15,
 * two_branch_regression.py:10-11. */
/* (Gram matrix and gather are two-stage and deterministic: per-CTA partial sums in the workspace -- *_workspace_bytes()
 * bytes, 16-byte aligned -- added in a fixed order; no floating-point atomics.) */
size_t ctk_first_patch_gram_workspace_bytes(int cin);
int ctk_first_patch_gram(const float* x, int n, int c_total, int c_offset, int cin, int H, int W, double* gram,
                         void* workspace, size_t workspace_bytes, void* stream);
int ctk_first_moments(const double* gram, const float* w, int cout, int cin, double count, float* moments, void* stream);
size_t ctk_first_wgrad_codes_workspace_bytes(int cin, int cout);
int ctk_first_wgrad_codes(const float* x, int n, int c_total, int c_offset, int cin, int H, int W, const void* codes_u32,
                          const void* dp_bf16, int cout, float slope, float* t1, float* sums, void* workspace,
                          size_t workspace_bytes, void* stream);
int ctk_first_wgrad_finalize(const float* t1, const double* gram, const float* w, const float* scale, const float* mean,
                             const float* invstd, float* sums, double count, int cout, int cin, float* dw,
                             void* stream);

/* out = maxpool2x2(leaky(y*scale + shift)), y bf16 NHWC [n,H,W,C] -> out bf16 NHWC [n,H/2,W/2,out_cstride] @ out_coffset.
 * Replaces: BatchNorm2d(train) apply + LeakyReLU + MaxPool2d, regression_model.py:15-17,24-26. */
int ctk_bn_act_pool_fwd(const void* y_bf16, int n, int H, int W, int channels, const float* scale, const float* shift,
                        float slope, void* out_bf16, int out_cstride, int out_coffset, void* stream);
/* Backward of MaxPool2d + LeakyReLU + BatchNorm2d(train).  dp = gradient of the pooled output (bf16 NHWC with channel
 * stride/offset).  reduce: sums[c] = sum(dA) (= dbeta), sums[C+c] = sum(dA*xhat) (= dgamma).  apply: dy (bf16 NHWC,
 * dense) = gamma*invstd*(dA - mean(dA) - xhat*mean(dA*xhat)).  The pool's argmax is recomputed (first maximum wins).
 * Replaces: the autograd backward (train_model.py:422) of regression_model.py:15-17,24-26 and
 * two_branch_regression.py:11-13,17-19,23-25,29-31. */
/* All three reductions are two-stage and deterministic: every CTA stores one row of partial sums into the workspace
 * (ctk_bn_bwd_reduce_workspace_bytes(channels) bytes, 16-byte aligned) and the rows are added in a fixed order with fp64
 * accumulation -- no floating-point atomics, identical launches give bit-identical sums. */
size_t ctk_bn_bwd_reduce_workspace_bytes(int channels);
int ctk_bn_bwd_reduce(const void* y_bf16, const void* dp_bf16, int dp_cstride, int dp_coffset, int n, int H, int W,
                      int channels, const float* scale, const float* shift, const float* mean, const float* invstd,
                      float slope, float* sums, void* workspace, size_t workspace_bytes, void* stream);
/* Same sums from the pooled activation P (the block's stored output) and dP alone: dA lives only at each window's
 * argmax, where the activation equals P, so f'(P) and xhat(P) = (leaky^-1(P) - beta)/gamma are recoverable.  Reads
 * 1 B per element of Y instead of 2.5 B.  One bf16 ulp of P moves xhat by ~|beta/gamma| 2^-8 and gamma == 0 makes it
 * unrecoverable (dgamma = 0 then): use ctk_bn_bwd_reduce_guarded unless the parameters are known to be benign. */
int ctk_bn_bwd_reduce_pooled(const void* pooled_bf16, int p_cstride, int p_coffset, const void* dp_bf16, int dp_cstride,
                             int dp_coffset, long long pooled_pixels, int channels, const float* gamma,
                             const float* beta, float slope, float* sums, void* workspace, size_t workspace_bytes,
                             void* stream);
/* The pooled reduction with a per-channel-group guard evaluated on the device (no host read of the parameters): a group of
 * 8 channels containing gamma == 0 or |beta| > 8 |gamma| is reduced from the raw conv output y like ctk_bn_bwd_reduce,
 * every other group from the pooled tensors.  This is the entry point the training path uses. */
int ctk_bn_bwd_reduce_guarded(const void* y_bf16, int n, int H, int W, const float* scale, const float* shift,
                              const float* mean, const float* invstd, const void* pooled_bf16, int p_cstride,
                              int p_coffset, const void* dp_bf16, int dp_cstride, int dp_coffset, int channels,
                              const float* gamma, const float* beta, float slope, float* sums, void* workspace,
                              size_t workspace_bytes, void* stream);
int ctk_bn_bwd_apply(const void* y_bf16, const void* dp_bf16, int dp_cstride, int dp_coffset, int n, int H, int W,
                     int channels, const float* scale, const float* shift, const float* mean, const float* invstd,
                     const float* sums, float slope, void* dy_bf16, void* stream);

/* Weight gradients.  dw is fp32 in the reference layout [Cout,Cin,3,3] and is overwritten.
 * ctk_conv3x3_wgrad_tc: tcgen05 GEMM over pixels with MN-major NHWC operands (cin % 64 == 0, cout % 128 == 0).
 * ctk_conv_first_wgrad: first layer (cin 1 or 2), x = fp32 NCHW input planes.
 * Replaces: aten::convolution_backward (weight gradient) of the nn.Conv2d layers at regression_model.py:14,23 and
 * two_branch_regression.py:10,16,22,28, reached through loss.backward() (train_model.py:422). */
/* Both are two-stage and deterministic: every CTA stores its partial sums in the workspace (*_workspace_bytes(cin, cout)
 * bytes, 16-byte aligned) and a second kernel adds them in a fixed order -- no floating-point atomics. */
size_t ctk_conv3x3_wgrad_tc_workspace_bytes(int cin, int cout);
int ctk_conv3x3_wgrad_tc(const void* dy_bf16, const void* x_bf16, int n, int H, int W, int cin, int cout, float* dw,
                         void* workspace, size_t workspace_bytes, void* stream);
size_t ctk_conv_first_wgrad_workspace_bytes(int cin, int cout);
int ctk_conv_first_wgrad(const void* dy_bf16, const float* x, int n, int c_total, int c_offset, int cin, int H, int W,
                         int cout, float* dw, void* workspace, size_t workspace_bytes, void* stream);

/* FC1 in training: feature-map transpose out[(c*HW+p)][n] = feat[n][p][c] (zero padded to ld columns), FC1 weight
 * transposed + permuted w_t[p*C+c][o] = w[o][c*HW+p], and a plain bf16-output GEMM C[M,N] = A[M,K] * B[N,K]^T
 * (M, N multiples of 128, K of 64).  With ctk_gemm_bf16_splitk they give FC1 forward, dX and dW:
 *   dfeat[n, p*C+c] = dZ1[n,:] . w_t[p*C+c,:]          dW1[o, c*HW+p] = dZ1^T[o,:] . featT[c*HW+p,:]
 * Replaces: aten::addmm / mm of nn.Linear, regression_model.py:36; two_branch_regression.py:42. */
int ctk_feat_transpose_bf16(const void* feat_bf16, int n, int hw, int channels, void* out_bf16, int ld, void* stream);
int ctk_pack_fc1_weight_t_bf16(const float* w, int out_features, int channels, int hw, void* w_t_bf16, void* stream);
int ctk_gemm_bf16_out_bf16(const void* a_bf16, const void* b_bf16, int M, int N, int K, void* c_bf16, void* stream);
/* same product with B given as [K][N] (N contiguous): FC1's dX reads the forward pass's packed weight [f1][HW*C] directly */
int ctk_gemm_bf16_bt_out_bf16(const void* a_bf16, const void* b_kn_bf16, int M, int N, int K, void* c_bf16, void* stream);

/* Small fp32 building blocks of the train-mode head and its backward (regression_model.py:37-46,
 * two_branch_regression.py:43-53,100).
 *  ctk_colstat:          z[n][f] = sum_s in[s*split_stride + n*row_stride + f] + bias[f]; stats = column sum / sum of squares
 *  ctk_bn1d_act_drop_fwd: a = leaky(z*scale+shift) * mask/(1-p)   (mask = 0/1 keep mask or NULL)
 *  ctk_sgemm_strided:    c[i][j] = sum_k a[i*a_i + k*a_k] * b[j*b_j + k*b_k] + bias[j]
 *  ctk_head_out_fwd/bwd: the last Linear(f,1) [+ Sigmoid*0.5] and its gradients
 *  ctk_bn1d_bwd_reduce/apply: backward of Dropout + LeakyReLU + BatchNorm1d(train); apply can also emit bf16 copies
 *                        dz_bf16[n][f] and dzT_bf16[f][n] (row stride ldt) for the FC1 backward GEMMs. */
int ctk_colstat(const float* in, int splits, long long split_stride, int row_stride, const float* bias, int n_rows,
                int features, float* z, float* stats, void* stream);
int ctk_bn1d_act_drop_fwd(const float* z, const float* scale, const float* shift, const float* mask, float drop_p,
                          float slope, int n_rows, int features, float* a, void* stream);
/* Keep-masks (1.0 keep / 0.0 drop) of the head's two nn.Dropout layers (regression_model.py:39,44;
 * two_branch_regression.py:45,50) drawn on the device in one launch: Philox-4x32-10 keyed by `seed`, counter =
 * (element / 4, offset); element kept when its uniform draw >= p.  Callers advance `offset` every step.  This is not
 * torch's generator stream -- parity runs hand their own masks to ctk_bn1d_act_drop_fwd instead. */
int ctk_dropout_masks(float* mask1, long long n1, float p1, float* mask2, long long n2, float p2, unsigned long long seed,
                      unsigned long long offset, void* stream);
int ctk_sgemm_strided(const float* a, long long a_i, long long a_k, const float* b, long long b_j, long long b_k,
                      const float* bias, int M, int N, int K, float* c, int ldc, void* stream);
int ctk_head_out_fwd(const float* a2, const float* w3, const float* b3, int n_rows, int features, int sigmoid_half,
                     float* out, void* stream);
int ctk_head_out_bwd(const float* dout, const float* out, const float* a2, const float* w3, int n_rows, int features,
                     int sigmoid_half, float* da2, float* dw3, float* db3, void* stream);
int ctk_bn1d_bwd_reduce(const float* da, const float* mask, float drop_p, const float* z, const float* scale,
                        const float* shift, const float* mean, const float* invstd, float slope, int n_rows,
                        int features, float* dact, float* sums, void* stream);
int ctk_bn1d_bwd_apply(const float* dact, const float* z, const float* scale, const float* mean, const float* invstd,
                       const float* sums, int n_rows, int features, float* dz, void* dz_bf16, void* dzT_bf16, int ldt,
                       void* stream);

/* ==========================================================================================
 * fp32 training path (ctk.set_precision(model, "fp32")): the reference's own arithmetic -- float32 operands, products
 * and stored tensors (train_model.py:419-424 runs the nn.Modules in fp32) -- on the CUDA cores, for parity runs.  FFMA
 * products, fp32 partial sums over 16 terms folded into fp64 running sums, fp64 batch statistics, fixed-order two-stage
 * cross-CTA reductions (bit-reproducible).  Activations are NHWC float32; `sn, sy, sx, sc` are ELEMENT strides of the
 * input read as [n][y][x][c], so the reference's NCHW input planes are consumed in place.
 * ========================================================================================== */
/* [Cout,Cin,3,3] -> [(tap*cin + ci)][cout] (rotate 0: forward operand) or [((8-tap)*cout + co)][cin] (rotate 1: the
 * operand that makes ctk_conv3x3_f32 compute the input gradient from dY). */
int ctk_pack_conv_weight_f32(const float* w, int cout, int cin, int rotate, float* out, void* stream);
/* y[n,H,W,cout] = conv3x3(x, w) without bias, stride 1, zero padding 1.  cout % 64 == 0.
 * Replaces: nn.Conv2d forward / input gradient, regression_model.py:14,23; two_branch_regression.py:10,16,22,28. */
int ctk_conv3x3_f32(const float* x, long long sn, long long sy, long long sx, long long sc, int n, int H, int W, int cin,
                    const float* w_packed, int cout, float* y, void* stream);
/* dw[Cout,Cin,3,3] = sum_pixels dY[p,co] * x[p+tap,ci]; split over pixel slices with fp64 partial sums in the workspace,
 * added in slice order.  cout % 64 == 0.  Replaces: aten::convolution_backward (weight gradient). */
size_t ctk_conv3x3_wgrad_f32_workspace_bytes(int n, int H, int W, int cin, int cout);
int ctk_conv3x3_wgrad_f32(const float* dy, const float* x, long long sn, long long sy, long long sx, long long sc, int n,
                          int H, int W, int cin, int cout, float* dw, void* workspace, size_t workspace_bytes,
                          void* stream);
/* sums[c] = sum_p y[p,c], sums[C+c] = sum_p y[p,c]^2 in fp64 (workspace: ctk_channel_sums_f64_workspace_bytes(C)), and
 * the BatchNorm constants from them (ctk_bn_finalize with fp64 sums).  Replaces: aten::native_batch_norm(training). */
size_t ctk_channel_sums_f64_workspace_bytes(int channels);
int ctk_channel_stats_f32(const float* y, long long pixels, int channels, double* sums, void* workspace,
                          size_t workspace_bytes, void* stream);
int ctk_bn_finalize_f64(const double* sums, double count, const float* bias, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                        float eps, int channels, float* scale, float* shift, float* mean, float* invstd, void* stream);
/* out[img*out_sn + (py*Wp+px)*out_sp + c*out_sc] = maxpool2x2(leaky((y-mean)*invstd*gamma + beta)): NHWC output with
 * (Hp*Wp*C, C, 1), or the NCHW-flatten order nn.Flatten hands to FC1 with (features, 1, Hp*Wp) (two_branch_regression.py:96). */
int ctk_bn_act_pool_fwd_f32(const float* y, int n, int H, int W, int channels, const float* mean, const float* invstd,
                            const float* gamma, const float* beta, float slope, float* out, long long out_sn,
                            long long out_sp, long long out_sc, void* stream);
/* Backward of MaxPool2d + LeakyReLU + BatchNorm2d(train) in fp32: sums (fp64; optional float copy = [d beta | d gamma]),
 * then the dense dY.  dp is read through the same three strides as the forward output. */
int ctk_bn_bwd_reduce_f32(const float* y, const float* dp, long long dp_sn, long long dp_sp, long long dp_sc, int n, int H,
                          int W, int channels, const float* mean, const float* invstd, const float* gamma,
                          const float* beta, float slope, double* sums, float* sums_f32, void* workspace,
                          size_t workspace_bytes, void* stream);
int ctk_bn_bwd_apply_f32(const float* y, const float* dp, long long dp_sn, long long dp_sp, long long dp_sc, int n, int H,
                         int W, int channels, const float* mean, const float* invstd, const float* gamma,
                         const float* beta, const double* sums, double count, float slope, float* dy, void* stream);
/* c[i*ldc + j] = sum_k a[i*a_i + k*a_k] * b[j*b_j + k*b_k] + bias[j], fp64 running sums: FC1 forward, dX and dW in the
 * reference's own layouts (regression_model.py:36; two_branch_regression.py:42). */
int ctk_gemm_f32(const float* a, long long a_i, long long a_k, const float* b, long long b_j, long long b_k,
                 const float* bias, int M, int N, int K, float* c, long long ldc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTK_H_ */
