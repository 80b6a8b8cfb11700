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

/* ------------------------------------------------------------------------------------------
 * Pearson r of channel 0 vs channel 1 of each [2,H,W] float32 tile.
 * Replaces: scipy.stats.pearsonr + the np.std()==0 guard, test-cross-talk-model.py:59-64.
 * r_out[i] is float64; NaN when either plane is constant; clipped to [-1, 1].
 * ------------------------------------------------------------------------------------------ */
size_t ctk_pearson_workspace_bytes(int n_tiles);
int ctk_pearson_f32(const float* tiles, int n_tiles, int plane_elems, double* r_out,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Parameter re-packing (derived cache; redo after every optimizer step / load_state_dict).
 * ------------------------------------------------------------------------------------------ */
/* Eval-mode BatchNorm folded with the conv/linear bias:  y = acc*scale + shift  where
 * scale = gamma/sqrt(rvar+eps), shift = (bias - rmean)*scale + beta.
 * Replaces: nn.BatchNorm2d/1d in eval(), regression_model.py:15,24,37,42; two_branch_regression.py:11,17,23,29,43,48. */
int ctk_fold_bn_eval(const float* bias, const float* gamma, const float* beta, const float* rmean,
                     const float* rvar, float eps, int channels, float* scale, float* shift, void* stream);
/* conv weight [Cout,Cin,3,3] fp32 -> [9][Cout][Cin] bf16 (tap-major, Cin contiguous = K-major GEMM B operand). */
int ctk_pack_conv_weight_bf16(const float* w, int cout, int cin, void* w_packed_bf16, void* stream);
/* first-layer weight [Cout,Cin,3,3] fp32 times per-channel scale -> fp32 [Cout][Cin*9] (BN folded into the taps). */
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
 * K % (64*splits) == 0.
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
 * MSE loss (mean reduction) and its gradient.  Replaces: torch.nn.MSELoss(), train_model.py:636,421.
 * loss_out: [1] fp32; grad_out (may be NULL): [n] fp32 = 2*(out-target)/n.
 * ------------------------------------------------------------------------------------------ */
int ctk_mse_loss(const float* out, const float* target, int n, float* loss_out, float* grad_out, void* stream);

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

#ifdef __cplusplus
}
#endif
#endif /* CTK_H_ */
