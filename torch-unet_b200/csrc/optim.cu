// MSE loss (+ gradient) and the multi-tensor Adam update.
//   ctk_mse_loss   replaces torch.nn.MSELoss(),            /root/reference/train_model.py:636,421
//   ctk_adam_multi replaces optim.Adam(lr, wd=1e-4).step(), /root/reference/train_model.py:637,424
// Adam is pure HBM streaming: 16 B read (p, g, m, v) + 12 B written (p, m, v) per parameter = 28 B/param;
// one launch covers every tensor of the model through a (tensor, chunk) table, 128-bit accesses.
#include "ctk_common.h"

#include <cmath>

namespace {

__global__ void mse_kernel(const float* __restrict__ out, const float* __restrict__ target, int n,
                           float* __restrict__ loss, float* __restrict__ grad) {
  __shared__ float sh[32];
  float acc = 0.f;
  const float inv_n = 1.f / static_cast<float>(n);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = out[i] - target[i];
    acc = fmaf(d, d, acc);
    if (grad) grad[i] = 2.f * d * inv_n;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) loss[0] = acc * inv_n;
  }
}

__global__ void scale_by_scalar_kernel(const float* __restrict__ in, const float* __restrict__ scalar, int n,
                                       float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * scalar[0];
}

struct AdamConsts {
  float beta1, beta2, one_minus_beta1, one_minus_beta2, eps, weight_decay, step_size, inv_sqrt_bc2, grad_scale;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamConsts& c) {
  g = fmaf(c.weight_decay, p, g * c.grad_scale);            // coupled L2
  m = fmaf(g - m, c.one_minus_beta1, m);                    // exp_avg.lerp_(g, 1-beta1)
  v = fmaf(c.one_minus_beta2 * g, g, v * c.beta2);          // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
  const float denom = fmaf(sqrtf(v), c.inv_sqrt_bc2, c.eps);
  p = fmaf(-c.step_size, m / denom, p);                     // param.addcdiv_(m, denom, -step_size)
}

__global__ void __launch_bounds__(256)
adam_multi_kernel(float* const* __restrict__ params, const float* const* __restrict__ grads,
                  float* const* __restrict__ exp_avg, float* const* __restrict__ exp_avg_sq,
                  const int64_t* __restrict__ numel, const int32_t* __restrict__ block_tensor,
                  const int32_t* __restrict__ block_chunk, AdamConsts c) {
  const int t = block_tensor[blockIdx.x];
  const int64_t begin = static_cast<int64_t>(block_chunk[blockIdx.x]) * CTK_ADAM_CHUNK;
  const int64_t n = numel[t];
  const int64_t end = begin + CTK_ADAM_CHUNK < n ? begin + CTK_ADAM_CHUNK : n;
  float* p = params[t];
  const float* g = grads[t];
  float* m = exp_avg[t];
  float* v = exp_avg_sq[t];
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                         reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  int64_t i = begin;
  if (aligned) {
    const int64_t nvec = (end - begin) / 4;
    for (int64_t k = threadIdx.x; k < nvec; k += blockDim.x) {
      const int64_t idx = begin + 4 * k;
      float4 pv = *reinterpret_cast<float4*>(p + idx);
      const float4 gv = __ldcs(reinterpret_cast<const float4*>(g + idx));
      float4 mv = *reinterpret_cast<float4*>(m + idx);
      float4 vv = *reinterpret_cast<float4*>(v + idx);
      adam_one(pv.x, gv.x, mv.x, vv.x, c);
      adam_one(pv.y, gv.y, mv.y, vv.y, c);
      adam_one(pv.z, gv.z, mv.z, vv.z, c);
      adam_one(pv.w, gv.w, mv.w, vv.w, c);
      *reinterpret_cast<float4*>(p + idx) = pv;
      *reinterpret_cast<float4*>(m + idx) = mv;
      *reinterpret_cast<float4*>(v + idx) = vv;
    }
    i = begin + 4 * nvec;
  }
  for (int64_t k = i + threadIdx.x; k < end; k += blockDim.x) {
    float pv = p[k], mv = m[k], vv = v[k];
    adam_one(pv, g[k], mv, vv, c);
    p[k] = pv; m[k] = mv; v[k] = vv;
  }
}

}  // namespace

extern "C" {

int ctk_mse_loss(const float* out, const float* target, int n, float* loss_out, float* grad_out, void* stream) {
  CTK_REQUIRE(out && target && loss_out && n > 0);
  mse_kernel<<<1, 256, 0, ctk::as_stream(stream)>>>(out, target, n, loss_out, grad_out);
  return ctk::check_launch();
}

int ctk_scale_by_scalar(const float* in, const float* scalar, int n, float* out, void* stream) {
  CTK_REQUIRE(in && scalar && out && n > 0);
  scale_by_scalar_kernel<<<(n + 255) / 256, 256, 0, ctk::as_stream(stream)>>>(in, scalar, n, out);
  return ctk::check_launch();
}

int ctk_adam_multi(void* const* params, void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                   const int64_t* numel, const int32_t* block_tensor, const int32_t* block_chunk, int n_blocks,
                   float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                   void* stream) {
  if (n_blocks == 0) return CTK_OK;
  CTK_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel && block_tensor && block_chunk && n_blocks > 0 &&
              step >= 1);
  const double bc1 = 1.0 - std::pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - std::pow(static_cast<double>(beta2), step);
  AdamConsts c;
  c.beta1 = beta1;
  c.beta2 = beta2;
  c.one_minus_beta1 = 1.f - beta1;
  c.one_minus_beta2 = 1.f - beta2;
  c.eps = eps;
  c.weight_decay = weight_decay;
  c.step_size = static_cast<float>(static_cast<double>(lr) / bc1);
  c.inv_sqrt_bc2 = static_cast<float>(1.0 / std::sqrt(bc2));
  c.grad_scale = grad_scale;
  adam_multi_kernel<<<n_blocks, 256, 0, ctk::as_stream(stream)>>>(
      reinterpret_cast<float* const*>(params), reinterpret_cast<const float* const*>(grads),
      reinterpret_cast<float* const*>(exp_avg), reinterpret_cast<float* const*>(exp_avg_sq), numel, block_tensor,
      block_chunk, c);
  return ctk::check_launch();
}

}  // extern "C"
