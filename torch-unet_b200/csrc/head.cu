// Regression head after FC1's matmul, eval mode (replaces /root/reference/regression_model.py:37-46 and
// two_branch_regression.py:43-53,100): split-K reduce + bias/BN1d fold + LeakyReLU + Linear(f1,f2) + BN1d fold
// + LeakyReLU + Linear(f2,1) [+ Sigmoid * 0.5].  One CTA per image; everything stays fp32.  Tiny and
// latency-bound (256 images x 66 kFLOP), so the only goals are coalesced reads and a single launch.
#include "ctk_common.h"
#include "ctk_ptx.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kMaxF1 = 512;
constexpr int kMaxF2 = 128;

__global__ void __launch_bounds__(kThreads)
head_eval_kernel(const float* __restrict__ partial, int splits, int m_stride, int f1, int f2,
                 const float* __restrict__ scale1, const float* __restrict__ shift1, const float* __restrict__ w2,
                 const float* __restrict__ scale2, const float* __restrict__ shift2, const float* __restrict__ w3,
                 const float* __restrict__ b3, float slope, int sigmoid_half, float* __restrict__ out) {
  __shared__ __align__(16) float h1[kMaxF1];
  __shared__ float h2[kMaxF2];
  const int img = blockIdx.x;
  for (int f = threadIdx.x; f < f1; f += kThreads) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[(static_cast<size_t>(s) * m_stride + img) * f1 + f];
    h1[f] = ctk::leaky(fmaf(acc, scale1[f], shift1[f]), slope);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < f2; j += kThreads / 32) {
    const float4* wrow = reinterpret_cast<const float4*>(w2 + static_cast<size_t>(j) * f1);
    float acc = 0.f;
    for (int i = lane; i < f1 / 4; i += 32) {
      const float4 wv = __ldg(wrow + i);
      const float4 hv = *reinterpret_cast<const float4*>(&h1[4 * i]);
      acc = fmaf(wv.x, hv.x, acc); acc = fmaf(wv.y, hv.y, acc);
      acc = fmaf(wv.z, hv.z, acc); acc = fmaf(wv.w, hv.w, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) h2[j] = ctk::leaky(fmaf(acc, scale2[j], shift2[j]), slope);
  }
  __syncthreads();
  if (warp == 0) {
    float acc = 0.f;
    for (int j = lane; j < f2; j += 32) acc = fmaf(w3[j], h2[j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      float z = acc + b3[0];
      if (sigmoid_half) z = 0.5f / (1.f + expf(-z));
      out[img] = z;
    }
  }
}

}  // namespace

extern "C" int ctk_head_eval(const float* fc1_partial, int splits, int m_stride, int n, int f1, int f2,
                             const float* scale1, const float* shift1, const float* w2, const float* scale2,
                             const float* shift2, const float* w3, const float* b3, float slope, int sigmoid_half,
                             float* out, void* stream) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(fc1_partial && scale1 && shift1 && w2 && scale2 && shift2 && w3 && b3 && out);
  CTK_REQUIRE(n > 0 && splits > 0 && m_stride >= n && f1 > 0 && f1 <= kMaxF1 && f1 % 4 == 0 && f2 > 0 && f2 <= kMaxF2);
  CTK_REQUIRE((reinterpret_cast<uintptr_t>(w2) & 15) == 0);
  head_eval_kernel<<<n, kThreads, 0, ctk::as_stream(stream)>>>(fc1_partial, splits, m_stride, f1, f2, scale1, shift1,
                                                               w2, scale2, shift2, w3, b3, slope, sigmoid_half, out);
  return ctk::check_launch();
}
