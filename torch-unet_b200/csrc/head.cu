// Regression head after FC1's matmul, eval mode (replaces /root/reference/regression_model.py:37-46 and
// two_branch_regression.py:43-53,100): split-K reduce + bias/BN1d fold + LeakyReLU + Linear(f1,f2) + BN1d fold
// + LeakyReLU + Linear(f2,1) [+ Sigmoid * 0.5].  Everything stays fp32.  Tiny (256 images x 66 kFLOP) and latency-bound:
// a CTA takes kImg images, so every row of FC2's weight is fetched once per four images and the loads of four dot
// products are in flight together (one CTA per image spent 57 us walking 128 dependent row fetches); 512 threads keep the
// chain of split-K partial loads (18 per value, six in flight) to four values per thread.
// Per image the arithmetic order is the one-image-per-CTA kernel's: results are bit-identical to round 1's.
#include "ctk_common.h"
#include "ctk_ptx.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kImg = 4;
constexpr int kMaxF1 = 512;
constexpr int kMaxF2 = 128;

__global__ void __launch_bounds__(kThreads)
head_eval_kernel(const float* __restrict__ partial, int splits, int m_stride, int n, int f1, int f2,
                 const float* __restrict__ scale1, const float* __restrict__ shift1, const float* __restrict__ w2,
                 const float* __restrict__ scale2, const float* __restrict__ shift2, const float* __restrict__ w3,
                 const float* __restrict__ b3, float slope, int sigmoid_half, float* __restrict__ out) {
  __shared__ __align__(16) float h1[kImg][kMaxF1];
  __shared__ float h2[kImg][kMaxF2];
  const int img0 = blockIdx.x * kImg;
  const int nimg = n - img0 < kImg ? n - img0 : kImg;
  const size_t sstride = static_cast<size_t>(m_stride) * f1;
  // ---- split-K partial sums of FC1 (added in split order) -> folded bias / BatchNorm1d -> LeakyReLU
  for (int idx = threadIdx.x; idx < kImg * f1; idx += kThreads) {
    const int im = idx / f1, f = idx - im * f1;
    float v = 0.f;
    if (im < nimg) {
      const float* src = partial + static_cast<size_t>(img0 + im) * f1 + f;
      float acc = 0.f;
      int s = 0;
      for (; s + 6 <= splits; s += 6) {                    // six independent loads in flight, same order of additions
        float a[6];
#pragma unroll
        for (int u = 0; u < 6; ++u) a[u] = src[(s + u) * sstride];
#pragma unroll
        for (int u = 0; u < 6; ++u) acc += a[u];
      }
      for (; s < splits; ++s) acc += src[s * sstride];
      v = ctk::leaky(fmaf(acc, scale1[f], shift1[f]), slope);
    }
    h1[im][f] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- FC2: a warp per output row; the row's weights are loaded once and used for the CTA's four images
  for (int j = warp; j < f2; j += kThreads / 32) {
    const float4* wrow = reinterpret_cast<const float4*>(w2 + static_cast<size_t>(j) * f1);
    float acc[kImg];
#pragma unroll
    for (int im = 0; im < kImg; ++im) acc[im] = 0.f;
    for (int i = lane; i < f1 / 4; i += 32) {
      const float4 wv = __ldg(wrow + i);
#pragma unroll
      for (int im = 0; im < kImg; ++im) {
        const float4 hv = *reinterpret_cast<const float4*>(&h1[im][4 * i]);
        acc[im] = fmaf(wv.x, hv.x, acc[im]); acc[im] = fmaf(wv.y, hv.y, acc[im]);
        acc[im] = fmaf(wv.z, hv.z, acc[im]); acc[im] = fmaf(wv.w, hv.w, acc[im]);
      }
    }
#pragma unroll
    for (int im = 0; im < kImg; ++im) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[im] += __shfl_xor_sync(0xffffffffu, acc[im], o);
    }
    if (lane == 0) {
      const float sc = scale2[j], sh = shift2[j];
#pragma unroll
      for (int im = 0; im < kImg; ++im) h2[im][j] = ctk::leaky(fmaf(acc[im], sc, sh), slope);
    }
  }
  __syncthreads();
  // ---- FC3 (+ Sigmoid * 0.5): one warp per image
  if (warp < nimg) {
    float acc = 0.f;
    for (int j = lane; j < f2; j += 32) acc = fmaf(w3[j], h2[warp][j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      float z = acc + b3[0];
      if (sigmoid_half) z = 0.5f / (1.f + expf(-z));
      out[img0 + warp] = z;
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
  head_eval_kernel<<<(n + kImg - 1) / kImg, kThreads, 0, ctk::as_stream(stream)>>>(
      fc1_partial, splits, m_stride, n, f1, f2, scale1, shift1, w2, scale2, shift2, w3, b3, slope, sigmoid_half, out);
  return ctk::check_launch();
}
