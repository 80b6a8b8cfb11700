// Train-mode regression head (everything after FC1's big matmul) and its backward pass, fp32 on the CUDA cores:
//   Z1 = FC1 partials + b1 -> BN1d(batch stats) -> LeakyReLU -> Dropout -> Linear(512,128) -> BN1d -> LeakyReLU -> Dropout
//   -> Linear(128,1) [-> Sigmoid * 0.5]
// Replaces (training mode) /root/reference/regression_model.py:37-46 and two_branch_regression.py:43-53,100 plus their
// autograd.  The tensors are tiny ([N,512], [N,128]); the kernels are generic building blocks (column statistics,
// strided small GEMM, elementwise BN/activation/dropout forward and backward) launched a few times per step.
#include "ctk_common.h"
#include "ctk_ptx.cuh"

namespace {

using namespace ctk;

// Z[n][f] = sum_s in[s*split_stride + n*row_stride + f] + bias[f];  stats[f] = sum_n Z, stats[F+f] = sum_n Z^2
// One CTA per 32 columns, 32 row groups (1024 threads): a thread owns rows rg, rg + 32, ... of one column and adds the
// split-K partials of a row in split order with four loads in flight.  (With 8 row groups and one load at a time the FC1
// call -- 18 splits x 256 rows -- was a 39 us chain of dependent L2 reads on 16 SMs.)  Row-group totals are combined in a
// fixed order: deterministic.
constexpr int kColstatGroups = 32;
__global__ void __launch_bounds__(32 * kColstatGroups)
colstat_kernel(const float* __restrict__ in, int splits, long long split_stride, int row_stride,
               const float* __restrict__ bias, int n_rows, int F, float* __restrict__ z, float* __restrict__ stats) {
  __shared__ float r1[kColstatGroups][32], r2[kColstatGroups][32];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rg = threadIdx.x >> 5;
  float s1 = 0.f, s2 = 0.f;
  if (col < F) {
    const float b = bias ? __ldg(bias + col) : 0.f;
    for (int n = rg; n < n_rows; n += kColstatGroups) {
      const float* src = in + static_cast<long long>(n) * row_stride + col;
      float v = b;
      int s = 0;
      for (; s + 4 <= splits; s += 4) {
        float t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) t[u] = src[(s + u) * split_stride];
#pragma unroll
        for (int u = 0; u < 4; ++u) v += t[u];
      }
      for (; s < splits; ++s) v += src[s * split_stride];
      if (z) z[static_cast<long long>(n) * F + col] = v;
      s1 += v;
      s2 = fmaf(v, v, s2);
    }
  }
  r1[rg][threadIdx.x & 31] = s1;
  r2[rg][threadIdx.x & 31] = s2;
  __syncthreads();
  if (rg == 0 && col < F && stats) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < kColstatGroups; ++i) { a += r1[i][threadIdx.x]; b += r2[i][threadIdx.x]; }
    stats[col] = a;
    stats[F + col] = b;
  }
}

// A = dropout(leaky(z*scale + shift)):  mask is a 0/1 keep-mask (nullptr = keep all), keep_scale = 1/(1-p)
__global__ void bn1d_act_drop_fwd_kernel(const float* __restrict__ z, const float* __restrict__ scale,
                                         const float* __restrict__ shift, const float* __restrict__ mask,
                                         float keep_scale, float slope, int F, long long total, float* __restrict__ a) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int f = static_cast<int>(i % F);
  float v = leaky(fmaf(z[i], scale[f], shift[f]), slope);
  if (mask) v *= mask[i] * keep_scale;
  a[i] = v;
}

// Keep-masks of nn.Dropout drawn on the device: Philox-4x32-10 keyed by (seed), counter = (element index / 4, stream
// offset); element i is kept when its uniform draw in [0, 1) is >= p.  One launch covers both Dropout layers of the
// head (mask1 [n1] then mask2 [n2] elements).  Not PyTorch's generator stream: parity runs inject their masks instead.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
__global__ void dropout_masks_kernel(float* __restrict__ mask1, long long n1, float p1, float* __restrict__ mask2,
                                     long long n2, float p2, unsigned long long seed, unsigned long long offset) {
  const long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;      // one Philox block = 4 elements
  const long long q1 = (n1 + 3) / 4, q2 = (n2 + 3) / 4;
  if (q >= q1 + q2) return;
  uint32_t c[4] = {static_cast<uint32_t>(q), static_cast<uint32_t>(q >> 32), static_cast<uint32_t>(offset),
                   static_cast<uint32_t>(offset >> 32)};
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const bool second = q >= q1;
  float* dst = second ? mask2 : mask1;
  const long long base = (second ? q - q1 : q) * 4, n = second ? n2 : n1;
  const float p = second ? p2 : p1;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (base + j < n) dst[base + j] = (static_cast<float>(c[j] >> 8) * (1.f / 16777216.f)) >= p ? 1.f : 0.f;
}

// C[i][j] = sum_k A[i*a_i + k*a_k] * B[j*b_j + k*b_k] (+ bias[j]); 32x32 tiles, K staged through shared memory
__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const float* __restrict__ A, long long a_i, long long a_k, const float* __restrict__ B,
                     long long b_j, long long b_k, const float* __restrict__ bias, int M, int N, int K,
                     float* __restrict__ C, int ldc) {
  __shared__ float sa[32][33], sb[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // ty 0..7
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  // the next K tile is fetched into registers while the current one is multiplied (the loop used to be a chain of exposed
  // L2 round trips: 50 us for the 256 x 128 x 512 product)
  float pa[4], pb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = ty + 8 * q, i = i0 + r, j = j0 + r, k = k0 + tx;
      pa[q] = (i < M && k < K) ? A[i * a_i + k * a_k] : 0.f;
      pb[q] = (j < N && k < K) ? B[j * b_j + k * b_k] : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      sa[ty + 8 * q][tx] = pa[q];
      sb[ty + 8 * q][tx] = pb[q];
    }
    __syncthreads();
    if (k0 + 32 < K) fetch(k0 + 32);
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float bv = sb[tx][k];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(sa[ty + 8 * r][k], bv, acc[r]);
    }
    __syncthreads();
  }
  const int j = j0 + tx;
  if (j < N) {
    const float b = bias ? bias[j] : 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ty + 8 * r;
      if (i < M) C[static_cast<long long>(i) * ldc + j] = acc[r] + b;
    }
  }
}

// out[n] = a2[n,:] . w3 + b3 (optionally 0.5*sigmoid)
__global__ void head_out_fwd_kernel(const float* __restrict__ a2, const float* __restrict__ w3,
                                    const float* __restrict__ b3, int n_rows, int F, int sigmoid_half,
                                    float* __restrict__ out) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= n_rows) return;
  float acc = 0.f;
  for (int j = lane; j < F; j += 32) acc = fmaf(a2[static_cast<long long>(n) * F + j], w3[j], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    float z = acc + b3[0];
    if (sigmoid_half) z = 0.5f / (1.f + expf(-z));
    out[n] = z;
  }
}

// dz3[n] = dout[n] * d(out)/dz;  dA2[n][j] = dz3[n]*w3[j];  dw3[j] = sum_n dz3[n]*a2[n][j];  db3 = sum_n dz3[n]
// one CTA per 32 columns j; the 8 warps split the rows, their partial sums are added in warp order (deterministic)
__global__ void __launch_bounds__(256)
head_out_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ a2,
                    const float* __restrict__ w3, int n_rows, int F, int sigmoid_half, float* __restrict__ da2,
                    float* __restrict__ dw3, float* __restrict__ db3) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const float w = j < F ? w3[j] : 0.f;
  float acc = 0.f;
  for (int n = rg; n < n_rows; n += 8) {
    float g = dout[n];
    if (sigmoid_half) g *= out[n] * (1.f - 2.f * out[n]);      // d/dz 0.5*sigmoid(z) = out*(1-2*out)
    if (j < F) {
      const long long i = static_cast<long long>(n) * F + j;
      da2[i] = g * w;
      acc = fmaf(g, a2[i], acc);
    }
  }
  red[rg][lane] = acc;
  __syncthreads();
  if (rg == 0 && j < F) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[r][lane];
    dw3[j] = t;
  }
  if (blockIdx.x == 0 && rg == 1) {                            // db3: lane l adds rows l, l+32, ..., then a fixed butterfly
    float t = 0.f;
    for (int n = lane; n < n_rows; n += 32) {
      float g = dout[n];
      if (sigmoid_half) g *= out[n] * (1.f - 2.f * out[n]);
      t += g;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) db3[0] = t;
  }
}

// backward of dropout + LeakyReLU + BN1d, pass 1: dact = dA*mask*keep_scale*leaky'(zn); sums[f] = sum dact, sums[F+f] = sum dact*xhat
__global__ void __launch_bounds__(32 * kColstatGroups)      // 32 row groups like colstat_kernel: 8 rows per thread
bn1d_bwd_reduce_kernel(const float* __restrict__ da, const float* __restrict__ mask, float keep_scale,
                       const float* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                       const float* __restrict__ mean, const float* __restrict__ invstd, float slope, int n_rows, int F,
                       float* __restrict__ dact, float* __restrict__ sums) {
  __shared__ float r1[kColstatGroups][32], r2[kColstatGroups][32];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rg = threadIdx.x >> 5;
  float s1 = 0.f, s2 = 0.f;
  if (col < F) {
    const float sc = scale[col], sh = shift[col], mu = mean[col], is = invstd[col];
    for (int n = rg; n < n_rows; n += kColstatGroups) {
      const long long i = static_cast<long long>(n) * F + col;
      const float zn = fmaf(z[i], sc, sh);
      float g = da[i] * (zn > 0.f ? 1.f : slope);
      if (mask) g *= mask[i] * keep_scale;
      dact[i] = g;
      s1 += g;
      s2 = fmaf(g, (z[i] - mu) * is, s2);
    }
  }
  r1[rg][threadIdx.x & 31] = s1;
  r2[rg][threadIdx.x & 31] = s2;
  __syncthreads();
  if (rg == 0 && col < F) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < kColstatGroups; ++i) { a += r1[i][threadIdx.x]; b += r2[i][threadIdx.x]; }
    sums[col] = a;
    sums[F + col] = b;
  }
}

// pass 2: dZ = scale*(dact - s1/N - xhat*s2/N); optional bf16 copies dz_bf16[n][f] (row stride F) and dzT_bf16[f][n] (row stride ldt)
__global__ void bn1d_bwd_apply_kernel(const float* __restrict__ dact, const float* __restrict__ z,
                                      const float* __restrict__ scale, const float* __restrict__ mean,
                                      const float* __restrict__ invstd, const float* __restrict__ sums, int n_rows,
                                      int F, float* __restrict__ dz, __nv_bfloat16* __restrict__ dz_bf16,
                                      __nv_bfloat16* __restrict__ dzT_bf16, int ldt) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(n_rows) * F) return;
  const int f = static_cast<int>(i % F);
  const int n = static_cast<int>(i / F);
  const float inv_n = 1.f / static_cast<float>(n_rows);
  const float xhat = (z[i] - mean[f]) * invstd[f];
  const float g = scale[f] * (dact[i] - sums[f] * inv_n - xhat * sums[F + f] * inv_n);
  dz[i] = g;
  if (dz_bf16) dz_bf16[i] = __float2bfloat16_rn(g);
  if (dzT_bf16) dzT_bf16[static_cast<long long>(f) * ldt + n] = __float2bfloat16_rn(g);
}

}  // namespace

extern "C" {

int ctk_colstat(const float* in, int splits, long long split_stride, int row_stride, const float* bias, int n_rows,
                int features, float* z, float* stats, void* stream) {
  CTK_REQUIRE(in && n_rows > 0 && features > 0 && splits > 0 && (z || stats));
  colstat_kernel<<<(features + 31) / 32, 32 * kColstatGroups, 0, ctk::as_stream(stream)>>>(in, splits, split_stride, row_stride, bias,
                                                                           n_rows, features, z, stats);
  return ctk::check_launch();
}

int ctk_bn1d_act_drop_fwd(const float* z, const float* scale, const float* shift, const float* mask, float drop_p,
                          float slope, int n_rows, int features, float* a, void* stream) {
  CTK_REQUIRE(z && scale && shift && a && n_rows > 0 && features > 0 && drop_p >= 0.f && drop_p < 1.f);
  const long long total = static_cast<long long>(n_rows) * features;
  bn1d_act_drop_fwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, ctk::as_stream(stream)>>>(
      z, scale, shift, mask, 1.f / (1.f - drop_p), slope, features, total, a);
  return ctk::check_launch();
}

int ctk_dropout_masks(float* mask1, long long n1, float p1, float* mask2, long long n2, float p2, unsigned long long seed,
                      unsigned long long offset, void* stream) {
  CTK_REQUIRE((mask1 || n1 == 0) && (mask2 || n2 == 0) && n1 >= 0 && n2 >= 0 && p1 >= 0.f && p1 < 1.f && p2 >= 0.f &&
              p2 < 1.f);
  const long long blocks4 = (n1 + 3) / 4 + (n2 + 3) / 4;
  if (blocks4 == 0) return CTK_OK;
  dropout_masks_kernel<<<static_cast<unsigned>((blocks4 + 255) / 256), 256, 0, ctk::as_stream(stream)>>>(
      mask1, n1, p1, mask2, n2, p2, seed, offset);
  return ctk::check_launch();
}

int ctk_sgemm_strided(const float* a, long long a_i, long long a_k, const float* b, long long b_j, long long b_k,
                      const float* bias, int M, int N, int K, float* c, int ldc, void* stream) {
  CTK_REQUIRE(a && b && c && M > 0 && N > 0 && K > 0 && ldc >= N);
  dim3 grid((N + 31) / 32, (M + 31) / 32);
  sgemm_strided_kernel<<<grid, 256, 0, ctk::as_stream(stream)>>>(a, a_i, a_k, b, b_j, b_k, bias, M, N, K, c, ldc);
  return ctk::check_launch();
}

int ctk_head_out_fwd(const float* a2, const float* w3, const float* b3, int n_rows, int features, int sigmoid_half,
                     float* out, void* stream) {
  CTK_REQUIRE(a2 && w3 && b3 && out && n_rows > 0 && features > 0);
  head_out_fwd_kernel<<<(n_rows + 7) / 8, 256, 0, ctk::as_stream(stream)>>>(a2, w3, b3, n_rows, features, sigmoid_half,
                                                                            out);
  return ctk::check_launch();
}

int ctk_head_out_bwd(const float* dout, const float* out, const float* a2, const float* w3, int n_rows, int features,
                     int sigmoid_half, float* da2, float* dw3, float* db3, void* stream) {
  CTK_REQUIRE(dout && out && a2 && w3 && da2 && dw3 && db3 && n_rows > 0 && n_rows <= 8192 && features > 0);
  head_out_bwd_kernel<<<(features + 31) / 32, 256, 0, ctk::as_stream(stream)>>>(dout, out, a2, w3, n_rows, features,
                                                                               sigmoid_half, da2, dw3, db3);
  return ctk::check_launch();
}

int ctk_bn1d_bwd_reduce(const float* da, const float* mask, float drop_p, const float* z, const float* scale,
                        const float* shift, const float* mean, const float* invstd, float slope, int n_rows,
                        int features, float* dact, float* sums, void* stream) {
  CTK_REQUIRE(da && z && scale && shift && mean && invstd && dact && sums && n_rows > 0 && features > 0);
  bn1d_bwd_reduce_kernel<<<(features + 31) / 32, 32 * kColstatGroups, 0, ctk::as_stream(stream)>>>(
      da, mask, 1.f / (1.f - drop_p), z, scale, shift, mean, invstd, slope, n_rows, features, dact, sums);
  return ctk::check_launch();
}

int ctk_bn1d_bwd_apply(const float* dact, const float* z, const float* scale, const float* mean, const float* invstd,
                       const float* sums, int n_rows, int features, float* dz, void* dz_bf16, void* dzT_bf16, int ldt,
                       void* stream) {
  CTK_REQUIRE(dact && z && scale && mean && invstd && sums && dz && n_rows > 0 && features > 0);
  CTK_REQUIRE(dzT_bf16 == nullptr || ldt >= n_rows);
  const long long total = static_cast<long long>(n_rows) * features;
  bn1d_bwd_apply_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, ctk::as_stream(stream)>>>(
      dact, z, scale, mean, invstd, sums, n_rows, features, dz, static_cast<__nv_bfloat16*>(dz_bf16),
      static_cast<__nv_bfloat16*>(dzT_bf16), ldt);
  return ctk::check_launch();
}

}  // extern "C"
