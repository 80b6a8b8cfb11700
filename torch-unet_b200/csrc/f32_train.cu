// The fp32 training path: every operand, product and stored tensor is IEEE float32 -- the reference's own arithmetic
// (train_model.py:419-424 runs the nn.Modules in fp32) -- on the CUDA cores, for parity runs.
//
// The tensor-core path keeps operands in bf16 (2^-9 per element) and tcgen05 accumulates with truncation (about -2^-24
// relative per MMA step: profiles/r2_tcgen05_accumulation_bias.txt); with train-mode BatchNorm amplifying a 1e-6 relative
// input perturbation to 6e-4 .. 5e-3 of the whole gradient (measured in pure fp32 on the CPU), neither can reproduce the
// reference's gradients to 1e-4.  These kernels can: FFMA products, fp32 partial sums over 16 terms folded into fp64
// running sums, fp64 batch statistics and BatchNorm-backward sums, every cross-CTA reduction a fixed-order second stage
// (no atomics: bit-reproducible).  They are plain tiled implicit GEMMs -- 5-15 TFLOP/s, a twentieth of the tcgen05 path
// and still ~50x the reference's CPU step -- selected with ctk.set_precision(model, "fp32").
//
//   ctk_conv3x3_f32            nn.Conv2d forward, and its input gradient when fed rotated weights   (regression_model.py:14,23;
//   ctk_conv3x3_wgrad_f32      weight gradient                                                       two_branch_regression.py:10-28)
//   ctk_pack_conv_weight_f32   [Cout,Cin,3,3] -> [(tap, cin)][cout] (forward) or [(8 - tap, cout)][cin] (input gradient)
//   ctk_channel_stats_f32      per-channel sum / sum of squares in fp64                              (nn.BatchNorm2d, train mode)
//   ctk_bn_finalize_f64        statistics -> scale / shift / mean / invstd, running statistics
//   ctk_bn_act_pool_fwd_f32    normalise + LeakyReLU + MaxPool2d(2,2), output with free strides (NHWC, or the NCHW-flatten
//                              order nn.Flatten feeds to FC1)
//   ctk_bn_bwd_reduce_f32 / ctk_bn_bwd_apply_f32   backward of the three
//   ctk_gemm_f32               C = A B^T with free strides, fp64 running sums (FC1 forward, dX, dW)
#include "ctk_common.h"

#include <algorithm>

namespace {

using namespace ctk;

struct TensorView {           // element strides of an [n][y][x][c] tensor (NHWC, or NCHW planes read in place)
  long long sn, sy, sx, sc;
};

__device__ __forceinline__ float leaky_f(float v, float slope) { return v > 0.f ? v : v * slope; }

// ---------------------------------------------------------------------------------------------------------------------
// Y[p][co] = sum_{tap, ci} X[p + tap][ci] * Wk[(tap * cin + ci)][co]      (3x3, stride 1, zero padding 1, no bias)
// CTA: 128 consecutive pixels x 64 output channels, K in chunks of 16; thread: 8 pixels x 4 channels.
constexpr int kBM = 128, kBN = 64, kBK = 16, kAPitch = kBM + 2;

__global__ void __launch_bounds__(256)
conv3x3_f32_kernel(const float* __restrict__ x, TensorView xv, long long pixels, int H, int W, int cin,
                   const float* __restrict__ wk, int cout, float* __restrict__ y) {
  __shared__ float As[kBK][kAPitch];
  __shared__ __align__(16) float Bs[kBK][kBN];
  const int tid = threadIdx.x;
  const int ktot = 9 * cin;
  const long long p0 = static_cast<long long>(blockIdx.x) * kBM;
  const int n0 = blockIdx.y * kBN;
  // A loads: lane = k within the chunk (consecutive input channels are contiguous), 8 pixels per thread
  const int kl = tid & 15, mq = tid >> 4;
  long long pix_off[8];
  int ph[8], pw[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long p = p0 + mq + 16 * j;
    const int w = static_cast<int>(p % W);
    const long long t = p / W;
    const int h = static_cast<int>(t % H);
    ph[j] = p < pixels ? h : -4; pw[j] = w;                 // rows past the end (ragged last tile) read as padding
    pix_off[j] = (t / H) * xv.sn + h * xv.sy + w * xv.sx;
  }
  const int nl = tid & 63, kq = tid >> 6;                 // B loads
  const int tm = tid & 15, tn = tid >> 4;                 // compute: pixels i*16 + tm, channels tn*4 .. tn*4+3
  float acc[8][4];
  double sum[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; sum[i][j] = 0.0; }
  for (int k0 = 0; k0 < ktot; k0 += kBK) {
    const int kk = k0 + kl;
    const bool kin = kk < ktot;
    const int tap = kin ? kk / cin : 0;
    const int ci = kk - tap * cin;
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    const long long koff = dy * xv.sy + dx * xv.sx + ci * xv.sc;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int hh = ph[j] + dy, ww = pw[j] + dx;
      const bool ok = kin && hh >= 0 && hh < H && ww >= 0 && ww < W;
      As[kl][mq + 16 * j] = ok ? __ldg(x + pix_off[j] + koff) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kq + 4 * j;
      Bs[k][nl] = (k0 + k < ktot) ? __ldg(wk + static_cast<long long>(k0 + k) * cout + n0 + nl) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a = As[k][i * 16 + tm];
        acc[i][0] = fmaf(a, b.x, acc[i][0]);
        acc[i][1] = fmaf(a, b.y, acc[i][1]);
        acc[i][2] = fmaf(a, b.z, acc[i][2]);
        acc[i][3] = fmaf(a, b.w, acc[i][3]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { sum[i][j] += static_cast<double>(acc[i][j]); acc[i][j] = 0.f; }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long p = p0 + i * 16 + tm;
    if (p < pixels)
      *reinterpret_cast<float4*>(y + p * cout + n0 + tn * 4) =
        make_float4(static_cast<float>(sum[i][0]), static_cast<float>(sum[i][1]), static_cast<float>(sum[i][2]),
                    static_cast<float>(sum[i][3]));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// dW partial[s][co][nn] = sum_{p in slice s} dY[p][co] * X[p + tap(nn)][ci(nn)],  nn = tap * cin + ci
// CTA: 64 co x 64 nn, K = pixels in chunks of 16 (one image row segment); thread: 4 co x 4 nn; fp64 running sums.
__global__ void __launch_bounds__(256)
conv3x3_wgrad_f32_kernel(const float* __restrict__ dy, const float* __restrict__ x, TensorView xv, long long pixels, int H,
                         int W, int cin, int cout, long long chunks_total, int chunks_per_slice,
                         double* __restrict__ part) {
  __shared__ __align__(16) float As[kBK][64];
  __shared__ __align__(16) float Bs[kBK][64];
  const int tid = threadIdx.x;
  const int ntot = 9 * cin;
  const int co0 = blockIdx.x * 64, nn0 = blockIdx.y * 64;
  const int cl = tid & 63, pq = tid >> 6;
  const int nn = nn0 + cl;
  const bool nin = nn < ntot;
  const int tap = nin ? nn / cin : 0;
  const int ci = nn - tap * cin;
  const int ddy = tap / 3 - 1, ddx = tap % 3 - 1;
  const int tm = tid & 15, tn = tid >> 4;
  float acc[4][4];
  double sum[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; sum[i][j] = 0.0; }
  const long long c_begin = static_cast<long long>(blockIdx.z) * chunks_per_slice;
  const long long c_end = min(c_begin + chunks_per_slice, chunks_total);
  for (long long ch = c_begin; ch < c_end; ++ch) {
    const long long p0 = ch * kBK;                       // 16 consecutive pixels (they may span rows when W < 16)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pk = pq + 4 * j;
      const long long p = p0 + pk;
      const bool pin = p < pixels;
      As[pk][cl] = pin ? __ldg(dy + p * cout + co0 + cl) : 0.f;
      const int w = static_cast<int>(p % W);
      const long long t = p / W;
      const int h = static_cast<int>(t % H);
      const long long img = t / H;
      const int hh = h + ddy, ww = w + ddx;
      const bool ok = pin && nin && hh >= 0 && hh < H && ww >= 0 && ww < W;
      Bs[pk][cl] = ok ? __ldg(x + img * xv.sn + hh * xv.sy + ww * xv.sx + ci * xv.sc) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { sum[i][j] += static_cast<double>(acc[i][j]); acc[i][j] = 0.f; }
  }
  double* dst = part + static_cast<long long>(blockIdx.z) * cout * ntot;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tm * 4 + i, n = nn0 + tn * 4 + j;
      if (n < ntot) dst[static_cast<long long>(co) * ntot + n] = sum[i][j];
    }
}

// dw[co][ci][tap] = sum over slices (in slice order) of part[s][co][tap * cin + ci]
__global__ void wgrad_f32_reduce_kernel(const double* __restrict__ part, int slices, int cout, int cin,
                                        float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int ntot = 9 * cin;
  if (i >= cout * ntot) return;
  double s = 0.0;
  for (int k = 0; k < slices; ++k) s += part[static_cast<long long>(k) * cout * ntot + i];
  const int co = i / ntot, nn = i - co * ntot;
  const int tap = nn / cin, ci = nn - tap * cin;
  dw[(static_cast<long long>(co) * cin + ci) * 9 + tap] = static_cast<float>(s);
}

// rotate == 0: out[(tap * cin + ci) * cout + co] = w[co][ci][tap]              (forward operand)
// rotate == 1: out[(tap * cout + co) * cin + ci] = w[co][ci][8 - tap]          (input-gradient operand: conv of dY)
__global__ void pack_conv_weight_f32_kernel(const float* __restrict__ w, int cout, int cin, int rotate,
                                            float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * cin * 9) return;
  const int tap = i % 9, ci = (i / 9) % cin, co = i / (9 * cin);
  if (rotate) out[(static_cast<long long>(8 - tap) * cout + co) * cin + ci] = w[i];
  else out[(static_cast<long long>(tap) * cin + ci) * cout + co] = w[i];
}

// ---------------------------------------------------------------------------------------------------------------------
// per-channel fp64 partial sums over pixels: part[cta][2C] (sum, sum of squares)
template <bool kSquares>
__device__ __forceinline__ void cta_channel_partials(const double (&s1)[2], const double (&s2)[2], int C, int cq, int slots,
                                                     double* red, double* __restrict__ row) {
  // red: [256][4] doubles; threads of one channel (different pixel slots) are added in slot order
  const int tid = threadIdx.x;
  red[tid * 4 + 0] = s1[0]; red[tid * 4 + 1] = s1[1]; red[tid * 4 + 2] = s2[0]; red[tid * 4 + 3] = s2[1];
  __syncthreads();
  const int c_l = tid % cq, slot = tid / cq;
  if (slot == 0) {
    for (int r = 0; r < (C + cq - 1) / cq; ++r) {
      const int c = c_l + r * cq;
      if (c >= C) break;
      double a = 0.0, b = 0.0;
      for (int s = 0; s < slots; ++s) { a += red[(s * cq + c_l) * 4 + r]; b += red[(s * cq + c_l) * 4 + 2 + r]; }
      row[c] = a;
      row[C + c] = b;
    }
  }
}

__global__ void __launch_bounds__(256)
channel_stats_f32_kernel(const float* __restrict__ y, long long pixels, int C, double* __restrict__ part) {
  __shared__ double red[256 * 4];
  const int cq = C < 256 ? C : 256, slots = 256 / cq;
  const int c_l = threadIdx.x % cq, slot = threadIdx.x / cq;
  double s1[2] = {0.0, 0.0}, s2[2] = {0.0, 0.0};
  for (long long p = static_cast<long long>(blockIdx.x) * slots + slot; p < pixels; p += static_cast<long long>(gridDim.x) * slots) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int c = c_l + r * cq;
      if (c < C) {
        const double v = static_cast<double>(__ldg(y + p * C + c));
        s1[r] += v;
        s2[r] += v * v;
      }
    }
  }
  cta_channel_partials<true>(s1, s2, C, cq, slots, red, part + static_cast<long long>(blockIdx.x) * 2 * C);
}

__global__ void bn_finalize_f64_kernel(const double* __restrict__ sums, double count, const float* __restrict__ bias,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float* __restrict__ running_mean, float* __restrict__ running_var,
                                       long long* __restrict__ num_batches_tracked, float momentum, float eps, int c,
                                       float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                       float* __restrict__ invstd_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && num_batches_tracked) num_batches_tracked[0] += 1;
  if (i >= c) return;
  const double m = sums[i] / count;
  double var = sums[c + i] / count - m * m;                   // biased variance normalises the batch
  var = var > 0.0 ? var : 0.0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float sc = gamma[i] * invstd;
  scale[i] = sc;
  shift[i] = beta[i] - static_cast<float>(m) * sc;
  mean_out[i] = static_cast<float>(m);
  invstd_out[i] = invstd;
  if (running_mean) {
    const float b = bias ? bias[i] : 0.f;                     // the raw conv output excludes the conv bias
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * (static_cast<float>(m) + b);
    running_var[i] = (1.f - momentum) * running_var[i] + momentum * static_cast<float>(unbiased);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward: out = maxpool2x2(leaky((y - mean) * invstd * gamma + beta)), computed as nn.BatchNorm2d does (not as a folded
// scale / shift), thread = (pooled pixel, channel)
__global__ void __launch_bounds__(256)
bn_act_pool_fwd_f32_kernel(const float* __restrict__ y, int H, int W, int C, const float* __restrict__ mean,
                           const float* __restrict__ invstd, const float* __restrict__ gamma,
                           const float* __restrict__ beta, float slope, float* __restrict__ out, long long on,
                           long long op, long long oc, long long total) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int Hp = H >> 1, Wp = W >> 1;
  const int c = static_cast<int>(i % C);
  const long long pp = i / C;
  const int px = static_cast<int>(pp % Wp);
  const long long t = pp / Wp;
  const int py = static_cast<int>(t % Hp);
  const long long img = t / Hp;
  const float mu = __ldg(mean + c), is = __ldg(invstd + c), g = __ldg(gamma + c), b = __ldg(beta + c);
  const float* src = y + ((img * H + 2 * py) * W + 2 * px) * C + c;
  float best = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float v = __ldg(src + ((j >> 1) * W + (j & 1)) * static_cast<long long>(C));
    const float z = leaky_f((v - mu) * is * g + b, slope);
    best = j == 0 ? z : fmaxf(best, z);
  }
  out[img * on + (static_cast<long long>(py) * Wp + px) * op + c * oc] = best;
}

// the four pre-activations of a window and the index of the first maximum of the activations
__device__ __forceinline__ int window_argmax_f32(const float* __restrict__ src, int W, int C, float mu, float is, float g,
                                                 float b, float slope, float (&yv)[4], float (&z)[4]) {
  int arg = 0;
  float best = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    yv[j] = __ldg(src + ((j >> 1) * W + (j & 1)) * static_cast<long long>(C));
    z[j] = (yv[j] - mu) * is * g + b;
    const float a = leaky_f(z[j], slope);
    if (j == 0 || a > best) { best = a; arg = j; }           // strictly greater: the first maximum wins
  }
  return arg;
}

// backward pass 1: per channel sum(dA) and sum(dA * xhat) in fp64, dA = dP * f'(z) at the window's argmax
__global__ void __launch_bounds__(256)
bn_bwd_reduce_f32_kernel(const float* __restrict__ y, const float* __restrict__ dp, long long dn, long long dpp,
                         long long dc, int H, int W, int C, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float slope, long long pooled_pixels,
                         double* __restrict__ part) {
  __shared__ double red[256 * 4];
  const int Hp = H >> 1, Wp = W >> 1;
  const int cq = C < 256 ? C : 256, slots = 256 / cq;
  const int c_l = threadIdx.x % cq, slot = threadIdx.x / cq;
  double s1[2] = {0.0, 0.0}, s2[2] = {0.0, 0.0};
  for (long long pp = static_cast<long long>(blockIdx.x) * slots + slot; pp < pooled_pixels;
       pp += static_cast<long long>(gridDim.x) * slots) {
    const int px = static_cast<int>(pp % Wp);
    const long long t = pp / Wp;
    const int py = static_cast<int>(t % Hp);
    const long long img = t / Hp;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int c = c_l + r * cq;
      if (c < C) {
        const float mu = __ldg(mean + c), is = __ldg(invstd + c);
        float yv[4], z[4];
        const int arg = window_argmax_f32(y + ((img * H + 2 * py) * W + 2 * px) * C + c, W, C, mu, is, __ldg(gamma + c),
                                          __ldg(beta + c), slope, yv, z);
        const float g = __ldg(dp + img * dn + (static_cast<long long>(py) * Wp + px) * dpp + c * dc);
        const float da = z[arg] > 0.f ? g : g * slope;
        s1[r] += static_cast<double>(da);
        s2[r] += static_cast<double>(da) * static_cast<double>((yv[arg] - mu) * is);
      }
    }
  }
  cta_channel_partials<true>(s1, s2, C, cq, slots, red, part + static_cast<long long>(blockIdx.x) * 2 * C);
}

// backward pass 2: dY = gamma * invstd * (dA - mean(dA) - xhat * mean(dA * xhat)), dense, thread = (pooled pixel, channel)
__global__ void __launch_bounds__(256)
bn_bwd_apply_f32_kernel(const float* __restrict__ y, const float* __restrict__ dp, long long dn, long long dpp,
                        long long dc, int H, int W, int C, const float* __restrict__ mean,
                        const float* __restrict__ invstd, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const double* __restrict__ sums, double inv_count, float slope,
                        float* __restrict__ dy, long long total) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int Hp = H >> 1, Wp = W >> 1;
  const int c = static_cast<int>(i % C);
  const long long pp = i / C;
  const int px = static_cast<int>(pp % Wp);
  const long long t = pp / Wp;
  const int py = static_cast<int>(t % Hp);
  const long long img = t / Hp;
  const float mu = __ldg(mean + c), is = __ldg(invstd + c), g = __ldg(gamma + c);
  const float m1 = static_cast<float>(sums[c] * inv_count), m2 = static_cast<float>(sums[C + c] * inv_count);
  const long long base = ((img * H + 2 * py) * W + 2 * px) * C + c;
  float yv[4], z[4];
  const int arg = window_argmax_f32(y + base, W, C, mu, is, g, __ldg(beta + c), slope, yv, z);
  const float gp = __ldg(dp + img * dn + (static_cast<long long>(py) * Wp + px) * dpp + c * dc);
  const float da_top = z[arg] > 0.f ? gp : gp * slope;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float xhat = (yv[j] - mu) * is;
    const float da = j == arg ? da_top : 0.f;
    dy[base + ((j >> 1) * W + (j & 1)) * static_cast<long long>(C)] = g * is * (da - m1 - xhat * m2);
  }
}

__global__ void f64_to_f32_kernel(const double* __restrict__ in, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = static_cast<float>(in[i]);
}

// ---------------------------------------------------------------------------------------------------------------------
// C[i][j] = sum_k A[i*a_i + k*a_k] * B[j*b_j + k*b_k] (+ bias[j]); 64 x 64 tiles, K chunks of 16, fp64 running sums
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, long long a_i, long long a_k, const float* __restrict__ B, long long b_j,
                long long b_k, const float* __restrict__ bias, int M, int N, int K, float* __restrict__ C, long long ldc) {
  __shared__ __align__(16) float As[kBK][64 + 4];
  __shared__ __align__(16) float Bs[kBK][64 + 4];
  const int tid = threadIdx.x;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  // loads: when the K stride is 1 lanes run along k, else along the row index, so that either layout is coalesced
  const bool a_kfast = a_k == 1, b_kfast = b_k == 1;
  const int tm = tid & 15, tn = tid >> 4;
  float acc[4][4];
  double sum[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; sum[i][j] = 0.0; }
  for (int k0 = 0; k0 < K; k0 += kBK) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int e = tid + 256 * s;                        // 1024 elements of each 16 x 64 tile
      {
        const int k = a_kfast ? (e & 15) : (e >> 6), r = a_kfast ? (e >> 4) : (e & 63);
        const int i = i0 + r;
        As[k][r] = (i < M && k0 + k < K) ? __ldg(A + i * a_i + (k0 + k) * a_k) : 0.f;
      }
      {
        const int k = b_kfast ? (e & 15) : (e >> 6), r = b_kfast ? (e >> 4) : (e & 63);
        const int j = j0 + r;
        Bs[k][r] = (j < N && k0 + k < K) ? __ldg(B + j * b_j + (k0 + k) * b_k) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { sum[i][j] += static_cast<double>(acc[i][j]); acc[i][j] = 0.f; }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = i0 + tm * 4 + i, c = j0 + tn * 4 + j;
      if (r < M && c < N) C[r * ldc + c] = static_cast<float>(sum[i][j] + (bias ? static_cast<double>(bias[c]) : 0.0));
    }
}

inline int stats_grid(long long items, int C) {
  const int cq = C < 256 ? C : 256, slots = 256 / cq;
  const long long blocks = (items + slots - 1) / slots;
  const long long cap = static_cast<long long>(ctk::num_sms()) * 4;
  return static_cast<int>(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace

extern "C" {

int ctk_pack_conv_weight_f32(const float* w, int cout, int cin, int rotate, float* out, void* stream) {
  CTK_REQUIRE(w && out && cout > 0 && cin > 0);
  const int total = cout * cin * 9;
  pack_conv_weight_f32_kernel<<<(total + 255) / 256, 256, 0, ctk::as_stream(stream)>>>(w, cout, cin, rotate, out);
  return ctk::check_launch();
}

int ctk_conv3x3_f32(const float* x, long long sn, long long sy, long long sx, long long sc, int n, int H, int W, int cin,
                    const float* w_packed, int cout, float* y, void* stream) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(x && w_packed && y && n > 0 && H > 0 && W > 0 && cin > 0 && cout > 0 && cout % kBN == 0);
  const long long pixels = static_cast<long long>(n) * H * W;
  CTK_REQUIRE((pixels + kBM - 1) / kBM < (1ll << 31) && (reinterpret_cast<uintptr_t>(y) & 15) == 0);
  const TensorView xv = {sn, sy, sx, sc};
  conv3x3_f32_kernel<<<dim3(static_cast<unsigned>((pixels + kBM - 1) / kBM), cout / kBN), 256, 0, ctk::as_stream(stream)>>>(
      x, xv, pixels, H, W, cin, w_packed, cout, y);
  return ctk::check_launch();
}

static int wgrad_f32_slices(long long chunks, int cout, int cin) {
  const long long tiles = static_cast<long long>(cout / 64) * ((9 * cin + 63) / 64);
  long long s = (static_cast<long long>(ctk::num_sms()) * 4 + tiles - 1) / tiles;
  if (s > chunks) s = chunks;
  return static_cast<int>(s < 1 ? 1 : s);
}

size_t ctk_conv3x3_wgrad_f32_workspace_bytes(int n, int H, int W, int cin, int cout) {
  if (n <= 0 || H <= 0 || W <= 0 || cin <= 0 || cout <= 0) return 0;
  const long long chunks = (static_cast<long long>(n) * H * W + kBK - 1) / kBK;
  return static_cast<size_t>(wgrad_f32_slices(chunks, cout, cin)) * cout * 9 * cin * sizeof(double);
}

int ctk_conv3x3_wgrad_f32(const float* dy, const float* x, long long sn, long long sy, long long sx, long long sc, int n,
                          int H, int W, int cin, int cout, float* dw, void* workspace, size_t workspace_bytes,
                          void* stream) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(dy && x && dw && n > 0 && H > 0 && W > 0 && cin > 0 && cout > 0 && cout % 64 == 0);
  const long long pixels = static_cast<long long>(n) * H * W;
  const long long chunks = (pixels + kBK - 1) / kBK;
  const int slices = wgrad_f32_slices(chunks, cout, cin);
  CTK_REQUIRE_WORKSPACE(workspace, workspace_bytes, static_cast<size_t>(slices) * cout * 9 * cin * sizeof(double));
  const int per = static_cast<int>((chunks + slices - 1) / slices);
  const TensorView xv = {sn, sy, sx, sc};
  cudaStream_t s = ctk::as_stream(stream);
  double* part = static_cast<double*>(workspace);
  conv3x3_wgrad_f32_kernel<<<dim3(cout / 64, (9 * cin + 63) / 64, slices), 256, 0, s>>>(dy, x, xv, pixels, H, W, cin, cout,
                                                                                      chunks, per, part);
  int st = ctk::check_launch();
  if (st != CTK_OK) return st;
  const int total = cout * 9 * cin;
  wgrad_f32_reduce_kernel<<<(total + 255) / 256, 256, 0, s>>>(part, slices, cout, cin, dw);
  return ctk::check_launch();
}

size_t ctk_channel_sums_f64_workspace_bytes(int channels) {
  return channels > 0 ? static_cast<size_t>(ctk::num_sms()) * 4 * 2 * channels * sizeof(double) : 0;
}

int ctk_channel_stats_f32(const float* y, long long pixels, int channels, double* sums, void* workspace,
                          size_t workspace_bytes, void* stream) {
  CTK_REQUIRE(y && sums && pixels > 0 && channels > 0 && channels <= 512 && (channels >= 256 ? channels % 256 == 0 : 256 % channels == 0));
  const int grid = stats_grid(pixels, channels);
  CTK_REQUIRE_WORKSPACE(workspace, workspace_bytes, static_cast<size_t>(grid) * 2 * channels * sizeof(double));
  cudaStream_t s = ctk::as_stream(stream);
  double* part = static_cast<double*>(workspace);
  channel_stats_f32_kernel<<<grid, 256, 0, s>>>(y, pixels, channels, part);
  int st = ctk::check_launch();
  if (st != CTK_OK) return st;
  return ctk::reduce_rows_f64(part, grid, 2 * channels, 2 * channels, sums, s);
}

int ctk_bn_finalize_f64(const double* sums, double count, const float* bias, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                        float eps, int channels, float* scale, float* shift, float* mean, float* invstd, void* stream) {
  CTK_REQUIRE(sums && gamma && beta && scale && shift && mean && invstd && channels > 0 && count >= 1.0);
  CTK_REQUIRE((running_mean == nullptr) == (running_var == nullptr));
  bn_finalize_f64_kernel<<<(channels + 127) / 128, 128, 0, ctk::as_stream(stream)>>>(
      sums, count, bias, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, channels, scale, shift,
      mean, invstd);
  return ctk::check_launch();
}

int ctk_bn_act_pool_fwd_f32(const float* y, int n, int H, int W, int channels, const float* mean, const float* invstd,
                            const float* gamma, const float* beta, float slope, float* out, long long out_sn,
                            long long out_sp, long long out_sc, void* stream) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(y && mean && invstd && gamma && beta && out && n > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && channels > 0);
  const long long total = static_cast<long long>(n) * (H / 2) * (W / 2) * channels;
  bn_act_pool_fwd_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, ctk::as_stream(stream)>>>(
      y, H, W, channels, mean, invstd, gamma, beta, slope, out, out_sn, out_sp, out_sc, total);
  return ctk::check_launch();
}

int ctk_bn_bwd_reduce_f32(const float* y, const float* dp, long long dp_sn, long long dp_sp, long long dp_sc, int n, int H,
                          int W, int channels, const float* mean, const float* invstd, const float* gamma,
                          const float* beta, float slope, double* sums, float* sums_f32, void* workspace,
                          size_t workspace_bytes, void* stream) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(y && dp && mean && invstd && gamma && beta && sums && n > 0 && H % 2 == 0 && W % 2 == 0);
  CTK_REQUIRE(channels > 0 && channels <= 512 && (channels >= 256 ? channels % 256 == 0 : 256 % channels == 0));
  const long long pooled = static_cast<long long>(n) * (H / 2) * (W / 2);
  const int grid = stats_grid(pooled, channels);
  CTK_REQUIRE_WORKSPACE(workspace, workspace_bytes, static_cast<size_t>(grid) * 2 * channels * sizeof(double));
  cudaStream_t s = ctk::as_stream(stream);
  double* part = static_cast<double*>(workspace);
  bn_bwd_reduce_f32_kernel<<<grid, 256, 0, s>>>(y, dp, dp_sn, dp_sp, dp_sc, H, W, channels, mean, invstd, gamma, beta, slope,
                                                pooled, part);
  int st = ctk::check_launch();
  if (st != CTK_OK) return st;
  st = ctk::reduce_rows_f64(part, grid, 2 * channels, 2 * channels, sums, s);
  if (st != CTK_OK || sums_f32 == nullptr) return st;
  f64_to_f32_kernel<<<(2 * channels + 255) / 256, 256, 0, s>>>(sums, 2 * channels, sums_f32);
  return ctk::check_launch();
}

int ctk_bn_bwd_apply_f32(const float* y, const float* dp, long long dp_sn, long long dp_sp, long long dp_sc, int n, int H,
                         int W, int channels, const float* mean, const float* invstd, const float* gamma,
                         const float* beta, const double* sums, double count, float slope, float* dy, void* stream) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(y && dp && mean && invstd && gamma && beta && sums && dy && n > 0 && H % 2 == 0 && W % 2 == 0 && channels > 0 &&
              count >= 1.0);
  const long long total = static_cast<long long>(n) * (H / 2) * (W / 2) * channels;
  bn_bwd_apply_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, ctk::as_stream(stream)>>>(
      y, dp, dp_sn, dp_sp, dp_sc, H, W, channels, mean, invstd, gamma, beta, sums, 1.0 / count, slope, dy, total);
  return ctk::check_launch();
}

int ctk_gemm_f32(const float* a, long long a_i, long long a_k, const float* b, long long b_j, long long b_k,
                 const float* bias, int M, int N, int K, float* c, long long ldc, void* stream) {
  CTK_REQUIRE(a && b && c && M > 0 && N > 0 && K > 0 && ldc >= N);
  gemm_f32_kernel<<<dim3((N + 63) / 64, (M + 63) / 64), 256, 0, ctk::as_stream(stream)>>>(a, a_i, a_k, b, b_j, b_k, bias, M, N,
                                                                                         K, c, ldc);
  return ctk::check_launch();
}

}  // extern "C"
