// Train-mode BatchNorm2d + LeakyReLU + MaxPool2d(2,2) around the tensor-core convolutions: statistics finalisation,
// the normalise/activate/pool pass, and the two backward passes (per-channel reductions, then the dense gradient of the
// raw conv output).  Replaces aten::native_batch_norm(+_backward), leaky_relu(+_backward) and
// max_pool2d_with_indices(+_backward) for nn.BatchNorm2d / nn.LeakyReLU / nn.MaxPool2d at
// /root/reference/regression_model.py:15-17,24-26 and two_branch_regression.py:11-13,... in training mode.
//
// All four passes are HBM streaming: Y (raw conv output, bf16 NHWC) is read once per pass with 16-byte accesses,
// one thread per pooled pixel x 8 channels.  No argmax is stored: the backward passes recompute the activations of
// the 2x2 window from Y and pick the first maximum in scan order, which is what PyTorch's max_pool2d does.
//   forward : 2 B/elem read + 0.5 B/elem written
//   reduce  : 2 B/elem + 0.5 B/elem (dP) read
//   apply   : 2 B/elem + 0.5 B/elem read, 2 B/elem written
#include <cstdlib>

#include "ctk_common.h"
#include "ctk_ptx.cuh"

namespace {

using namespace ctk;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// ---------------------------------------------------------------- statistics -> scale/shift (+ running stats)
// moments == 0: sums = [sum(y), sum(y^2)];  moments == 1: sums = [mean, biased variance] (first block, from the Gram matrix)
__global__ void bn_finalize_kernel(const float* __restrict__ sums, int moments, double count, const float* __restrict__ bias,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ num_batches_tracked, float momentum, float eps, int c,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && num_batches_tracked) num_batches_tracked[0] += 1;
  if (i >= c) return;
  const double m = moments ? static_cast<double>(sums[i]) : static_cast<double>(sums[i]) / count;
  double var = moments ? static_cast<double>(sums[c + i])
                       : static_cast<double>(sums[c + i]) / count - m * m;      // biased variance normalises the batch
  var = var > 0.0 ? var : 0.0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float sc = gamma[i] * invstd;
  scale[i] = sc;
  shift[i] = beta[i] - static_cast<float>(m) * sc;
  mean_out[i] = static_cast<float>(m);
  invstd_out[i] = invstd;
  if (running_mean) {
    const float b = bias ? bias[i] : 0.f;                              // the raw conv output excludes the conv bias
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * (static_cast<float>(m) + b);
    running_var[i] = (1.f - momentum) * running_var[i] + momentum * static_cast<float>(unbiased);
  }
}

// ---------------------------------------------------------------- forward: normalise + LeakyReLU + 2x2 max-pool
// A thread keeps ONE 8-channel group for its whole grid-stride walk (the stride is a multiple of c8), so the per-channel
// constants are loaded once, and it handles two pooled pixels per iteration with all eight 16-byte loads issued up front.
__global__ void __launch_bounds__(256)
bn_act_pool_fwd_kernel(const uint4* __restrict__ y, int H, int W, int c8, const float* __restrict__ scale,
                       const float* __restrict__ shift, float slope, __nv_bfloat16* __restrict__ out, int out_cstride,
                       int out_coffset, long long pooled_pixels) {
  const int Hp = H >> 1, Wp = W >> 1;
  const int slots = blockDim.x / c8;
  const int cg = threadIdx.x % c8, slot = threadIdx.x / c8;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = __ldg(scale + cg * 8 + i); sh[i] = __ldg(shift + cg * 8 + i); }
  const long long stride = static_cast<long long>(gridDim.x) * slots;
  for (long long pix0 = blockIdx.x * static_cast<long long>(slots) + slot; pix0 < pooled_pixels; pix0 += 2 * stride) {
    uint4 raw[2][4];
    long long obase[2];
    bool on[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long pix = pix0 + u * stride;
      on[u] = pix < pooled_pixels;
      const long long q = on[u] ? pix : pix0;
      const int px = static_cast<int>(q % Wp);
      const long long t = q / Wp;
      const int py = static_cast<int>(t % Hp);
      const long long n = t / Hp;
      const long long base = ((n * H + 2 * py) * W + 2 * px) * c8 + cg;
      obase[u] = q * static_cast<long long>(out_cstride) + out_coffset + cg * 8;
#pragma unroll
      for (int j = 0; j < 4; ++j) raw[u][j] = __ldcs(y + base + ((j >> 1) * W + (j & 1)) * static_cast<long long>(c8));
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float best[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f[8];
        unpack8(raw[u][j], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float a = fmaf(f[i], sc[i], sh[i]);
          best[i] = j == 0 ? a : fmaxf(best[i], a);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) best[i] = leaky(best[i], slope);     // monotone: activation after the max
      if (on[u]) *reinterpret_cast<uint4*>(out + obase[u]) = pack8(best);
    }
  }
}

// gradient of the activation output for the four positions of one window (8 channels): routes dP to the first maximum
__device__ __forceinline__ void window_grads_raw(const uint4 (&raw)[4], const float (&sc)[8], const float (&sh)[8],
                                                 float slope, const float (&dp)[8], float (&yv)[4][8],
                                                 float (&da)[4][8]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) unpack8(raw[j], yv[j]);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float z[4], a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { z[j] = fmaf(yv[j][i], sc[i], sh[i]); a[j] = leaky(z[j], slope); }
    int arg = 0;
    float best = a[0];
#pragma unroll
    for (int j = 1; j < 4; ++j)
      if (a[j] > best) { best = a[j]; arg = j; }          // strictly greater: the first maximum wins
#pragma unroll
    for (int j = 0; j < 4; ++j) da[j][i] = j == arg ? dp[i] * (z[j] > 0.f ? 1.f : slope) : 0.f;
  }
}
__device__ __forceinline__ void window_grads(const uint4* __restrict__ y, long long base, int W, int c8,
                                             const float (&sc)[8], const float (&sh)[8], float slope,
                                             const float (&dp)[8], float (&yv)[4][8], float (&da)[4][8]) {
  uint4 raw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) raw[j] = __ldg(y + base + ((j >> 1) * W + (j & 1)) * static_cast<long long>(c8));
  window_grads_raw(raw, sc, sh, slope, dp, yv, da);
}

// The pooled-tensor reduction below rebuilds xhat from the bf16-rounded pooled activation as (leaky^-1(P) - beta) / gamma:
// one bf16 ulp of P moves xhat by about |beta / gamma| * 2^-8, and gamma == 0 makes it unrecoverable.  A channel group
// (8 channels, one thread) whose parameters leave that regime is "guarded": the pooled kernel skips it and the
// kernel that reads the raw conv output handles it instead.  Both kernels evaluate this same predicate.
constexpr float kPooledGuardRatio = 8.f;
__device__ __forceinline__ bool pooled_guarded(const float* __restrict__ gamma, const float* __restrict__ beta, int cg) {
  bool guarded = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float g = fabsf(__ldg(gamma + cg * 8 + i)), b = fabsf(__ldg(beta + cg * 8 + i));
    guarded |= !(g > 0.f && b <= kPooledGuardRatio * g);      // also true for NaN parameters
  }
  return guarded;
}

// ---------------------------------------------------------------- backward pass 1: sum(dA), sum(dA * xhat) per channel
// guard_gamma != nullptr: only the guarded channel groups are processed (the rest write zero partial sums)
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const uint4* __restrict__ y, const __nv_bfloat16* __restrict__ dp, int dp_cstride, int dp_coffset,
                     int H, int W, int c8, const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ mean, const float* __restrict__ invstd, float slope,
                     const float* __restrict__ guard_gamma, const float* __restrict__ guard_beta,
                     float* __restrict__ part, long long pooled_pixels) {
  extern __shared__ float red[];                       // [256][16]
  const int Hp = H >> 1, Wp = W >> 1;
  const int cg = threadIdx.x % c8;
  const int slot = threadIdx.x / c8;
  const int slots = blockDim.x / c8;
  const bool active = guard_gamma == nullptr || pooled_guarded(guard_gamma, guard_beta, cg);
  float sc[8], sh[8], mu[8], is[8], s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = __ldg(scale + cg * 8 + i); sh[i] = __ldg(shift + cg * 8 + i);
    mu[i] = __ldg(mean + cg * 8 + i); is[i] = __ldg(invstd + cg * 8 + i);
    s1[i] = 0.f; s2[i] = 0.f;
  }
  for (long long pix = blockIdx.x * static_cast<long long>(slots) + slot; active && pix < pooled_pixels;
       pix += static_cast<long long>(gridDim.x) * slots) {
    const int px = static_cast<int>(pix % Wp);
    const long long t = pix / Wp;
    const int py = static_cast<int>(t % Hp);
    const long long n = t / Hp;
    float dpv[8], yv[4][8], da[4][8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dp + pix * dp_cstride + dp_coffset + cg * 8)), dpv);
    window_grads(y, ((n * H + 2 * py) * W + 2 * px) * c8 + cg, W, c8, sc, sh, slope, dpv, yv, da);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s1[i] += da[j][i];
        s2[i] = fmaf(da[j][i], (yv[j][i] - mu[i]) * is[i], s2[i]);
      }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[threadIdx.x * 16 + i] = s1[i]; red[threadIdx.x * 16 + 8 + i] = s2[i]; }
  __syncthreads();
  if (slot == 0) {
    const int c = c8 * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = 0.f, b = 0.f;
      for (int s = 0; s < slots; ++s) { a += red[(s * c8 + cg) * 16 + i]; b += red[(s * c8 + cg) * 16 + 8 + i]; }
      // one row of partial sums per CTA; ctk::reduce_rows_f32 adds the rows in a fixed order (no atomics: deterministic)
      part[static_cast<size_t>(blockIdx.x) * 2 * c + cg * 8 + i] = a;
      part[static_cast<size_t>(blockIdx.x) * 2 * c + c + cg * 8 + i] = b;
    }
  }
}

// ---------------------------------------------------------------- backward pass 1 from the POOLED tensors only
// dA is non-zero only at each window's argmax, where the activation equals the stored pooled output P.  So
//   sum(dA)       = sum_w dP * f'(P)            f'(P) = P > 0 ? 1 : slope
//   sum(dA*xhat)  = sum_w dP * f'(P) * xhat(P)  xhat(P) = (leaky^-1(P) - beta) / gamma
// needs 1 B per element of Y (P and dP, both quarter size) instead of 2.5 B.
__global__ void __launch_bounds__(256)
bn_bwd_reduce_pooled_kernel(const __nv_bfloat16* __restrict__ pooled, int p_cstride, int p_coffset,
                            const __nv_bfloat16* __restrict__ dp, int dp_cstride, int dp_coffset, int c8,
                            const float* __restrict__ gamma, const float* __restrict__ beta, float slope, int guard,
                            float* __restrict__ part, long long pooled_pixels) {
  extern __shared__ float red[];                       // [256][16]
  const int cg = threadIdx.x % c8;
  const int slot = threadIdx.x / c8;
  const int slots = blockDim.x / c8;
  float ig[8], be[8], s1[8], s2[8];
  const float inv_slope = 1.f / slope;
  const bool active = !(guard && pooled_guarded(gamma, beta, cg));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float g = __ldg(gamma + cg * 8 + i);
    ig[i] = g != 0.f ? 1.f / g : 0.f;
    be[i] = __ldg(beta + cg * 8 + i);
    s1[i] = 0.f; s2[i] = 0.f;
  }
  const long long stride = static_cast<long long>(gridDim.x) * slots;
  for (long long pix0 = blockIdx.x * static_cast<long long>(slots) + slot; active && pix0 < pooled_pixels;
       pix0 += 4 * stride) {
    uint4 rp[4], rd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {                      // eight independent 16-byte loads in flight per thread
      const long long pix = pix0 + u * stride;
      const long long q = pix < pooled_pixels ? pix : pix0;
      rp[u] = __ldcs(reinterpret_cast<const uint4*>(pooled + q * p_cstride + p_coffset + cg * 8));
      rd[u] = __ldcs(reinterpret_cast<const uint4*>(dp + q * dp_cstride + dp_coffset + cg * 8));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (pix0 + u * stride >= pooled_pixels) break;
      float pv[8], dv[8];
      unpack8(rp[u], pv);
      unpack8(rd[u], dv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool pos = pv[i] > 0.f;
        const float da = pos ? dv[i] : dv[i] * slope;
        const float z = pos ? pv[i] : pv[i] * inv_slope;
        s1[i] += da;
        s2[i] = fmaf(da, (z - be[i]) * ig[i], s2[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[threadIdx.x * 16 + i] = s1[i]; red[threadIdx.x * 16 + 8 + i] = s2[i]; }
  __syncthreads();
  if (slot == 0) {
    const int c = c8 * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = 0.f, b = 0.f;
      for (int s = 0; s < slots; ++s) { a += red[(s * c8 + cg) * 16 + i]; b += red[(s * c8 + cg) * 16 + 8 + i]; }
      // one row of partial sums per CTA; ctk::reduce_rows_f32 adds the rows in a fixed order (no atomics: deterministic)
      part[static_cast<size_t>(blockIdx.x) * 2 * c + cg * 8 + i] = a;
      part[static_cast<size_t>(blockIdx.x) * 2 * c + c + cg * 8 + i] = b;
    }
  }
}

// ---------------------------------------------------------------- backward pass 2: dense gradient of the raw conv output
// dY = sc (dA - m1 - xhat m2) with xhat = (y - mu) invstd  ==  sc dA + A y + B,  A = -sc invstd m2,  B = -sc m1 - A mu.
// Same thread <-> channel-group mapping as the forward pass: constants live in registers for the whole walk.
template <int kU, int kMinBlocks>          // pooled pixels per thread and iteration, resident CTAs the register budget allows
__global__ void __launch_bounds__(256, kMinBlocks)
bn_bwd_apply_kernel(const uint4* __restrict__ y, const __nv_bfloat16* __restrict__ dp, int dp_cstride, int dp_coffset,
                    int H, int W, int c8, const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ sums,
                    float inv_count, float slope, uint4* __restrict__ dy, long long pooled_pixels) {
  const int Hp = H >> 1, Wp = W >> 1;
  const int c = c8 * 8;
  const int slots = blockDim.x / c8;
  const int cg = threadIdx.x % c8, slot = threadIdx.x / c8;
  float sc[8], sh[8], ca[8], cb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = __ldg(scale + cg * 8 + i); sh[i] = __ldg(shift + cg * 8 + i);
    const float mu = __ldg(mean + cg * 8 + i), is = __ldg(invstd + cg * 8 + i);
    const float m1 = __ldg(sums + cg * 8 + i) * inv_count, m2 = __ldg(sums + c + cg * 8 + i) * inv_count;
    ca[i] = -sc[i] * is * m2;
    cb[i] = -sc[i] * m1 - ca[i] * mu;
  }
  const long long stride = static_cast<long long>(gridDim.x) * slots;
  for (long long pix0 = blockIdx.x * static_cast<long long>(slots) + slot; pix0 < pooled_pixels; pix0 += kU * stride) {
    uint4 raw[kU][4], rdp[kU];
    long long base[kU];
    bool on[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long pix = pix0 + u * stride;
      on[u] = pix < pooled_pixels;
      const long long q = on[u] ? pix : pix0;
      const int px = static_cast<int>(q % Wp);
      const long long t = q / Wp;
      const int py = static_cast<int>(t % Hp);
      const long long n = t / Hp;
      base[u] = ((n * H + 2 * py) * W + 2 * px) * c8 + cg;
      rdp[u] = __ldcs(reinterpret_cast<const uint4*>(dp + q * dp_cstride + dp_coffset + cg * 8));
#pragma unroll
      for (int j = 0; j < 4; ++j) raw[u][j] = __ldcs(y + base[u] + ((j >> 1) * W + (j & 1)) * static_cast<long long>(c8));
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (!on[u]) break;
      // one 32-bit word (two channels) of the four window positions at a time keeps the live set small
      uint32_t o[4][4];
      const uint32_t dw[4] = {rdp[u].x, rdp[u].y, rdp[u].z, rdp[u].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float g[4][2];
#pragma unroll
        for (int hsel = 0; hsel < 2; ++hsel) {
          const int i = 2 * k + hsel;
          const float dpi = hsel ? __uint_as_float(dw[k] & 0xffff0000u) : __uint_as_float(dw[k] << 16);
          float yv[4], z[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t w = k == 0 ? raw[u][j].x : k == 1 ? raw[u][j].y : k == 2 ? raw[u][j].z : raw[u][j].w;
            yv[j] = hsel ? __uint_as_float(w & 0xffff0000u) : __uint_as_float(w << 16);
            z[j] = fmaf(yv[j], sc[i], sh[i]);
          }
          // LeakyReLU is monotone, so the first maximum of the activations is the first maximum of z
          int arg = 0;
          float best = z[0];
#pragma unroll
          for (int j = 1; j < 4; ++j)
            if (z[j] > best) { best = z[j]; arg = j; }
          const float top = sc[i] * dpi * (best > 0.f ? 1.f : slope);
#pragma unroll
          for (int j = 0; j < 4; ++j) g[j][hsel] = fmaf(ca[i], yv[j], cb[i]) + (j == arg ? top : 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j][k] = pack_bf16x2(g[j][0], g[j][1]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        __stcs(dy + base[u] + ((j >> 1) * W + (j & 1)) * static_cast<long long>(c8), make_uint4(o[j][0], o[j][1], o[j][2], o[j][3]));
    }
  }
}

// EXPERIMENTAL (branch wip/overlap-streams): CTK_BN_BLOCK=128 launches the streaming BatchNorm passes with 128-thread CTAs
// so that one of them fits beside a resident tensor-core conv CTA (register budget).  Default 256.
inline int bn_block_threads() {
  static const int v = [] {
    const char* e = getenv("CTK_BN_BLOCK");
    const int t = e ? atoi(e) : 256;
    return (t == 64 || t == 128 || t == 256) ? t : 256;
  }();
  return v;
}

// grid for the pooled-pixel walks: `slots` pooled pixels per CTA pass, a few CTAs per SM
inline int grid_for_pixels(long long pooled_pixels, int slots, int per_iter) {
  const long long blocks = (pooled_pixels + static_cast<long long>(slots) * per_iter - 1) / (static_cast<long long>(slots) * per_iter);
  const long long cap = static_cast<long long>(ctk::num_sms()) * 8;
  return static_cast<int>(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace

extern "C" {

static int bn_finalize_impl(const float* sums, int moments, double count, const float* bias, const float* gamma,
                            const float* beta, float* running_mean, float* running_var, long long* num_batches_tracked,
                            float momentum, float eps, int channels, float* scale, float* shift, float* mean,
                            float* invstd, void* stream) {
  CTK_REQUIRE(sums && gamma && beta && scale && shift && mean && invstd && channels > 0 && count >= 1.0);
  CTK_REQUIRE((running_mean == nullptr) == (running_var == nullptr));
  bn_finalize_kernel<<<(channels + 127) / 128, 128, 0, ctk::as_stream(stream)>>>(
      sums, moments, count, bias, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, channels,
      scale, shift, mean, invstd);
  return ctk::check_launch();
}

int ctk_bn_finalize(const float* sums, double count, const float* bias, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps,
                    int channels, float* scale, float* shift, float* mean, float* invstd, void* stream) {
  return bn_finalize_impl(sums, 0, count, bias, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps,
                          channels, scale, shift, mean, invstd, stream);
}

int ctk_bn_finalize_moments(const float* moments, double count, const float* bias, const float* gamma, const float* beta,
                            float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                            float eps, int channels, float* scale, float* shift, float* mean, float* invstd,
                            void* stream) {
  return bn_finalize_impl(moments, 1, count, bias, gamma, beta, running_mean, running_var, num_batches_tracked, momentum,
                          eps, channels, scale, shift, mean, invstd, stream);
}

int ctk_bn_act_pool_fwd(const void* y_bf16, int n, int H, int W, int channels, const float* scale, const float* shift,
                        float slope, void* out_bf16, int out_cstride, int out_coffset, void* stream) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(y_bf16 && scale && shift && out_bf16 && n > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0);
  CTK_REQUIRE(channels > 0 && channels % 8 == 0 && out_cstride % 8 == 0 && out_coffset % 8 == 0 &&
              out_coffset + channels <= out_cstride);
  CTK_REQUIRE(channels <= 2048);
  const int c8 = channels / 8;
  const int threads = ((bn_block_threads() >= c8 ? bn_block_threads() : 256) / c8) * c8;   // whole channel groups per CTA
  const long long pooled = static_cast<long long>(n) * (H / 2) * (W / 2);
  bn_act_pool_fwd_kernel<<<grid_for_pixels(pooled, threads / c8, 2), threads, 0, ctk::as_stream(stream)>>>(
      static_cast<const uint4*>(y_bf16), H, W, c8, scale, shift, slope, static_cast<__nv_bfloat16*>(out_bf16),
      out_cstride, out_coffset, pooled);
  return ctk::check_launch();
}

// rows of partial sums: the exact kernel's grid plus the pooled kernel's grid (the guarded entry point runs both)
// full == false: the guarded entry point's second kernel, which normally finds no guarded channel group and only has to
// write its rows of zero partial sums -- one CTA per SM keeps that at a few microseconds
static int bn_reduce_grid_exact(long long pooled, int channels, bool full = true) {
  const int slots = 256 / (channels / 8);
  const long long blocks = (pooled + slots - 1) / slots;
  const long long cap = static_cast<long long>(ctk::num_sms()) * (full ? 4 : 1);
  return static_cast<int>(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}
static int bn_reduce_grid_pooled(long long pooled, int channels) {
  const int slots = 256 / (channels / 8);
  // every CTA ends with one row of 2 * channels partial sums: few, long-running CTAs (>= 64 pixels per slot)
  const long long blocks = (pooled + static_cast<long long>(slots) * 64 - 1) / (static_cast<long long>(slots) * 64);
  const long long cap = static_cast<long long>(ctk::num_sms()) * 3;
  return static_cast<int>(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

size_t ctk_bn_bwd_reduce_workspace_bytes(int channels) {
  return channels > 0 ? static_cast<size_t>(ctk::num_sms()) * 7 * 2 * channels * sizeof(float) : 0;
}

int ctk_bn_bwd_reduce(const void* y_bf16, const void* dp_bf16, int dp_cstride, int dp_coffset, int n, int H, int W,
                      int channels, const float* scale, const float* shift, const float* mean, const float* invstd,
                      float slope, float* sums, void* workspace, size_t workspace_bytes, void* stream) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(y_bf16 && dp_bf16 && scale && shift && mean && invstd && sums && n > 0 && H % 2 == 0 && W % 2 == 0);
  CTK_REQUIRE(channels > 0 && channels % 8 == 0 && channels <= 2048 && 256 % (channels / 8) == 0 &&
              dp_cstride % 8 == 0 && dp_coffset % 8 == 0 && dp_coffset + channels <= dp_cstride);
  cudaStream_t s = ctk::as_stream(stream);
  const long long pooled = static_cast<long long>(n) * (H / 2) * (W / 2);
  const int grid = bn_reduce_grid_exact(pooled, channels);
  CTK_REQUIRE_WORKSPACE(workspace, workspace_bytes, static_cast<size_t>(grid) * 2 * channels * sizeof(float));
  float* part = static_cast<float*>(workspace);
  bn_bwd_reduce_kernel<<<grid, 256, 256 * 16 * sizeof(float), s>>>(
      static_cast<const uint4*>(y_bf16), static_cast<const __nv_bfloat16*>(dp_bf16), dp_cstride, dp_coffset, H, W,
      channels / 8, scale, shift, mean, invstd, slope, nullptr, nullptr, part, pooled);
  int st = ctk::check_launch();
  if (st != CTK_OK) return st;
  return ctk::reduce_rows_f32(part, grid, 2 * channels, 2 * channels, sums, s);
}

static int bn_reduce_pooled_impl(const void* y_bf16, int H, int W, const float* scale, const float* shift,
                                 const float* mean, const float* invstd, const void* pooled_bf16, int p_cstride,
                                 int p_coffset, const void* dp_bf16, int dp_cstride, int dp_coffset,
                                 long long pooled_pixels, int channels, const float* gamma, const float* beta,
                                 float slope, float* sums, void* workspace, size_t workspace_bytes, void* stream) {
  if (pooled_pixels == 0) return CTK_OK;
  CTK_REQUIRE(pooled_bf16 && dp_bf16 && gamma && beta && sums && pooled_pixels > 0 && slope > 0.f);
  CTK_REQUIRE(channels > 0 && channels % 8 == 0 && channels <= 2048 && 256 % (channels / 8) == 0);
  CTK_REQUIRE(dp_cstride % 8 == 0 && dp_coffset % 8 == 0 && dp_coffset + channels <= dp_cstride && p_cstride % 8 == 0 &&
              p_coffset % 8 == 0 && p_coffset + channels <= p_cstride);
  cudaStream_t s = ctk::as_stream(stream);
  const bool guard = y_bf16 != nullptr;
  const int grid_p = bn_reduce_grid_pooled(pooled_pixels, channels);
  const int grid_e = guard ? bn_reduce_grid_exact(pooled_pixels, channels, false) : 0;
  CTK_REQUIRE_WORKSPACE(workspace, workspace_bytes, static_cast<size_t>(grid_p + grid_e) * 2 * channels * sizeof(float));
  float* part = static_cast<float*>(workspace);
  bn_bwd_reduce_pooled_kernel<<<grid_p, 256, 256 * 16 * sizeof(float), s>>>(
      static_cast<const __nv_bfloat16*>(pooled_bf16), p_cstride, p_coffset, static_cast<const __nv_bfloat16*>(dp_bf16),
      dp_cstride, dp_coffset, channels / 8, gamma, beta, slope, guard ? 1 : 0, part, pooled_pixels);
  int st = ctk::check_launch();
  if (st != CTK_OK) return st;
  if (guard) {
    // guarded channel groups (none for freshly initialised or moderately trained BatchNorms: the CTAs then exit at once)
    bn_bwd_reduce_kernel<<<grid_e, 256, 256 * 16 * sizeof(float), s>>>(
        static_cast<const uint4*>(y_bf16), static_cast<const __nv_bfloat16*>(dp_bf16), dp_cstride, dp_coffset, H, W,
        channels / 8, scale, shift, mean, invstd, slope, gamma, beta,
        part + static_cast<size_t>(grid_p) * 2 * channels, pooled_pixels);
    st = ctk::check_launch();
    if (st != CTK_OK) return st;
  }
  return ctk::reduce_rows_f32(part, grid_p + grid_e, 2 * channels, 2 * channels, sums, s);
}

int ctk_bn_bwd_reduce_pooled(const void* pooled_bf16, int p_cstride, int p_coffset, const void* dp_bf16, int dp_cstride,
                             int dp_coffset, long long pooled_pixels, int channels, const float* gamma,
                             const float* beta, float slope, float* sums, void* workspace, size_t workspace_bytes,
                             void* stream) {
  return bn_reduce_pooled_impl(nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr, pooled_bf16, p_cstride, p_coffset, dp_bf16,
                               dp_cstride, dp_coffset, pooled_pixels, channels, gamma, beta, slope, sums, workspace,
                               workspace_bytes, stream);
}

int ctk_bn_bwd_reduce_guarded(const void* y_bf16, int n, int H, int W, const float* scale, const float* shift,
                              const float* mean, const float* invstd, const void* pooled_bf16, int p_cstride,
                              int p_coffset, const void* dp_bf16, int dp_cstride, int dp_coffset, int channels,
                              const float* gamma, const float* beta, float slope, float* sums, void* workspace,
                              size_t workspace_bytes, void* stream) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(y_bf16 && scale && shift && mean && invstd && n > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0);
  return bn_reduce_pooled_impl(y_bf16, H, W, scale, shift, mean, invstd, pooled_bf16, p_cstride, p_coffset, dp_bf16,
                               dp_cstride, dp_coffset, static_cast<long long>(n) * (H / 2) * (W / 2), channels, gamma,
                               beta, slope, sums, workspace, workspace_bytes, stream);
}

int ctk_bn_bwd_apply(const void* y_bf16, const void* dp_bf16, int dp_cstride, int dp_coffset, int n, int H, int W,
                     int channels, const float* scale, const float* shift, const float* mean, const float* invstd,
                     const float* sums, float slope, void* dy_bf16, void* stream) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(y_bf16 && dp_bf16 && scale && shift && mean && invstd && sums && dy_bf16 && n > 0 && H % 2 == 0 &&
              W % 2 == 0);
  CTK_REQUIRE(channels > 0 && channels % 8 == 0 && dp_cstride % 8 == 0 && dp_coffset % 8 == 0 &&
              dp_coffset + channels <= dp_cstride);
  CTK_REQUIRE(channels <= 2048);
  const int c8 = channels / 8;
  const int threads = ((bn_block_threads() >= c8 ? bn_block_threads() : 256) / c8) * c8;
  const long long pooled = static_cast<long long>(n) * (H / 2) * (W / 2);
  const float inv_count = 1.f / (static_cast<float>(n) * H * W);
  // CTK_BN_APPLY_VARIANT: 0 = two pooled pixels per iteration, 2 CTAs per SM (128 registers); 1 / 2 = one pixel, 3 / 4 CTAs;
  // 3 = two pixels, 3 CTAs
  static const int variant = [] { const char* e = getenv("CTK_BN_APPLY_VARIANT"); return e ? atoi(e) : 1; }();
#define CTK_BN_APPLY(U, B)                                                                                              \
  bn_bwd_apply_kernel<U, B><<<grid_for_pixels(pooled, threads / c8, U), threads, 0, ctk::as_stream(stream)>>>(        \
      static_cast<const uint4*>(y_bf16), static_cast<const __nv_bfloat16*>(dp_bf16), dp_cstride, dp_coffset, H, W, c8, \
      scale, shift, mean, invstd, sums, inv_count, slope, static_cast<uint4*>(dy_bf16), pooled)
  if (variant == 1) CTK_BN_APPLY(1, 3);
  else if (variant == 2) CTK_BN_APPLY(1, 4);
  else if (variant == 3) CTK_BN_APPLY(2, 3);
  else CTK_BN_APPLY(2, 2);
#undef CTK_BN_APPLY
  return ctk::check_launch();
}

}  // extern "C"
