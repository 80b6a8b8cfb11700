// Parameter re-packing kernels: the reference's state_dict layout (fp32 NCHW conv weights, [out,in] FC
// weights with NCHW-flatten columns, separate BN tensors) -> the derived cache the compute kernels read
// (tap-major bf16 conv weights, NHWC-column bf16 FC1 weight, folded BN scale/shift).  Pure data movement.
#include "ctk_common.h"

#include <algorithm>
#include <cuda_bf16.h>

namespace {

__global__ void fold_bn_kernel(const float* __restrict__ bias, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ rmean,
                               const float* __restrict__ rvar, float eps, int c, float* __restrict__ scale,
                               float* __restrict__ shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  // y = gamma * (acc + bias - mean) / sqrt(var + eps) + beta
  const float sc = gamma[i] / sqrtf(rvar[i] + eps);
  const float b = bias ? bias[i] : 0.f;
  scale[i] = sc;
  shift[i] = (b - rmean[i]) * sc + beta[i];
}

// out[tap][co][ci] = w[co][ci][tap]
__global__ void pack_conv_kernel(const float* __restrict__ w, int cout, int cin, __nv_bfloat16* __restrict__ out) {
  const size_t total = static_cast<size_t>(9) * cout * cin;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % cin);
    const int co = static_cast<int>((i / cin) % cout);
    const int tap = static_cast<int>(i / (static_cast<size_t>(cin) * cout));
    out[i] = __float2bfloat16_rn(w[(static_cast<size_t>(co) * cin + ci) * 9 + tap]);
  }
}

// fp32-class operand: out[tap][co][3*cin] = [ hi(w) | hi(w) | lo(w) ] along K, to be multiplied with activations laid out
// [ hi(x) | lo(x) | hi(x) ]: x*w ~ hi*hi + lo*hi + hi*lo (the dropped lo*lo term is ~2^-17 relative)
__global__ void pack_conv_split_kernel(const float* __restrict__ w, int cout, int cin, __nv_bfloat16* __restrict__ out) {
  const size_t total = static_cast<size_t>(9) * cout * cin;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % cin);
    const int co = static_cast<int>((i / cin) % cout);
    const int tap = static_cast<int>(i / (static_cast<size_t>(cin) * cout));
    const float v = w[(static_cast<size_t>(co) * cin + ci) * 9 + tap];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* row = out + (static_cast<size_t>(tap) * cout + co) * 3 * cin;
    row[ci] = hi;
    row[cin + ci] = hi;
    row[2 * cin + ci] = lo;
  }
}

// dgrad operand: out[tap'][ci][co] = w[co][ci][8 - tap']  (weights rotated by 180 degrees, Cin/Cout swapped)
__global__ void pack_conv_dgrad_kernel(const float* __restrict__ w, int cout, int cin, __nv_bfloat16* __restrict__ out) {
  const size_t total = static_cast<size_t>(9) * cout * cin;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % cout);
    const int ci = static_cast<int>((i / cout) % cin);
    const int tap = static_cast<int>(i / (static_cast<size_t>(cin) * cout));
    out[i] = __float2bfloat16_rn(w[(static_cast<size_t>(co) * cin + ci) * 9 + (8 - tap)]);
  }
}

// Both operand layouts of up to eight conv weights in ONE launch (the training step packs every tensor-core conv's forward
// and dgrad copy once per step: twelve 5-us launches otherwise).  Same values as pack_conv_kernel / pack_conv_dgrad_kernel.
constexpr int kPackMulti = 8;
struct PackMulti {
  const float* w[kPackMulti];
  __nv_bfloat16* fwd[kPackMulti];
  __nv_bfloat16* dgrad[kPackMulti];
  int cout[kPackMulti], cin[kPackMulti];
  long long begin[kPackMulti + 1];
  int n;
};
__global__ void pack_conv_multi_kernel(const PackMulti p) {
  const long long total = p.begin[p.n];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int t = 0;
    while (t + 1 < p.n && i >= p.begin[t + 1]) ++t;
    const long long j = i - p.begin[t];
    const int cin = p.cin[t], cout = p.cout[t];
    const int ci = static_cast<int>(j % cin);
    const int co = static_cast<int>((j / cin) % cout);
    const int tap = static_cast<int>(j / (static_cast<long long>(cin) * cout));
    const __nv_bfloat16 v = __float2bfloat16_rn(p.w[t][(static_cast<size_t>(co) * cin + ci) * 9 + tap]);
    p.fwd[t][j] = v;                                                                       // [tap][co][ci]
    p.dgrad[t][(static_cast<size_t>(8 - tap) * cin + ci) * cout + co] = v;                 // [8 - tap][ci][co]
  }
}

// out[(c*HW + p)][n] = feat[n][p*cstride + c]   (NHWC activations -> NCHW-flatten rows, batch contiguous, zero padded to ld)
__global__ void feat_transpose_kernel(const __nv_bfloat16* __restrict__ feat, int n, int hw, int channels,
                                      __nv_bfloat16* __restrict__ out, int ld) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int p = blockIdx.z;
  const int c0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int nn = n0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (nn < n && c < channels)
        ? feat[(static_cast<size_t>(nn) * hw + p) * channels + c] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, nn = n0 + threadIdx.x;
    if (c < channels && nn < ld) out[(static_cast<size_t>(c) * hw + p) * ld + nn] = tile[threadIdx.x][j];
  }
}

// Same transpose for channels % 64 == 0 and ld % 8 == 0 (every model shape): one CTA moves 64 channels x 256 images of
// one pixel.  In: one full 128-byte line per image (16-byte loads); through shared memory as bf16 channel pairs with an
// odd word pitch (both phases conflict-free); out: 8 images per 16-byte store, 64 contiguous bytes per channel row and
// warp instruction.  The 32x32 element-wise kernel above ran at 18 % of HBM bandwidth (ncu: issue-bound).
constexpr int kFtPitch = 33;    // words per image row in shared memory (32 channel pairs + 1)
__global__ void __launch_bounds__(256)
feat_transpose_vec_kernel(const __nv_bfloat16* __restrict__ feat, int n, int hw, int channels,
                          __nv_bfloat16* __restrict__ out, int ld) {
  __shared__ uint32_t tile[256 * kFtPitch];
  const int p = blockIdx.z;
  const int c0 = blockIdx.x * 64, n0 = blockIdx.y * 256;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // load: 8 lanes cover the 128 bytes of one image, a warp instruction covers 4 images
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = k * 32 + warp * 4 + (lane >> 3);
    const int nn = n0 + r;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (nn < n)
      v = __ldg(reinterpret_cast<const uint4*>(feat + (static_cast<size_t>(nn) * hw + p) * channels + c0) + (lane & 7));
    uint32_t* dst = tile + r * kFtPitch + 4 * (lane & 7);
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
  __syncthreads();
  // store: lane -> (channel pair lane & 7, image group lane >> 3); a warp instruction covers 8 pairs x 32 images
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int item = k * 8 + warp;                  // 32 items = 4 pair groups x 8 image blocks of 32
    const int cp = (item & 3) * 8 + (lane & 7);
    const int r0 = (item >> 2) * 32 + (lane >> 3) * 8;
    if (n0 + r0 >= ld) continue;
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = tile[(r0 + i) * kFtPitch + cp];
    uint4 lo, hi;                                   // even channel = low halves, odd channel = high halves
    lo.x = __byte_perm(w[0], w[1], 0x5410); hi.x = __byte_perm(w[0], w[1], 0x7632);
    lo.y = __byte_perm(w[2], w[3], 0x5410); hi.y = __byte_perm(w[2], w[3], 0x7632);
    lo.z = __byte_perm(w[4], w[5], 0x5410); hi.z = __byte_perm(w[4], w[5], 0x7632);
    lo.w = __byte_perm(w[6], w[7], 0x5410); hi.w = __byte_perm(w[6], w[7], 0x7632);
    const size_t row = static_cast<size_t>(c0 + 2 * cp) * hw + p;
    *reinterpret_cast<uint4*>(out + row * ld + n0 + r0) = lo;
    *reinterpret_cast<uint4*>(out + (row + hw) * ld + n0 + r0) = hi;
  }
}

// w_t[p*C + c][o] = w[o][c*HW + p]  (FC1 weight, NHWC-ordered rows, output features contiguous: K-major B operand of dfeat)
__global__ void pack_fc1_t_kernel(const float* __restrict__ w, int out_features, int channels, int hw,
                                  __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[32][33];
  const int c = blockIdx.z;
  const int p0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int o = o0 + j, p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (o < out_features && p < hw) ? w[(static_cast<size_t>(o) * channels + c) * hw + p] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, o = o0 + threadIdx.x;
    if (p < hw && o < out_features)
      out[(static_cast<size_t>(p) * channels + c) * out_features + o] = __float2bfloat16_rn(tile[threadIdx.x][j]);
  }
}

__global__ void pack_first_kernel(const float* __restrict__ w, const float* __restrict__ scale, int cout, int k,
                                  float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * k) return;
  out[i] = scale ? w[i] * scale[i / k] : w[i];
}

// out[o][p*C + c] = w[o][c*HW + p]; one block per (row o, 64-pixel x 64-channel tile), transposed through smem:
// 256-byte fp32 row reads (float4 per thread), 128-byte bf16 row writes (four channels per thread)
__global__ void __launch_bounds__(256)
pack_fc1_kernel(const float* __restrict__ w, int channels, int hw, __nv_bfloat16* __restrict__ out,
                __nv_bfloat16* __restrict__ out_lo) {
  __shared__ float tile[64][65];                               // [channel][pixel]
  const size_t row = blockIdx.z;
  const int c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  const float* src = w + row * static_cast<size_t>(channels) * hw;
  __nv_bfloat16* dst = out + row * static_cast<size_t>(channels) * hw;
  __nv_bfloat16* dst_lo = out_lo ? out_lo + row * static_cast<size_t>(channels) * hw : nullptr;
  const int t = threadIdx.x;
  const bool vec = (hw % 4 == 0) && p0 + 64 <= hw;
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int j = pass * 16 + (t >> 4), q = (t & 15) * 4;      // channel row j, pixels q..q+3
    const int c = c0 + j;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < channels) {
      if (vec) {
        v = __ldcs(reinterpret_cast<const float4*>(src + static_cast<size_t>(c) * hw + p0 + q));
      } else {
        const float* s0 = src + static_cast<size_t>(c) * hw;
        if (p0 + q < hw) v.x = s0[p0 + q];
        if (p0 + q + 1 < hw) v.y = s0[p0 + q + 1];
        if (p0 + q + 2 < hw) v.z = s0[p0 + q + 2];
        if (p0 + q + 3 < hw) v.w = s0[p0 + q + 3];
      }
    }
    tile[j][q] = v.x; tile[j][q + 1] = v.y; tile[j][q + 2] = v.z; tile[j][q + 3] = v.w;
  }
  __syncthreads();
  const bool cvec = (channels % 4 == 0) && c0 + 64 <= channels;
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int j = pass * 16 + (t >> 4), q = (t & 15) * 4;      // pixel row j, channels q..q+3
    const int p = p0 + j;
    if (p >= hw) continue;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = tile[q + k][j];
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      hi[k] = __float2bfloat16_rn(v[k]);
      lo[k] = __float2bfloat16_rn(v[k] - __bfloat162float(hi[k]));
    }
    const size_t o = static_cast<size_t>(p) * channels + c0 + q;
    if (cvec) {
      *reinterpret_cast<uint2*>(dst + o) = *reinterpret_cast<const uint2*>(hi);
      if (dst_lo) *reinterpret_cast<uint2*>(dst_lo + o) = *reinterpret_cast<const uint2*>(lo);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (c0 + q + k < channels) {
          dst[o + k] = hi[k];
          if (dst_lo) dst_lo[o + k] = lo[k];
        }
    }
  }
}

}  // namespace

extern "C" {

int ctk_fold_bn_eval(const float* bias, const float* gamma, const float* beta, const float* rmean, const float* rvar,
                     float eps, int channels, float* scale, float* shift, void* stream) {
  CTK_REQUIRE(gamma && beta && rmean && rvar && scale && shift && channels > 0);
  fold_bn_kernel<<<(channels + 127) / 128, 128, 0, ctk::as_stream(stream)>>>(bias, gamma, beta, rmean, rvar, eps,
                                                                             channels, scale, shift);
  return ctk::check_launch();
}

int ctk_pack_conv_weight_bf16(const float* w, int cout, int cin, void* w_packed_bf16, void* stream) {
  CTK_REQUIRE(w && w_packed_bf16 && cout > 0 && cin > 0);
  const size_t total = static_cast<size_t>(9) * cout * cin;
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 4096));
  pack_conv_kernel<<<blocks, 256, 0, ctk::as_stream(stream)>>>(w, cout, cin,
                                                               static_cast<__nv_bfloat16*>(w_packed_bf16));
  return ctk::check_launch();
}

int ctk_pack_conv_weight_dgrad_bf16(const float* w, int cout, int cin, void* w_packed_bf16, void* stream) {
  CTK_REQUIRE(w && w_packed_bf16 && cout > 0 && cin > 0);
  const size_t total = static_cast<size_t>(9) * cout * cin;
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 4096));
  pack_conv_dgrad_kernel<<<blocks, 256, 0, ctk::as_stream(stream)>>>(w, cout, cin,
                                                                     static_cast<__nv_bfloat16*>(w_packed_bf16));
  return ctk::check_launch();
}

int ctk_pack_conv_weights_train(int n_layers, const float* const* w, const int* cout, const int* cin,
                                void* const* w_fwd_bf16, void* const* w_dgrad_bf16, void* stream) {
  if (n_layers == 0) return CTK_OK;
  CTK_REQUIRE(n_layers > 0 && n_layers <= kPackMulti && w && cout && cin && w_fwd_bf16 && w_dgrad_bf16);
  PackMulti p = {};
  p.n = n_layers;
  for (int t = 0; t < n_layers; ++t) {
    CTK_REQUIRE(w[t] && w_fwd_bf16[t] && w_dgrad_bf16[t] && cout[t] > 0 && cin[t] > 0);
    p.w[t] = w[t];
    p.fwd[t] = static_cast<__nv_bfloat16*>(w_fwd_bf16[t]);
    p.dgrad[t] = static_cast<__nv_bfloat16*>(w_dgrad_bf16[t]);
    p.cout[t] = cout[t];
    p.cin[t] = cin[t];
    p.begin[t + 1] = p.begin[t] + 9ll * cout[t] * cin[t];
  }
  const long long total = p.begin[n_layers];
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
  pack_conv_multi_kernel<<<blocks, 256, 0, ctk::as_stream(stream)>>>(p);
  return ctk::check_launch();
}

int ctk_feat_transpose_bf16(const void* feat_bf16, int n, int hw, int channels, void* out_bf16, int ld, void* stream) {
  CTK_REQUIRE(feat_bf16 && out_bf16 && n > 0 && hw > 0 && hw <= 65535 && channels > 0 && ld >= n);
  if (channels % 64 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(feat_bf16) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0) {
    dim3 grid(channels / 64, (ld + 255) / 256, hw);
    feat_transpose_vec_kernel<<<grid, 256, 0, ctk::as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(feat_bf16), n, hw, channels, static_cast<__nv_bfloat16*>(out_bf16), ld);
    return ctk::check_launch();
  }
  dim3 grid((channels + 31) / 32, (ld + 31) / 32, hw);
  feat_transpose_kernel<<<grid, dim3(32, 8), 0, ctk::as_stream(stream)>>>(
      static_cast<const __nv_bfloat16*>(feat_bf16), n, hw, channels, static_cast<__nv_bfloat16*>(out_bf16), ld);
  return ctk::check_launch();
}

int ctk_pack_fc1_weight_t_bf16(const float* w, int out_features, int channels, int hw, void* w_t_bf16, void* stream) {
  CTK_REQUIRE(w && w_t_bf16 && out_features > 0 && channels > 0 && channels <= 65535 && hw > 0);
  dim3 grid((hw + 31) / 32, (out_features + 31) / 32, channels);
  pack_fc1_t_kernel<<<grid, dim3(32, 8), 0, ctk::as_stream(stream)>>>(w, out_features, channels, hw,
                                                                      static_cast<__nv_bfloat16*>(w_t_bf16));
  return ctk::check_launch();
}

int ctk_pack_first_weight(const float* w, const float* scale, int cout, int cin, float* w_folded, void* stream) {
  CTK_REQUIRE(w && w_folded && cout > 0 && cin > 0);
  const int total = cout * cin * 9;
  pack_first_kernel<<<(total + 255) / 256, 256, 0, ctk::as_stream(stream)>>>(w, scale, cout, cin * 9, w_folded);
  return ctk::check_launch();
}

int ctk_pack_fc1_weight_bf16(const float* w, int out_features, int channels, int hw, void* w_packed_bf16,
                             void* stream) {
  CTK_REQUIRE(w && w_packed_bf16 && out_features > 0 && out_features <= 65535 && channels > 0 && hw > 0);
  dim3 grid((hw + 63) / 64, (channels + 63) / 64, out_features);
  pack_fc1_kernel<<<grid, 256, 0, ctk::as_stream(stream)>>>(w, channels, hw, static_cast<__nv_bfloat16*>(w_packed_bf16),
                                                            nullptr);
  return ctk::check_launch();
}

int ctk_pack_fc1_weight_split_bf16(const float* w, int out_features, int channels, int hw, void* w_hi_bf16,
                                   void* w_lo_bf16, void* stream) {
  CTK_REQUIRE(w && w_hi_bf16 && w_lo_bf16 && out_features > 0 && out_features <= 65535 && channels > 0 && hw > 0);
  dim3 grid((hw + 63) / 64, (channels + 63) / 64, out_features);
  pack_fc1_kernel<<<grid, 256, 0, ctk::as_stream(stream)>>>(w, channels, hw, static_cast<__nv_bfloat16*>(w_hi_bf16),
                                                            static_cast<__nv_bfloat16*>(w_lo_bf16));
  return ctk::check_launch();
}

int ctk_pack_conv_weight_split_bf16(const float* w, int cout, int cin, void* w_packed_bf16, void* stream) {
  CTK_REQUIRE(w && w_packed_bf16 && cout > 0 && cin > 0);
  const size_t total = static_cast<size_t>(9) * cout * cin;
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 4096));
  pack_conv_split_kernel<<<blocks, 256, 0, ctk::as_stream(stream)>>>(w, cout, cin,
                                                                     static_cast<__nv_bfloat16*>(w_packed_bf16));
  return ctk::check_launch();
}

}  // extern "C"
