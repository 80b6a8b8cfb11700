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

__global__ void pack_first_kernel(const float* __restrict__ w, const float* __restrict__ scale, int cout, int k,
                                  float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * k) return;
  out[i] = w[i] * scale[i / k];
}

// out[o][p*C + c] = w[o][c*HW + p]; one block per (row o, 32-pixel x 32-channel tile), transposed through smem
__global__ void pack_fc1_kernel(const float* __restrict__ w, int channels, int hw, __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[32][33];
  const size_t row = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const float* src = w + row * static_cast<size_t>(channels) * hw;
  __nv_bfloat16* dst = out + row * static_cast<size_t>(channels) * hw;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < channels && p < hw) ? src[static_cast<size_t>(c) * hw + p] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    if (c < channels && p < hw) dst[static_cast<size_t>(p) * channels + c] = __float2bfloat16_rn(tile[threadIdx.x][j]);
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

int ctk_pack_first_weight(const float* w, const float* scale, int cout, int cin, float* w_folded, void* stream) {
  CTK_REQUIRE(w && scale && w_folded && cout > 0 && cin > 0);
  const int total = cout * cin * 9;
  pack_first_kernel<<<(total + 255) / 256, 256, 0, ctk::as_stream(stream)>>>(w, scale, cout, cin * 9, w_folded);
  return ctk::check_launch();
}

int ctk_pack_fc1_weight_bf16(const float* w, int out_features, int channels, int hw, void* w_packed_bf16,
                             void* stream) {
  CTK_REQUIRE(w && w_packed_bf16 && out_features > 0 && out_features <= 65535 && channels > 0 && hw > 0);
  dim3 grid((hw + 31) / 32, (channels + 31) / 32, out_features);
  pack_fc1_kernel<<<grid, dim3(32, 8), 0, ctk::as_stream(stream)>>>(w, channels, hw,
                                                                    static_cast<__nv_bfloat16*>(w_packed_bf16));
  return ctk::check_launch();
}

}  // extern "C"
