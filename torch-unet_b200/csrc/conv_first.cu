// First conv block, eval mode: Conv2d(cin in {1,2}, cout, 3, 1, 1) + folded BN + LeakyReLU + MaxPool2d(2,2),
// fp32 NCHW planes in, NHWC bf16 out.  Replaces /root/reference/regression_model.py:14-17 (cin=2, cout=128)
// and two_branch_regression.py:10-13 (cin=1, cout=64).  K = 9*cin is far too small for a GEMM tile, so
// this is a direct convolution on the fp32 pipe: each thread keeps the taps of CPT output channels in
// registers and walks pooled pixels; a pooled pixel's channels are written by consecutive threads so every
// store instruction of a warp covers whole 128-byte lines of the NHWC output.
#include "ctk_common.h"
#include "ctk_ptx.cuh"

namespace {

constexpr int kTileW = 32;   // pooled pixels per block along W
constexpr int kTileH = 8;    // pooled pixels per block along H
constexpr int kThreads = 256;
constexpr int kInW = 2 * kTileW + 2;   // 66
constexpr int kInH = 2 * kTileH + 2;   // 18
constexpr int kInPitch = 68;

template <int CIN, int COUT, int CPT>
__global__ void __launch_bounds__(kThreads)
conv_first_eval_kernel(const float* __restrict__ x, int c_total, int c_offset, int H, int W,
                       const float* __restrict__ w_folded, const float* __restrict__ shift, float slope,
                       __nv_bfloat16* __restrict__ out, int out_cstride, int out_coffset) {
  constexpr int G = COUT / CPT;             // threads per pooled pixel
  constexpr int SLOTS = kThreads / G;       // pooled pixels processed concurrently
  static_assert(kThreads % G == 0 && (kTileW * kTileH) % SLOTS == 0, "tile/thread mismatch");
  __shared__ float s_in[CIN][kInH][kInPitch];

  const int n = blockIdx.z;
  const int py0 = blockIdx.y * kTileH, px0 = blockIdx.x * kTileW;
  const int Hp = H / 2, Wp = W / 2;

  // stage the (2*tile+2)^2 input window, zero padded
  for (int c = 0; c < CIN; ++c) {
    const float* plane = x + (static_cast<size_t>(n) * c_total + c_offset + c) * H * W;
    for (int i = threadIdx.x; i < kInH * kInW; i += kThreads) {
      const int r = i / kInW, q = i % kInW;
      const int gy = 2 * py0 - 1 + r, gx = 2 * px0 - 1 + q;
      float v = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(plane + static_cast<size_t>(gy) * W + gx);
      s_in[c][r][q] = v;
    }
  }

  const int cg = threadIdx.x % G;
  const int slot = threadIdx.x / G;
  float wr[CPT][CIN * 9];
  float sh[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    sh[j] = __ldg(shift + cg * CPT + j);
#pragma unroll
    for (int k = 0; k < CIN * 9; ++k) wr[j][k] = __ldg(w_folded + (cg * CPT + j) * (CIN * 9) + k);
  }
  __syncthreads();

  for (int p = slot; p < kTileW * kTileH; p += SLOTS) {
    const int py = p / kTileW, px = p % kTileW;
    const int gy = py0 + py, gx = px0 + px;
    float patch[CIN][4][4];
#pragma unroll
    for (int c = 0; c < CIN; ++c)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float2 a = *reinterpret_cast<const float2*>(&s_in[c][2 * py + r][2 * px]);
        const float2 b = *reinterpret_cast<const float2*>(&s_in[c][2 * py + r][2 * px + 2]);
        patch[c][r][0] = a.x; patch[c][r][1] = a.y; patch[c][r][2] = b.x; patch[c][r][3] = b.y;
      }
    float res[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      float acc[4] = {sh[j], sh[j], sh[j], sh[j]};
#pragma unroll
      for (int c = 0; c < CIN; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float wv = wr[j][c * 9 + ky * 3 + kx];
            acc[0] = fmaf(wv, patch[c][ky][kx], acc[0]);
            acc[1] = fmaf(wv, patch[c][ky][kx + 1], acc[1]);
            acc[2] = fmaf(wv, patch[c][ky + 1][kx], acc[2]);
            acc[3] = fmaf(wv, patch[c][ky + 1][kx + 1], acc[3]);
          }
      const float m01 = fmaxf(ctk::leaky(acc[0], slope), ctk::leaky(acc[1], slope));
      const float m23 = fmaxf(ctk::leaky(acc[2], slope), ctk::leaky(acc[3], slope));
      res[j] = fmaxf(m01, m23);
    }
    if (gy < Hp && gx < Wp) {
      __nv_bfloat16* dst =
          out + (static_cast<size_t>(n) * Hp * Wp + static_cast<size_t>(gy) * Wp + gx) * out_cstride + out_coffset + cg * CPT;
      if constexpr (CPT == 8) {
        uint4 v;
        v.x = ctk::pack_bf16x2(res[0], res[1]); v.y = ctk::pack_bf16x2(res[2], res[3]);
        v.z = ctk::pack_bf16x2(res[4], res[5]); v.w = ctk::pack_bf16x2(res[6], res[7]);
        *reinterpret_cast<uint4*>(dst) = v;
      } else {
        uint2 v;
        v.x = ctk::pack_bf16x2(res[0], res[1]); v.y = ctk::pack_bf16x2(res[2], res[3]);
        *reinterpret_cast<uint2*>(dst) = v;
      }
    }
  }
}

}  // namespace

extern "C" int ctk_conv_first_eval(const float* x, int n, int c_total, int c_offset, int cin, int H, int W,
                                   const float* w_folded, const float* shift, int cout, float slope, void* out_bf16,
                                   int out_cstride, int out_coffset, void* stream) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(x && w_folded && shift && out_bf16 && n > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0);
  CTK_REQUIRE(c_offset >= 0 && c_offset + cin <= c_total && out_coffset >= 0 && out_coffset + cout <= out_cstride);
  CTK_REQUIRE(out_cstride % 8 == 0 && out_coffset % 8 == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0);
  CTK_REQUIRE(n <= 65535);
  dim3 grid((W / 2 + kTileW - 1) / kTileW, (H / 2 + kTileH - 1) / kTileH, n);
  cudaStream_t s = ctk::as_stream(stream);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(out_bf16);
  if (cin == 1 && cout == 64) {
    conv_first_eval_kernel<1, 64, 8><<<grid, kThreads, 0, s>>>(x, c_total, c_offset, H, W, w_folded, shift, slope, out,
                                                               out_cstride, out_coffset);
  } else if (cin == 2 && cout == 128) {
    conv_first_eval_kernel<2, 128, 4><<<grid, kThreads, 0, s>>>(x, c_total, c_offset, H, W, w_folded, shift, slope,
                                                                out, out_cstride, out_coffset);
  } else {
    return CTK_ERR_UNSUPPORTED;
  }
  return ctk::check_launch();
}
