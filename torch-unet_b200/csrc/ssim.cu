// Mean structural similarity of channel 0 vs channel 1 of every tile -- the `ssim(images[j][0], images[j][1],
// data_range=...)` column of the reference's evaluation loop (/root/reference/test-cross-talk-model.py:80-82), i.e.
// scikit-image's structural_similarity with its defaults: 7x7 uniform window, K1 = 0.01, K2 = 0.03, sample covariance,
// float32 arithmetic for float32 images, mean (float64) over the image cropped by 3 pixels.
//
// The arithmetic follows what scikit-image executes, step by step, so that the result agrees with it to float64
// summation order rather than to "float32 noise":
//   * the five window means (x, y, x*x, y*y, x*y; the products rounded to float32 first) are scipy.ndimage.uniform_filter's:
//     one 1-D pass down the rows, rounded to float32, then one along the columns, rounded to float32, each 7-tap mean
//     accumulated in float64;
//   * the SSIM map is evaluated in float32 in the library's operation order with explicitly rounded intrinsics
//     (no FMA contraction);
//   * only pixels whose window lies inside the tile enter the mean (the crop), so the filter's boundary mode never matters.
// Work split: one CTA per (strip of output rows, tile); a thread owns a column for the row pass and a pixel for the
// column pass, with the row-pass results of a few rows exchanged through shared memory.  Each tile's two planes are read
// once from HBM (512 KB, re-touched 7x through L1/L2); the kernel is arithmetic-bound on the fp64 adds, not HBM-bound.
#include "ctk_common.h"

namespace {

constexpr int kWin = 7, kPad = 3;
constexpr int kRows = 4;              // rows of row-pass results staged per exchange
constexpr int kThreads = 256;
constexpr int kStrip = 32;            // output rows per strip
constexpr int kSlices = 8;            // CTAs (and partial sums) per tile: slice s takes strips s, s + 8, ...

struct MinMax { float lo, hi; };

// range[n] = max over both planes - min over both planes (float32, like np.max([...]) - np.min([...]) at the call site);
// also clears the tile's accumulator.
__global__ void __launch_bounds__(kThreads)
ssim_range_kernel(const float* __restrict__ tiles, int plane_elems, float* __restrict__ range, double* __restrict__ acc) {
  __shared__ float s_lo[kThreads / 32], s_hi[kThreads / 32];
  const float4* src = reinterpret_cast<const float4*>(tiles + static_cast<size_t>(blockIdx.x) * 2 * plane_elems);
  const int n4 = plane_elems / 2;     // two planes, four floats per load
  float lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < n4; i += kThreads) {
    const float4 v = __ldg(src + i);
    lo = fminf(fminf(lo, v.x), fminf(fminf(v.y, v.z), v.w));
    hi = fmaxf(fmaxf(hi, v.x), fmaxf(fmaxf(v.y, v.z), v.w));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; ++w) { lo = fminf(lo, s_lo[w]); hi = fmaxf(hi, s_hi[w]); }
    range[blockIdx.x] = __fsub_rn(hi, lo);
  }
}

__device__ __forceinline__ float mean7(double s) { return static_cast<float>(s / 7.0); }

__global__ void __launch_bounds__(kThreads)
ssim_strip_kernel(const float* __restrict__ tiles, int H, int W, const float* __restrict__ range,
                  double* __restrict__ acc) {
  extern __shared__ float vbuf[];                       // [kRows][5][W] row-pass means
  __shared__ double s_part[kThreads / 32];
  const int tile = blockIdx.y;
  const float* x = tiles + static_cast<size_t>(tile) * 2 * H * W;
  const float* y = x + static_cast<size_t>(H) * W;
  const float R = range[tile];
  const float t1 = __fmul_rn(0.01f, R), t2 = __fmul_rn(0.03f, R);
  const float c1 = __fmul_rn(t1, t1), c2 = __fmul_rn(t2, t2);              // (K * data_range) ** 2 in float32
  const float cov_norm = static_cast<float>(49.0 / 48.0);                  // NP / (NP - 1), cast to the arrays' float32
  double total = 0.0;

  // a tile is covered by kSlices CTAs; CTA s takes the row strips s, s + kSlices, ... and leaves ONE partial sum, added up
  // in slice order by the finalize kernel (no floating-point atomics: the result is reproducible bit for bit)
  for (int r_begin = kPad + blockIdx.x * kStrip; r_begin < H - kPad; r_begin += kSlices * kStrip) {
  const int r_end = min(r_begin + kStrip, H - kPad);
  for (int r0 = r_begin; r0 < r_end; r0 += kRows) {
    const int rows = min(kRows, r_end - r0);
    // ---- pass 1 (down the rows, scipy's axis 0): thread <-> column
    for (int c = threadIdx.x; c < W; c += kThreads) {
      for (int rr = 0; rr < rows; ++rr) {
        const int r = r0 + rr;
        double sx = 0.0, sy = 0.0, sxx = 0.0, syy = 0.0, sxy = 0.0;
#pragma unroll
        for (int k = -kPad; k <= kPad; ++k) {
          const float a = __ldg(x + static_cast<size_t>(r + k) * W + c);
          const float b = __ldg(y + static_cast<size_t>(r + k) * W + c);
          sx += static_cast<double>(a);
          sy += static_cast<double>(b);
          sxx += static_cast<double>(__fmul_rn(a, a));
          syy += static_cast<double>(__fmul_rn(b, b));
          sxy += static_cast<double>(__fmul_rn(a, b));
        }
        float* v = vbuf + static_cast<size_t>(rr) * 5 * W + c;
        v[0] = mean7(sx);
        v[W] = mean7(sy);
        v[2 * W] = mean7(sxx);
        v[3 * W] = mean7(syy);
        v[4 * W] = mean7(sxy);
      }
    }
    __syncthreads();
    // ---- pass 2 (along the columns, axis 1) + the SSIM map: thread <-> pixel
    const int wc = W - 2 * kPad;
    for (int i = threadIdx.x; i < rows * wc; i += kThreads) {
      const int rr = i / wc, c = kPad + i % wc;
      const float* v = vbuf + static_cast<size_t>(rr) * 5 * W + c;
      float u[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        double s = 0.0;
#pragma unroll
        for (int k = -kPad; k <= kPad; ++k) s += static_cast<double>(v[q * W + k]);
        u[q] = mean7(s);
      }
      const float ux = u[0], uy = u[1], uxx = u[2], uyy = u[3], uxy = u[4];
      const float vx = __fmul_rn(cov_norm, __fsub_rn(uxx, __fmul_rn(ux, ux)));
      const float vy = __fmul_rn(cov_norm, __fsub_rn(uyy, __fmul_rn(uy, uy)));
      const float vxy = __fmul_rn(cov_norm, __fsub_rn(uxy, __fmul_rn(ux, uy)));
      const float a1 = __fadd_rn(__fmul_rn(__fmul_rn(2.f, ux), uy), c1);                       // 2 * ux * uy + C1
      const float a2 = __fadd_rn(__fmul_rn(2.f, vxy), c2);                                      // 2 * vxy + C2
      const float b1 = __fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), c1);          // ux**2 + uy**2 + C1
      const float b2 = __fadd_rn(__fadd_rn(vx, vy), c2);                                        // vx + vy + C2
      const float s = __fdiv_rn(__fmul_rn(a1, a2), __fmul_rn(b1, b2));                          // (A1 * A2) / (B1 * B2)
      total += static_cast<double>(s);
    }
    __syncthreads();
  }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = total;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += s_part[w];
    acc[static_cast<size_t>(tile) * kSlices + blockIdx.x] = t;
  }
}

__global__ void ssim_finalize_kernel(const double* __restrict__ acc, int n, double count, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double t = 0.0;
#pragma unroll
  for (int s = 0; s < kSlices; ++s) t += acc[static_cast<size_t>(i) * kSlices + s];
  out[i] = t / count;
}

}  // namespace

extern "C" {

size_t ctk_tile_ssim_workspace_bytes(int n_tiles) {
  if (n_tiles <= 0) return 0;
  return static_cast<size_t>(n_tiles) * (kSlices * sizeof(double) + sizeof(float)) + 8;
}

int ctk_tile_ssim_f32(const float* tiles, int n_tiles, int H, int W, double* ssim_out, void* workspace,
                      size_t workspace_bytes, void* stream) {
  if (n_tiles == 0) return CTK_OK;
  CTK_REQUIRE(tiles && ssim_out && workspace && n_tiles > 0 && n_tiles <= 65535);
  CTK_REQUIRE(H >= kWin && W >= kWin && (static_cast<long long>(H) * W) % 2 == 0);
  CTK_REQUIRE((reinterpret_cast<uintptr_t>(tiles) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0);
  const size_t smem = static_cast<size_t>(kRows) * 5 * W * sizeof(float);
  if (smem > 46 * 1024) return CTK_ERR_UNSUPPORTED;     // W <= 588
  if (workspace_bytes < ctk_tile_ssim_workspace_bytes(n_tiles)) return CTK_ERR_WORKSPACE;
  cudaStream_t s = ctk::as_stream(stream);
  double* acc = static_cast<double*>(workspace);
  float* range = reinterpret_cast<float*>(acc + static_cast<size_t>(n_tiles) * kSlices);
  ssim_range_kernel<<<n_tiles, kThreads, 0, s>>>(tiles, H * W, range, acc);
  int st = ctk::check_launch();
  if (st != CTK_OK) return st;
  ssim_strip_kernel<<<dim3(kSlices, n_tiles), kThreads, smem, s>>>(tiles, H, W, range, acc);
  st = ctk::check_launch();
  if (st != CTK_OK) return st;
  const double count = static_cast<double>(H - 2 * kPad) * (W - 2 * kPad);
  ssim_finalize_kernel<<<(n_tiles + 255) / 256, 256, 0, s>>>(acc, n_tiles, count, ssim_out);
  return ctk::check_launch();
}

}  // extern "C"
