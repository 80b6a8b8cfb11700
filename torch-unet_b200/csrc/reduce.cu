// Second stage of every cross-CTA reduction in libctk: the producing kernel writes ONE row of partial sums per CTA
// into a caller-owned workspace (plain stores, no atomics), and these kernels add the rows up in a fixed order with
// fp64 accumulation.  The result therefore does not depend on which CTA finished first: two identical launches give
// bit-identical batch statistics, BatchNorm-backward sums and weight gradients (reductions behind
// aten::native_batch_norm(_backward) / aten::convolution_backward, /root/reference/regression_model.py:14-26,
// two_branch_regression.py:10-31, reached from train_model.py:420-422).
#include "ctk_common.h"

namespace {

// out[c] = sum_{r < rows} part[r * stride + c].  One CTA per 32 columns; warp w adds rows w, w+32, ... (eight loads in
// flight), the 32 warp totals are then added in warp order by the first warp.
template <typename T>
__global__ void __launch_bounds__(1024) reduce_rows_kernel(const T* __restrict__ part, int rows, long long stride, int cols,
                                                           T* __restrict__ out) {
  __shared__ double s_part[32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double acc = 0.0;
  if (c < cols) {
    int r = warp;
    for (; r + 7 * 32 < rows; r += 8 * 32) {
      T v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = part[static_cast<long long>(r + u * 32) * stride + c];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += static_cast<double>(v[u]);
    }
    for (; r < rows; r += 32) acc += static_cast<double>(part[static_cast<long long>(r) * stride + c]);
  }
  s_part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < cols) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < 32; ++w) tot += s_part[w][lane];
    out[c] = static_cast<T>(tot);
  }
}

}  // namespace

namespace ctk {

int reduce_rows_f32(const float* part, int rows, long long stride, int cols, float* out, cudaStream_t s) {
  if (cols <= 0) return CTK_OK;
  reduce_rows_kernel<float><<<(cols + 31) / 32, 1024, 0, s>>>(part, rows, stride, cols, out);
  return check_launch();
}

int reduce_rows_f64(const double* part, int rows, long long stride, int cols, double* out, cudaStream_t s) {
  if (cols <= 0) return CTK_OK;
  reduce_rows_kernel<double><<<(cols + 31) / 32, 1024, 0, s>>>(part, rows, stride, cols, out);
  return check_launch();
}

}  // namespace ctk
