// Per-tile Pearson r of channel 0 vs channel 1 (replaces scipy.stats.pearsonr at
// /root/reference/test-cross-talk-model.py:59-64).  HBM-bound: each tile is read exactly once
// (2 planes x plane_elems x 4 B); five running sums are kept in fp64 so the one-pass formula is
// exact to ~1e-13, far inside the 1e-6 tolerance against the float32 SciPy value.
#include "ctk_common.h"

namespace {

constexpr int kSlices = 8;      // CTAs per tile: 256 tiles -> 2048 CTAs, >13 per SM
constexpr int kThreads = 256;
constexpr int kPartial = 8;     // sx, sy, sxx, syy, sxy, (min0,max0), (min1,max1) packed below

struct Acc {
  double sx, sy, sxx, syy, sxy;
  float mn0, mx0, mn1, mx1;
};

__device__ __forceinline__ void accumulate(Acc& a, const float4& x, const float4& y) {
  const float xs[4] = {x.x, x.y, x.z, x.w};
  const float ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double dx = static_cast<double>(xs[i]);
    const double dy = static_cast<double>(ys[i]);
    a.sx += dx;
    a.sy += dy;
    a.sxx = fma(dx, dx, a.sxx);
    a.syy = fma(dy, dy, a.syy);
    a.sxy = fma(dx, dy, a.sxy);
    a.mn0 = fminf(a.mn0, xs[i]);
    a.mx0 = fmaxf(a.mx0, xs[i]);
    a.mn1 = fminf(a.mn1, ys[i]);
    a.mx1 = fmaxf(a.mx1, ys[i]);
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// grid = n_tiles * kSlices (1-D).  partial[tile][slice][kPartial] doubles.
__global__ void __launch_bounds__(kThreads) pearson_partial_kernel(const float* __restrict__ tiles, int plane_elems,
                                                                   double* __restrict__ partial) {
  const int tile = blockIdx.x / kSlices;
  const int slice = blockIdx.x % kSlices;
  const float4* p0 = reinterpret_cast<const float4*>(tiles + static_cast<size_t>(tile) * 2 * plane_elems);
  const float4* p1 = p0 + plane_elems / 4;
  const int nvec = plane_elems / 4;
  const int per_slice = (nvec + kSlices - 1) / kSlices;
  const int begin = slice * per_slice;
  const int end = min(nvec, begin + per_slice);

  Acc a = {0.0, 0.0, 0.0, 0.0, 0.0, INFINITY, -INFINITY, INFINITY, -INFINITY};
  int i = begin + threadIdx.x;
  // 4 independent 16-byte loads per plane in flight per thread
  for (; i + 3 * kThreads < end; i += 4 * kThreads) {
    float4 x[4], y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = __ldcs(p0 + i + j * kThreads);
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = __ldcs(p1 + i + j * kThreads);
#pragma unroll
    for (int j = 0; j < 4; ++j) accumulate(a, x[j], y[j]);
  }
  for (; i < end; i += kThreads) accumulate(a, __ldcs(p0 + i), __ldcs(p1 + i));

  a.sx = warp_sum(a.sx);
  a.sy = warp_sum(a.sy);
  a.sxx = warp_sum(a.sxx);
  a.syy = warp_sum(a.syy);
  a.sxy = warp_sum(a.sxy);
  a.mn0 = warp_min(a.mn0);
  a.mx0 = warp_max(a.mx0);
  a.mn1 = warp_min(a.mn1);
  a.mx1 = warp_max(a.mx1);

  __shared__ double sh[kThreads / 32][5];
  __shared__ float shm[kThreads / 32][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh[warp][0] = a.sx; sh[warp][1] = a.sy; sh[warp][2] = a.sxx; sh[warp][3] = a.syy; sh[warp][4] = a.sxy;
    shm[warp][0] = a.mn0; shm[warp][1] = a.mx0; shm[warp][2] = a.mn1; shm[warp][3] = a.mx1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s[5] = {0, 0, 0, 0, 0};
    float m[4] = {INFINITY, -INFINITY, INFINITY, -INFINITY};
    for (int w = 0; w < kThreads / 32; ++w) {
      for (int k = 0; k < 5; ++k) s[k] += sh[w][k];
      m[0] = fminf(m[0], shm[w][0]); m[1] = fmaxf(m[1], shm[w][1]);
      m[2] = fminf(m[2], shm[w][2]); m[3] = fmaxf(m[3], shm[w][3]);
    }
    double* out = partial + (static_cast<size_t>(tile) * kSlices + slice) * kPartial;
    for (int k = 0; k < 5; ++k) out[k] = s[k];
    // pack (min,max) float pairs bit-exactly into one double slot each
    out[5] = __hiloint2double(__float_as_int(m[1]), __float_as_int(m[0]));
    out[6] = __hiloint2double(__float_as_int(m[3]), __float_as_int(m[2]));
    out[7] = 0.0;
  }
}

// One warp per tile: the tile's kSlices x kPartial partial sums are staged in shared memory by two coalesced loads and
// added by lane 0 in slice order (the same order as ever: bit-identical r).  One thread per tile walking its 64 dependent
// global loads took 8 us of the 36 us the metric costs per 256 tiles (ncu, round 1); this is launch-latency bound.
__global__ void __launch_bounds__(256) pearson_finalize_kernel(const double* __restrict__ partial, int n_tiles, int plane_elems,
                                                               double* __restrict__ r_out) {
  __shared__ double stage[8][kSlices * kPartial];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * 8 + warp;
  if (tile >= n_tiles) return;
  const double* src = partial + static_cast<size_t>(tile) * kSlices * kPartial;
  for (int i = lane; i < kSlices * kPartial; i += 32) stage[warp][i] = src[i];
  __syncwarp();
  if (lane != 0) return;
  double s[5] = {0, 0, 0, 0, 0};
  float mn0 = INFINITY, mx0 = -INFINITY, mn1 = INFINITY, mx1 = -INFINITY;
  for (int sl = 0; sl < kSlices; ++sl) {
    const double* p = stage[warp] + sl * kPartial;
    for (int k = 0; k < 5; ++k) s[k] += p[k];
    mn0 = fminf(mn0, __int_as_float(__double2loint(p[5])));
    mx0 = fmaxf(mx0, __int_as_float(__double2hiint(p[5])));
    mn1 = fminf(mn1, __int_as_float(__double2loint(p[6])));
    mx1 = fmaxf(mx1, __int_as_float(__double2hiint(p[6])));
  }
  const double n = static_cast<double>(plane_elems);
  double r;
  if (!(mx0 > mn0) || !(mx1 > mn1)) {
    r = __longlong_as_double(0x7ff8000000000000ll);   // constant plane -> NaN (test-cross-talk-model.py:61-62)
  } else {
    const double cxy = s[4] - s[0] * s[1] / n;
    const double cxx = s[2] - s[0] * s[0] / n;
    const double cyy = s[3] - s[1] * s[1] / n;
    r = cxy / sqrt(cxx * cyy);
    r = fmin(1.0, fmax(-1.0, r));                      // scipy clips to [-1, 1]
  }
  r_out[tile] = r;
}

}  // namespace

extern "C" {

size_t ctk_pearson_workspace_bytes(int n_tiles) {
  if (n_tiles <= 0) return 0;
  return static_cast<size_t>(n_tiles) * kSlices * kPartial * sizeof(double);
}

int ctk_pearson_f32(const float* tiles, int n_tiles, int plane_elems, double* r_out, void* workspace,
                    size_t workspace_bytes, void* stream) {
  if (n_tiles == 0) return CTK_OK;
  CTK_REQUIRE(tiles && r_out && workspace && n_tiles > 0 && plane_elems > 0 && plane_elems % 4 == 0);
  CTK_REQUIRE((reinterpret_cast<uintptr_t>(tiles) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0);
  if (workspace_bytes < ctk_pearson_workspace_bytes(n_tiles)) return CTK_ERR_WORKSPACE;
  cudaStream_t s = ctk::as_stream(stream);
  double* partial = static_cast<double*>(workspace);
  pearson_partial_kernel<<<static_cast<unsigned>(n_tiles) * kSlices, kThreads, 0, s>>>(tiles, plane_elems, partial);
  int st = ctk::check_launch();
  if (st != CTK_OK) return st;
  pearson_finalize_kernel<<<(n_tiles + 7) / 8, 256, 0, s>>>(partial, n_tiles, plane_elems, r_out);
  return ctk::check_launch();
}

}  // extern "C"
