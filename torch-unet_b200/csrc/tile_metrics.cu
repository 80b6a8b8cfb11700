// Per-tile comparison metrics of channel 0 vs channel 1, fused into one pass over the input tiles plus one histogram
// pass that re-reads them from L2 (replaces the host loop of
// /root/reference/test-cross-talk-model.py:52-85 for three of its metrics):
//   Pearson r               scipy.stats.pearsonr(img0.flatten(), img1.flatten()), NaN on a constant plane      (:59-64)
//   RMSE                    np.sqrt(np.mean((img0 - img1) ** 2))                                                (:79)
//   histogram correlation   pearsonr(np.histogram(img0, bins=256)[0], np.histogram(img1, bins=256)[0]), NaN on a
//                           flat histogram                                                                      (:65-70)
// Pass 1 (tile_sums_kernel): one read of both planes, fp64 running sums (x, y, xx, yy, xy, (x-y)^2) and min / max.
// Pass 2 (tile_hist_kernel): np.histogram's uniform-bin rule repeated operation by operation in float32 (first guess
// ((x - min) / (max - min)) * 256 truncated, corrected by at most one bin against float32 edges k*step + min), counted in
// per-warp shared-memory histograms.  Integer counts are bit-exact against NumPy >= 2.
// Pass 3 (tile_hist_corr_kernel): one warp per tile, fp64 Pearson of the two 256-bin count vectors.
#include "ctk_common.h"

namespace {

constexpr int kSlices = 8;      // CTAs per tile
constexpr int kThreads = 256;
constexpr int kPartial = 8;     // sx, sy, sxx, syy, sxy, sdd, (min0,max0), (min1,max1)
constexpr int kBins = 256;

struct Acc {
  double sx, sy, sxx, syy, sxy, sdd;
  float mn0, mx0, mn1, mx1;
};

__device__ __forceinline__ void accumulate(Acc& a, const float4& x, const float4& y) {
  const float xs[4] = {x.x, x.y, x.z, x.w};
  const float ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double dx = static_cast<double>(xs[i]);
    const double dy = static_cast<double>(ys[i]);
    const double d = static_cast<double>(__fsub_rn(xs[i], ys[i]));     // the reference subtracts in float32
    a.sx += dx;
    a.sy += dy;
    a.sxx = fma(dx, dx, a.sxx);
    a.syy = fma(dy, dy, a.syy);
    a.sxy = fma(dx, dy, a.sxy);
    a.sdd = fma(d, d, a.sdd);
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

__global__ void __launch_bounds__(kThreads) tile_sums_kernel(const float* __restrict__ tiles, int plane_elems,
                                                             double* __restrict__ partial) {
  const int tile = blockIdx.x / kSlices;
  const int slice = blockIdx.x % kSlices;
  const float4* p0 = reinterpret_cast<const float4*>(tiles + static_cast<size_t>(tile) * 2 * plane_elems);
  const float4* p1 = p0 + plane_elems / 4;
  const int nvec = plane_elems / 4;
  const int per_slice = (nvec + kSlices - 1) / kSlices;
  const int begin = slice * per_slice;
  const int end = min(nvec, begin + per_slice);
  Acc a = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, INFINITY, -INFINITY, INFINITY, -INFINITY};
  int i = begin + threadIdx.x;
  for (; i + 3 * kThreads < end; i += 4 * kThreads) {
    float4 x[4], y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = __ldg(p0 + i + j * kThreads);      // default caching: pass 2 re-reads from L2
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = __ldg(p1 + i + j * kThreads);
#pragma unroll
    for (int j = 0; j < 4; ++j) accumulate(a, x[j], y[j]);
  }
  for (; i < end; i += kThreads) accumulate(a, __ldg(p0 + i), __ldg(p1 + i));
  a.sx = warp_sum(a.sx); a.sy = warp_sum(a.sy); a.sxx = warp_sum(a.sxx);
  a.syy = warp_sum(a.syy); a.sxy = warp_sum(a.sxy); a.sdd = warp_sum(a.sdd);
  a.mn0 = warp_min(a.mn0); a.mx0 = warp_max(a.mx0); a.mn1 = warp_min(a.mn1); a.mx1 = warp_max(a.mx1);
  __shared__ double sh[kThreads / 32][6];
  __shared__ float shm[kThreads / 32][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh[warp][0] = a.sx; sh[warp][1] = a.sy; sh[warp][2] = a.sxx; sh[warp][3] = a.syy; sh[warp][4] = a.sxy; sh[warp][5] = a.sdd;
    shm[warp][0] = a.mn0; shm[warp][1] = a.mx0; shm[warp][2] = a.mn1; shm[warp][3] = a.mx1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s[6] = {0, 0, 0, 0, 0, 0};
    float m[4] = {INFINITY, -INFINITY, INFINITY, -INFINITY};
    for (int w = 0; w < kThreads / 32; ++w) {
      for (int k = 0; k < 6; ++k) s[k] += sh[w][k];
      m[0] = fminf(m[0], shm[w][0]); m[1] = fmaxf(m[1], shm[w][1]);
      m[2] = fminf(m[2], shm[w][2]); m[3] = fmaxf(m[3], shm[w][3]);
    }
    double* out = partial + (static_cast<size_t>(tile) * kSlices + slice) * kPartial;
    for (int k = 0; k < 6; ++k) out[k] = s[k];
    out[6] = __hiloint2double(__float_as_int(m[1]), __float_as_int(m[0]));
    out[7] = __hiloint2double(__float_as_int(m[3]), __float_as_int(m[2]));
  }
}

// r, rmse and the per-plane (min, max) the histogram pass needs
__global__ void tile_finalize_kernel(const double* __restrict__ partial, int n_tiles, int plane_elems,
                                     double* __restrict__ r_out, float* __restrict__ rmse_out,
                                     float* __restrict__ minmax) {
  const int tile = blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= n_tiles) return;
  double s[6] = {0, 0, 0, 0, 0, 0};
  float mn0 = INFINITY, mx0 = -INFINITY, mn1 = INFINITY, mx1 = -INFINITY;
  for (int sl = 0; sl < kSlices; ++sl) {
    const double* p = partial + (static_cast<size_t>(tile) * kSlices + sl) * kPartial;
    for (int k = 0; k < 6; ++k) s[k] += p[k];
    mn0 = fminf(mn0, __int_as_float(__double2loint(p[6])));
    mx0 = fmaxf(mx0, __int_as_float(__double2hiint(p[6])));
    mn1 = fminf(mn1, __int_as_float(__double2loint(p[7])));
    mx1 = fmaxf(mx1, __int_as_float(__double2hiint(p[7])));
  }
  const double n = static_cast<double>(plane_elems);
  double r;
  if (!(mx0 > mn0) || !(mx1 > mn1)) {
    r = __longlong_as_double(0x7ff8000000000000ll);   // constant plane -> NaN (test-cross-talk-model.py:61-62)
  } else {
    const double cxy = s[4] - s[0] * s[1] / n;
    const double cxx = s[2] - s[0] * s[0] / n;
    const double cyy = s[3] - s[1] * s[1] / n;
    r = fmin(1.0, fmax(-1.0, cxy / sqrt(cxx * cyy)));
  }
  if (r_out) r_out[tile] = r;
  if (rmse_out) rmse_out[tile] = static_cast<float>(sqrt(s[5] / n));
  minmax[tile * 4 + 0] = mn0; minmax[tile * 4 + 1] = mx0; minmax[tile * 4 + 2] = mn1; minmax[tile * 4 + 3] = mx1;
}

// np.histogram's bin of x for 256 uniform bins over [first, last], every operation in float32 with round-to-nearest and
// no contraction (numpy/lib/_histograms_impl.py; restated and pinned in oracle/crosstalk_oracle.py::histogram256_f32)
struct BinRule {
  float first, last, step, denom;
};
__device__ __forceinline__ BinRule make_rule(float mn, float mx) {
  BinRule b;
  b.first = mn; b.last = mx;
  if (mn == mx) { b.first = __fsub_rn(mn, 0.5f); b.last = __fadd_rn(mx, 0.5f); }
  b.denom = __fsub_rn(b.last, b.first);
  b.step = __fdiv_rn(b.denom, 256.0f);
  return b;
}
__device__ __forceinline__ float edge(const BinRule& b, int k) {
  return k == kBins ? b.last : __fadd_rn(__fmul_rn(static_cast<float>(k), b.step), b.first);
}
__device__ __forceinline__ int bin_of(const BinRule& b, float x) {
  int idx = static_cast<int>(__fmul_rn(__fdiv_rn(__fsub_rn(x, b.first), b.denom), 256.0f));    // truncation
  if (idx == kBins) idx = kBins - 1;
  if (x < edge(b, idx)) idx -= 1;
  if (x >= edge(b, idx + 1) && idx != kBins - 1) idx += 1;
  return idx;
}

__global__ void __launch_bounds__(kThreads) tile_hist_kernel(const float* __restrict__ tiles, int plane_elems,
                                                             const float* __restrict__ minmax,
                                                             unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[kThreads / 32][2][kBins];      // per-warp histograms: 16 KB
  const int tile = blockIdx.x / kSlices;
  const int slice = blockIdx.x % kSlices;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (kThreads / 32) * 2 * kBins; i += kThreads) (&sh[0][0][0])[i] = 0u;
  __syncthreads();
  const BinRule r0 = make_rule(minmax[tile * 4 + 0], minmax[tile * 4 + 1]);
  const BinRule r1 = make_rule(minmax[tile * 4 + 2], minmax[tile * 4 + 3]);
  const float4* p0 = reinterpret_cast<const float4*>(tiles + static_cast<size_t>(tile) * 2 * plane_elems);
  const float4* p1 = p0 + plane_elems / 4;
  const int nvec = plane_elems / 4;
  const int per_slice = (nvec + kSlices - 1) / kSlices;
  const int begin = slice * per_slice;
  const int end = min(nvec, begin + per_slice);
  for (int i = begin + threadIdx.x; i < end; i += kThreads) {
    const float4 x = __ldcs(p0 + i), y = __ldcs(p1 + i);
    atomicAdd(&sh[warp][0][bin_of(r0, x.x)], 1u); atomicAdd(&sh[warp][0][bin_of(r0, x.y)], 1u);
    atomicAdd(&sh[warp][0][bin_of(r0, x.z)], 1u); atomicAdd(&sh[warp][0][bin_of(r0, x.w)], 1u);
    atomicAdd(&sh[warp][1][bin_of(r1, y.x)], 1u); atomicAdd(&sh[warp][1][bin_of(r1, y.y)], 1u);
    atomicAdd(&sh[warp][1][bin_of(r1, y.z)], 1u); atomicAdd(&sh[warp][1][bin_of(r1, y.w)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * kBins; i += kThreads) {
    unsigned int c = 0;
    for (int w = 0; w < kThreads / 32; ++w) c += (&sh[w][0][0])[i];
    if (c) atomicAdd(hist + static_cast<size_t>(tile) * 2 * kBins + i, c);
  }
}

__global__ void __launch_bounds__(128) tile_hist_corr_kernel(const unsigned int* __restrict__ hist, int n_tiles,
                                                             double* __restrict__ out) {
  const int tile = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tile >= n_tiles) return;
  const unsigned int* h0 = hist + static_cast<size_t>(tile) * 2 * kBins;
  const unsigned int* h1 = h0 + kBins;
  double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
  unsigned int mn0 = 0xffffffffu, mx0 = 0u, mn1 = 0xffffffffu, mx1 = 0u;
  for (int i = lane; i < kBins; i += 32) {
    const unsigned int a = h0[i], b = h1[i];
    const double x = static_cast<double>(a), y = static_cast<double>(b);
    sx += x; sy += y; sxx += x * x; syy += y * y; sxy += x * y;      // integers < 2^53: exact
    mn0 = min(mn0, a); mx0 = max(mx0, a); mn1 = min(mn1, b); mx1 = max(mx1, b);
  }
  sx = warp_sum(sx); sy = warp_sum(sy); sxx = warp_sum(sxx); syy = warp_sum(syy); sxy = warp_sum(sxy);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn0 = min(mn0, __shfl_xor_sync(0xffffffffu, mn0, o)); mx0 = max(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
    mn1 = min(mn1, __shfl_xor_sync(0xffffffffu, mn1, o)); mx1 = max(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
  }
  if (lane == 0) {
    double r;
    if (mn0 == mx0 || mn1 == mx1) {
      r = __longlong_as_double(0x7ff8000000000000ll);   // np.std(hist) == 0 -> NaN (test-cross-talk-model.py:67-68)
    } else {
      const double n = static_cast<double>(kBins);
      const double cxy = sxy - sx * sy / n, cxx = sxx - sx * sx / n, cyy = syy - sy * sy / n;
      r = fmin(1.0, fmax(-1.0, cxy / sqrt(cxx * cyy)));
    }
    out[tile] = r;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Normalised mutual information of the two planes after np.digitize into 256 levels (test-cross-talk-model.py:71-74,84;
// sklearn.metrics.normalized_mutual_info_score).  label(x) = number of float32 edges k*step + min (last edge = max) that
// are <= x, exactly as np.digitize(x, np.linspace(min, max, 256)) computes it in NumPy >= 2.
struct DigitRule {
  float lo, hi, step;
};
__device__ __forceinline__ DigitRule make_digit_rule(float mn, float mx) {
  DigitRule d;
  d.lo = mn; d.hi = mx;
  d.step = __fdiv_rn(__fsub_rn(mx, mn), 255.0f);
  return d;
}
__device__ __forceinline__ float digit_edge(const DigitRule& d, int k) {
  return k == 255 ? d.hi : __fadd_rn(__fmul_rn(static_cast<float>(k), d.step), d.lo);
}
__device__ __forceinline__ int digit_of(const DigitRule& d, float x) {      // 0-based label (NumPy's minus one)
  int k = 255;
  if (d.step > 0.f) {
    k = static_cast<int>(__fdiv_rn(__fsub_rn(x, d.lo), d.step));
    k = min(max(k, 0), 255);
  }
  // largest k with edge_k <= x.  One step each way for ordinary planes; planes whose range spans fewer than 256 float32
  // values have runs of equal edges, which the guess can miss by more than one.
  while (k > 0 && digit_edge(d, k) > x) k -= 1;
  while (k < 255 && digit_edge(d, k + 1) <= x) k += 1;
  return k;
}

__global__ void __launch_bounds__(kThreads) tile_joint_kernel(const float* __restrict__ tiles, int plane_elems,
                                                              const float* __restrict__ minmax,
                                                              unsigned int* __restrict__ joint) {
  const int tile = blockIdx.x / kSlices;
  const int slice = blockIdx.x % kSlices;
  const DigitRule r0 = make_digit_rule(minmax[tile * 4 + 0], minmax[tile * 4 + 1]);
  const DigitRule r1 = make_digit_rule(minmax[tile * 4 + 2], minmax[tile * 4 + 3]);
  const float4* p0 = reinterpret_cast<const float4*>(tiles + static_cast<size_t>(tile) * 2 * plane_elems);
  const float4* p1 = p0 + plane_elems / 4;
  unsigned int* jt = joint + static_cast<size_t>(tile) * kBins * kBins;
  const int nvec = plane_elems / 4;
  const int per_slice = (nvec + kSlices - 1) / kSlices;
  const int begin = slice * per_slice;
  const int end = min(nvec, begin + per_slice);
  for (int i = begin + threadIdx.x; i < end; i += kThreads) {
    const float4 x = __ldcs(p0 + i), y = __ldcs(p1 + i);
    atomicAdd(jt + digit_of(r0, x.x) * kBins + digit_of(r1, y.x), 1u);
    atomicAdd(jt + digit_of(r0, x.y) * kBins + digit_of(r1, y.y), 1u);
    atomicAdd(jt + digit_of(r0, x.z) * kBins + digit_of(r1, y.z), 1u);
    atomicAdd(jt + digit_of(r0, x.w) * kBins + digit_of(r1, y.w), 1u);
  }
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kThreads / 32; ++w) t += sh[w];
  return t;
}

// one CTA per tile: marginals, MI and the two entropies from the 256 x 256 contingency table (fp64, natural logs)
__global__ void __launch_bounds__(kThreads) tile_nmi_kernel(const unsigned int* __restrict__ joint, int plane_elems,
                                                            double* __restrict__ out) {
  __shared__ unsigned int row_sum[kBins], col_sum[kBins];
  __shared__ double sh[kThreads / 32];
  const int tile = blockIdx.x;
  const unsigned int* jt = joint + static_cast<size_t>(tile) * kBins * kBins;
  for (int i = threadIdx.x; i < kBins; i += kThreads) { row_sum[i] = 0u; col_sum[i] = 0u; }
  __syncthreads();
  for (int i = threadIdx.x; i < kBins; i += kThreads) {          // thread i sums row i and column i
    unsigned int r = 0u, c = 0u;
    for (int j = 0; j < kBins; ++j) { r += jt[i * kBins + j]; c += jt[j * kBins + i]; }
    row_sum[i] = r; col_sum[i] = c;
  }
  __syncthreads();
  const double n = static_cast<double>(plane_elems);
  const double log_n = log(n);
  double mi = 0.0;
  for (int idx = threadIdx.x; idx < kBins * kBins; idx += kThreads) {
    const unsigned int c = jt[idx];
    if (c == 0u) continue;
    const int i = idx / kBins, j = idx % kBins;
    const double nz = static_cast<double>(c);
    const double outer = static_cast<double>(static_cast<unsigned long long>(row_sum[i]) * col_sum[j]);
    const double log_outer = -log(outer) + log_n + log_n;
    double t = (nz / n) * (log(nz) - log_n) + (nz / n) * log_outer;
    if (fabs(t) < 2.220446049250313e-16) t = 0.0;
    mi += t;
  }
  mi = block_sum(mi, sh);
  double hr = 0.0, hc = 0.0;
  int nr = 0, nc = 0;
  for (int i = threadIdx.x; i < kBins; i += kThreads) {
    if (row_sum[i]) { const double c = static_cast<double>(row_sum[i]); hr -= (c / n) * (log(c) - log_n); nr = 1; }
    if (col_sum[i]) { const double c = static_cast<double>(col_sum[i]); hc -= (c / n) * (log(c) - log_n); nc = 1; }
  }
  hr = block_sum(hr, sh);
  hc = block_sum(hc, sh);
  const int rows_used = __syncthreads_count(nr), cols_used = __syncthreads_count(nc);
  if (threadIdx.x == 0) {
    double r;
    if (rows_used == 1 && cols_used == 1) r = 1.0;                  // both labelings constant
    else if (rows_used == 1 || cols_used == 1) r = 0.0;
    else {
      mi = fmax(mi, 0.0);
      r = fabs(mi) < 2.220446049250313e-16 ? 0.0 : mi / (0.5 * (hr + hc));
    }
    out[tile] = r;
  }
}

}  // namespace

extern "C" {

size_t ctk_tile_metrics_workspace_bytes(int n_tiles) {
  if (n_tiles <= 0) return 0;
  return static_cast<size_t>(n_tiles) * (kSlices * kPartial * sizeof(double) + 4 * sizeof(float));
}

int ctk_tile_metrics_f32(const float* tiles, int n_tiles, int plane_elems, double* pearson_out, float* rmse_out,
                         double* hist_corr_out, unsigned int* hist_out, void* workspace, size_t workspace_bytes,
                         void* stream) {
  if (n_tiles == 0) return CTK_OK;
  CTK_REQUIRE(tiles && workspace && n_tiles > 0 && plane_elems > 0 && plane_elems % 4 == 0);
  CTK_REQUIRE((hist_corr_out == nullptr) || (hist_out != nullptr));
  CTK_REQUIRE((reinterpret_cast<uintptr_t>(tiles) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0);
  if (workspace_bytes < ctk_tile_metrics_workspace_bytes(n_tiles)) return CTK_ERR_WORKSPACE;
  cudaStream_t s = ctk::as_stream(stream);
  double* partial = static_cast<double*>(workspace);
  float* minmax = reinterpret_cast<float*>(partial + static_cast<size_t>(n_tiles) * kSlices * kPartial);
  tile_sums_kernel<<<static_cast<unsigned>(n_tiles) * kSlices, kThreads, 0, s>>>(tiles, plane_elems, partial);
  int st = ctk::check_launch();
  if (st != CTK_OK) return st;
  tile_finalize_kernel<<<(n_tiles + 127) / 128, 128, 0, s>>>(partial, n_tiles, plane_elems, pearson_out, rmse_out, minmax);
  st = ctk::check_launch();
  if (st != CTK_OK || hist_out == nullptr) return st;
  CTK_CUDA_TRY(cudaMemsetAsync(hist_out, 0, sizeof(unsigned int) * 2 * kBins * static_cast<size_t>(n_tiles), s));
  tile_hist_kernel<<<static_cast<unsigned>(n_tiles) * kSlices, kThreads, 0, s>>>(tiles, plane_elems, minmax, hist_out);
  st = ctk::check_launch();
  if (st != CTK_OK || hist_corr_out == nullptr) return st;
  tile_hist_corr_kernel<<<(n_tiles + 3) / 4, 128, 0, s>>>(hist_out, n_tiles, hist_corr_out);
  return ctk::check_launch();
}

size_t ctk_tile_nmi_workspace_bytes(int n_tiles) {
  if (n_tiles <= 0) return 0;
  return ctk_tile_metrics_workspace_bytes(n_tiles) + static_cast<size_t>(n_tiles) * kBins * kBins * sizeof(unsigned int);
}

int ctk_tile_nmi_f32(const float* tiles, int n_tiles, int plane_elems, double* nmi_out, void* workspace,
                     size_t workspace_bytes, void* stream) {
  if (n_tiles == 0) return CTK_OK;
  CTK_REQUIRE(tiles && nmi_out && workspace && n_tiles > 0 && plane_elems > 0 && plane_elems % 4 == 0);
  CTK_REQUIRE((reinterpret_cast<uintptr_t>(tiles) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0);
  if (workspace_bytes < ctk_tile_nmi_workspace_bytes(n_tiles)) return CTK_ERR_WORKSPACE;
  cudaStream_t s = ctk::as_stream(stream);
  double* partial = static_cast<double*>(workspace);
  float* minmax = reinterpret_cast<float*>(partial + static_cast<size_t>(n_tiles) * kSlices * kPartial);
  unsigned int* joint = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + ctk_tile_metrics_workspace_bytes(n_tiles));
  tile_sums_kernel<<<static_cast<unsigned>(n_tiles) * kSlices, kThreads, 0, s>>>(tiles, plane_elems, partial);
  int st = ctk::check_launch();
  if (st != CTK_OK) return st;
  tile_finalize_kernel<<<(n_tiles + 127) / 128, 128, 0, s>>>(partial, n_tiles, plane_elems, nullptr, nullptr, minmax);
  st = ctk::check_launch();
  if (st != CTK_OK) return st;
  CTK_CUDA_TRY(cudaMemsetAsync(joint, 0, sizeof(unsigned int) * kBins * kBins * static_cast<size_t>(n_tiles), s));
  tile_joint_kernel<<<static_cast<unsigned>(n_tiles) * kSlices, kThreads, 0, s>>>(tiles, plane_elems, minmax, joint);
  st = ctk::check_launch();
  if (st != CTK_OK) return st;
  tile_nmi_kernel<<<static_cast<unsigned>(n_tiles), kThreads, 0, s>>>(joint, plane_elems, nmi_out);
  return ctk::check_launch();
}

}  // extern "C"
