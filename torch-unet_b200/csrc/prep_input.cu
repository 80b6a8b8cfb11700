// Input pipeline on the device: raw TIFF pixel payloads -> the [N,2,H,W] float32 batch the models
// take.  Replaces, per plane, the host code of /root/reference/train_model.py:166-167 (iio.imread(...).astype(np.float32)),
// :211-216 (normalize_image: (img - min) / (max - min) in float32, unchanged if the plane is constant) and the flips of
// :225-232 (TF.hflip / TF.vflip applied to both planes of a sample).
//
// One 4-CTA cluster per plane: every thread keeps its 64 pixels in registers, the plane's min / max are combined across
// the cluster through distributed shared memory, and the normalised pixels are written once -- the plane is read from HBM
// exactly once (8 B or 4 B per pixel in, 4 B out).  The arithmetic is IEEE float32 with round-to-nearest (__fsub_rn /
// __fdiv_rn), i.e. bit-identical to NumPy's.
#include "ctk_common.h"
#include "ctk_ptx.cuh"

#include <cooperative_groups.h>

namespace {

namespace cg = cooperative_groups;

constexpr int kCluster = 4;
constexpr int kThreads = 256;
constexpr int kVpt = 64;                          // pixels per thread: 4 CTAs x 256 threads x 64 = 65 536 = 256 x 256

template <typename T>
__device__ __forceinline__ float to_f32(T v) { return static_cast<float>(v); }     // cvt.rn.f32.f64 == astype(np.float32)

__device__ __forceinline__ float block_reduce(float v, bool want_max, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = want_max ? fmaxf(v, u) : fminf(v, u);
  }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int w = 1; w < kThreads / 32; ++w) r = want_max ? fmaxf(r, sh[w]) : fminf(r, sh[w]);
  __syncthreads();
  return r;
}

// plane p = blockIdx.x / kCluster of [n][2] planes; flags[n]: bit 0 = horizontal flip, bit 1 = vertical flip
template <typename T>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads)
prepare_tiles_cluster_kernel(const T* __restrict__ raw, const unsigned char* __restrict__ flags, int H, int W,
                             float* __restrict__ out) {
  __shared__ float sh[kThreads / 32];
  __shared__ float s_minmax[2];
  cg::cluster_group cluster = cg::this_cluster();
  const int plane = blockIdx.x / kCluster;
  const int rank = static_cast<int>(cluster.block_rank());
  const size_t plane_elems = static_cast<size_t>(H) * W;
  const T* src = raw + static_cast<size_t>(plane) * plane_elems;
  // thread t of CTA r owns elements (r * 256 + t) * 4 + {0..3} + j * 4096, j = 0..15: 16-byte (f32) / 32-byte (f64) loads
  float v[kVpt];
  const int base = (rank * kThreads + threadIdx.x) * 4;
#pragma unroll
  for (int j = 0; j < kVpt / 4; ++j) {
    const size_t i = static_cast<size_t>(base) + static_cast<size_t>(j) * (kCluster * kThreads * 4);
    if constexpr (sizeof(T) == 8) {
      const double2 a = __ldcs(reinterpret_cast<const double2*>(src + i));
      const double2 b = __ldcs(reinterpret_cast<const double2*>(src + i + 2));
      v[4 * j] = to_f32(a.x); v[4 * j + 1] = to_f32(a.y); v[4 * j + 2] = to_f32(b.x); v[4 * j + 3] = to_f32(b.y);
    } else {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(src + i));
      v[4 * j] = a.x; v[4 * j + 1] = a.y; v[4 * j + 2] = a.z; v[4 * j + 3] = a.w;
    }
  }
  float mn = v[0], mx = v[0];
#pragma unroll
  for (int k = 1; k < kVpt; ++k) { mn = fminf(mn, v[k]); mx = fmaxf(mx, v[k]); }
  mn = block_reduce(mn, false, sh);
  mx = block_reduce(mx, true, sh);
  if (threadIdx.x == 0) { s_minmax[0] = mn; s_minmax[1] = mx; }
  cluster.sync();
  for (int r = 0; r < kCluster; ++r) {
    const float* peer = cluster.map_shared_rank(s_minmax, r);
    mn = fminf(mn, peer[0]);
    mx = fmaxf(mx, peer[1]);
  }
  cluster.sync();                                   // nobody exits while a peer still reads its shared memory
  const bool scale = mx > mn;                       // normalize_image leaves a constant plane unchanged
  const float denom = __fsub_rn(mx, mn);
  const unsigned char f = flags ? flags[plane >> 1] : 0;
  float* dst = out + static_cast<size_t>(plane) * plane_elems;
#pragma unroll
  for (int j = 0; j < kVpt / 4; ++j) {
    const int i = base + j * (kCluster * kThreads * 4);
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = scale ? __fdiv_rn(__fsub_rn(v[4 * j + k], mn), denom) : v[4 * j + k];
    int y = i / W, x = i - y * W;                   // W % 4 == 0: the four pixels share a row
    if (f & 2) y = H - 1 - y;
    if (f & 1) {
      x = W - 4 - x;
      *reinterpret_cast<float4*>(dst + static_cast<size_t>(y) * W + x) = make_float4(o[3], o[2], o[1], o[0]);
    } else {
      *reinterpret_cast<float4*>(dst + static_cast<size_t>(y) * W + x) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

// any plane size: one CTA per plane, second pass re-reads the plane (L2)
template <typename T>
__global__ void __launch_bounds__(kThreads)
prepare_tiles_generic_kernel(const T* __restrict__ raw, const unsigned char* __restrict__ flags, int H, int W,
                             float* __restrict__ out) {
  __shared__ float sh[kThreads / 32];
  const int plane = blockIdx.x;
  const int plane_elems = H * W;
  const T* src = raw + static_cast<size_t>(plane) * plane_elems;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < plane_elems; i += kThreads) {
    const float x = to_f32(src[i]);
    mn = fminf(mn, x);
    mx = fmaxf(mx, x);
  }
  mn = block_reduce(mn, false, sh);
  mx = block_reduce(mx, true, sh);
  const bool scale = mx > mn;
  const float denom = __fsub_rn(mx, mn);
  const unsigned char f = flags ? flags[plane >> 1] : 0;
  float* dst = out + static_cast<size_t>(plane) * plane_elems;
  for (int i = threadIdx.x; i < plane_elems; i += kThreads) {
    const float x = to_f32(src[i]);
    int yy = i / W, xx = i - yy * W;
    if (f & 2) yy = H - 1 - yy;
    if (f & 1) xx = W - 1 - xx;
    dst[static_cast<size_t>(yy) * W + xx] = scale ? __fdiv_rn(__fsub_rn(x, mn), denom) : x;
  }
}

template <typename T>
int launch_prepare(const T* raw, const unsigned char* flags, int n, int H, int W, float* out, cudaStream_t s) {
  const long long plane_elems = static_cast<long long>(H) * W;
  if (plane_elems == kCluster * kThreads * kVpt && W % 4 == 0 && (reinterpret_cast<uintptr_t>(raw) & 31) == 0) {
    prepare_tiles_cluster_kernel<T><<<static_cast<unsigned>(n) * 2 * kCluster, kThreads, 0, s>>>(raw, flags, H, W, out);
  } else {
    prepare_tiles_generic_kernel<T><<<static_cast<unsigned>(n) * 2, kThreads, 0, s>>>(raw, flags, H, W, out);
  }
  return ctk::check_launch();
}

}  // namespace

extern "C" int ctk_prepare_tiles(const void* raw, int raw_is_f64, const unsigned char* flip_flags, int n, int H, int W,
                                 float* out, void* stream) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(raw && out && n > 0 && H > 0 && W > 0 && static_cast<long long>(H) * W < (1ll << 30));
  CTK_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(raw) & 7) == 0);
  cudaStream_t s = ctk::as_stream(stream);
  if (raw_is_f64) return launch_prepare(static_cast<const double*>(raw), flip_flags, n, H, W, out, s);
  return launch_prepare(static_cast<const float*>(raw), flip_flags, n, H, W, out, s);
}
