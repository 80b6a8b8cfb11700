// Train-mode forward statistics and weight gradient of the FIRST conv block without ever materialising its
// full-resolution output (1 G elements per branch at batch 256 -- more than half of all pre-pool activations).
//
// The first conv has only T = 9*Cin taps, so everything that is linear or quadratic in its output follows from the
// patch Gram matrix of the input,  S[t] = sum_p x[p+t],  G[t][t'] = sum_p x[p+t] x[p+t']  (zero padded like the conv):
//   batch mean / variance of channel c :  mean = w_c.S / M,   var = w_c^T (G/M - s s^T) w_c            (exact, fp64)
//   sum_p Y[p,c] x[p+t]                :  (G w_c)[t]
// Forward therefore is: Gram -> moments -> BN constants -> the fused eval-style kernel (conv + scale/shift + LeakyReLU +
// pool) that writes only the pooled output.  Backward needs dW[c,t] = sum_p dY[p,c] x[p+t] with
//   dY = scale_c (dA - m1_c - xhat m2_c)   =>   dW[c,t] = scale_c [ T1[c,t] - m1_c S_t - m2_c invstd_c ((G w_c)[t] - mu_c S_t) ]
// where only T1[c,t] = sum_windows dP[w,c] f'(z*) x[p*(w,c)+t] touches the data: a gather driven by the 4-bit
// arg-max / sign codes the forward kernel stored per pooled element.
// Replaces (train mode) nn.Conv2d + nn.BatchNorm2d statistics + their backward for the first block:
// /root/reference/regression_model.py:14-15 and two_branch_regression.py:10-11.
#include <cstdlib>

#include "ctk_common.h"
#include "ctk_ptx.cuh"

namespace {

using namespace ctk;

// ------------------------------------------------------------------------------------------------ patch Gram matrix
// gram layout (double): [0, T) = S, [T, T + T*T) = G row-major (full, symmetric).
// Items = T sums + T(T+1)/2 upper-triangle products.  A warp owns one row of a 32 x 8 pixel tile at a time (lane = pixel)
// and a compile-time slice ("role") of the items, so every FFMA is unconditional: 9*CIN LDS feed ITEMS/ROLES FFMAs.
template <int CIN, int kRoles>
struct GramCfg {
  static constexpr int T = 9 * CIN;
  static constexpr int ITEMS = T + T * (T + 1) / 2;
  static constexpr int ROLES = kRoles;
  static constexpr int PER = (ITEMS + ROLES - 1) / ROLES;
  static constexpr int ROWG = 8 / ROLES;             // warps that share a role split the tile's rows
  static constexpr int TW = 32, TH = 64;             // pixel tile per iteration: long enough to hide the next tile's loads
  static constexpr int SLOTS = ((TH + 2) * (TW + 2) + 255) / 256;   // staging elements per thread and channel
};

template <int CIN, int kRoles, int ROLE>
__device__ __forceinline__ void gram_accumulate(const float (&v)[9 * CIN], float (&acc)[GramCfg<CIN, kRoles>::PER]) {
  using C = GramCfg<CIN, kRoles>;
  int item = 0;
#pragma unroll
  for (int a = 0; a < C::T; ++a, ++item)
    if (item % C::ROLES == ROLE) acc[item / C::ROLES] += v[a];
#pragma unroll
  for (int a = 0; a < C::T; ++a)
#pragma unroll
    for (int b = a; b < C::T; ++b, ++item)
      if (item % C::ROLES == ROLE) acc[item / C::ROLES] = fmaf(v[a], v[b], acc[item / C::ROLES]);
}

template <int CIN, int kRoles, int ROLE>
__device__ __forceinline__ void gram_role(const float* __restrict__ x, int n_img, int c_total, int c_offset, int H, int W,
                                          double* __restrict__ part, float (*s_in)[GramCfg<CIN, kRoles>::TH + 2][35],
                                          double* s_red) {
  using C = GramCfg<CIN, kRoles>;
  constexpr int T = C::T, TW = C::TW, TH = C::TH, SLOTS = C::SLOTS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rowg = warp / C::ROLES;
  // fp32 partial sums (<= TH / ROWG adds each), folded into fp64 running sums after every tile: the variance
  // w^T (G/M - s s^T) w cancels one or two digits, so the Gram entries need more than fp32 over 10^7 pixels
  float acc[C::PER];
  double dacc[C::PER];
#pragma unroll
  for (int i = 0; i < C::PER; ++i) { acc[i] = 0.f; dacc[i] = 0.0; }
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  const long long total = static_cast<long long>(n_img) * tiles_x * tiles_y;
  // this thread's staging slots of the (TH+2) x (TW+2) halo
  int sr[SLOTS], sq[SLOTS];
#pragma unroll
  for (int k = 0; k < SLOTS; ++k) {
    const int i = threadIdx.x + 256 * k;
    sr[k] = i / (TW + 2);
    sq[k] = i - sr[k] * (TW + 2);
  }
  // the next tile's halo is fetched into registers while the current one is being multiplied
  float pf[CIN][SLOTS];
  auto fetch = [&](long long tile) {
    const int tx = static_cast<int>(tile % tiles_x);
    const int ty = static_cast<int>((tile / tiles_x) % tiles_y);
    const int img = static_cast<int>(tile / (static_cast<long long>(tiles_x) * tiles_y));
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float* plane = x + (static_cast<size_t>(img) * c_total + c_offset + c) * H * W;
#pragma unroll
      for (int k = 0; k < SLOTS; ++k) {
        const int gy = ty * TH - 1 + sr[k], gx = tx * TW - 1 + sq[k];
        pf[c][k] = (sr[k] < TH + 2 && gy >= 0 && gy < H && gx >= 0 && gx < W)
                       ? __ldg(plane + static_cast<size_t>(gy) * W + gx) : 0.f;
      }
    }
  };
  if (blockIdx.x < total) fetch(blockIdx.x);
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tx = static_cast<int>(tile % tiles_x);
    const int ty = static_cast<int>((tile / tiles_x) % tiles_y);
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CIN; ++c)
#pragma unroll
      for (int k = 0; k < SLOTS; ++k)
        if (sr[k] < TH + 2) s_in[c][sr[k]][sq[k]] = pf[c][k];
    __syncthreads();
    if (tile + gridDim.x < total) fetch(tile + gridDim.x);
    // explicit shared-space loads with immediate offsets: through the generic pointer every one of the 9*CIN loads of a
    // row cost an extra shared-window address computation (ncu: 525 M instructions for 198 M FFMAs at CIN = 2)
    const uint32_t s_base = static_cast<uint32_t>(__cvta_generic_to_shared(&s_in[0][0][0])) + static_cast<uint32_t>(lane * 4);
#pragma unroll 2
    for (int py = rowg; py < TH; py += C::ROWG) {
      const int gy = ty * TH + py, gx = tx * TW + lane;
      float v[T];
      const bool in = gy < H && gx < W;
      const uint32_t row_addr = s_base + static_cast<uint32_t>(py * 35 * 4);
#pragma unroll
      for (int c = 0; c < CIN; ++c)
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          float t;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(row_addr + static_cast<uint32_t>((c * (TH + 2) * 35 + (k / 3) * 35 + (k % 3)) * 4)));
          v[c * 9 + k] = in ? t : 0.f;
        }
      gram_accumulate<CIN, kRoles, ROLE>(v, acc);
    }
#pragma unroll
    for (int i = 0; i < C::PER; ++i) { dacc[i] += static_cast<double>(acc[i]); acc[i] = 0.f; }
  }
  // lanes -> warp total, the ROWG warps of a role -> CTA total (shared memory), one fp64 partial per item per CTA
#pragma unroll
  for (int i = 0; i < C::PER; ++i) {
    double sres = dacc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sres += __shfl_xor_sync(0xffffffffu, sres, o);
    if (lane == 0) s_red[warp * C::PER + i] = sres;
  }
  __syncthreads();
  if (rowg == 0) {
    for (int i = lane; i < C::PER; i += 32) {
      const int item = i * C::ROLES + ROLE;
      if (item >= C::ITEMS) continue;
      double tot = 0.0;
      for (int g = 0; g < C::ROWG; ++g) tot += s_red[(g * C::ROLES + ROLE) * C::PER + i];
      part[static_cast<size_t>(blockIdx.x) * C::ITEMS + item] = tot;      // one row per CTA, added up by gram_reduce_kernel
    }
  }
}

// kRoles splits the items over the warps of a CTA: fewer accumulators (and fp64 running sums) per thread buy resident
// CTAs.  With one role the Cin = 1 kernel needs ~240 registers (54 fp32 + 54 fp64 sums): 8 warps per SM, and the 9 shared-
// memory loads + 54 FFMAs of a row wait on each other with nothing else to issue; kMinBlocks tells the compiler the
// register budget of the split variants.
template <int CIN, int kRoles, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks)
patch_gram_kernel(const float* __restrict__ x, int n_img, int c_total, int c_offset, int H, int W,
                  double* __restrict__ part) {
  using C = GramCfg<CIN, kRoles>;
  static_assert(kRoles == 1 || kRoles == 2 || kRoles == 4, "roles must divide the 8 warps");
  __shared__ float s_in[CIN][C::TH + 2][35];
  __shared__ double s_red[8 * C::PER];
  const int role = (threadIdx.x >> 5) % C::ROLES;
  if constexpr (C::ROLES == 1) {
    gram_role<CIN, kRoles, 0>(x, n_img, c_total, c_offset, H, W, part, s_in, s_red);
  } else if constexpr (C::ROLES == 2) {
    if (role == 0) gram_role<CIN, kRoles, 0>(x, n_img, c_total, c_offset, H, W, part, s_in, s_red);   // warp-uniform
    else gram_role<CIN, kRoles, 1>(x, n_img, c_total, c_offset, H, W, part, s_in, s_red);
  } else {
    switch (role) {          // warp-uniform
      case 0: gram_role<CIN, kRoles, 0>(x, n_img, c_total, c_offset, H, W, part, s_in, s_red); break;
      case 1: gram_role<CIN, kRoles, 1>(x, n_img, c_total, c_offset, H, W, part, s_in, s_red); break;
      case 2: gram_role<CIN, kRoles, 2>(x, n_img, c_total, c_offset, H, W, part, s_in, s_red); break;
      default: gram_role<CIN, kRoles, 3>(x, n_img, c_total, c_offset, H, W, part, s_in, s_red); break;
    }
  }
}

// Second stage: item totals over the CTA rows in row order (one warp per item: lane l adds rows l, l+32, ..., then a
// fixed butterfly), expanded from the upper triangle to the full symmetric matrix.  Deterministic: no atomics.
__global__ void gram_reduce_kernel(const double* __restrict__ part, int rows, int T, double* __restrict__ gram) {
  const int items = T + T * (T + 1) / 2;
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (item >= items) return;
  double acc = 0.0;
  for (int r = lane; r < rows; r += 32) acc += part[static_cast<size_t>(r) * items + item];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane != 0) return;
  if (item < T) {
    gram[item] = acc;
  } else {
    int rem = item - T, a = 0;                 // upper-triangle index -> (a, b), a <= b
    while (rem >= T - a) { rem -= T - a; ++a; }
    const int b = a + rem;
    gram[T + a * T + b] = acc;
    gram[T + b * T + a] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ moments from the Gram matrix
// The CTA first forms s = S / M and the patch covariance C = G / M - s s^T in shared memory (T <= 18), then a thread per
// channel evaluates w.s and w^T C w: 2 T^2 fp64 FMAs.  (Dividing inside the T^2 loop -- three fp64 divisions per term and
// thread, straight from global memory -- made this 25 us per launch.)
__global__ void first_moments_kernel(const double* __restrict__ gram, const float* __restrict__ w, int cout, int T,
                                     double count, float* __restrict__ moments) {
  __shared__ double s_mean[18], s_cov[18 * 18];
  const double inv = 1.0 / count;
  for (int i = threadIdx.x; i < T; i += blockDim.x) s_mean[i] = gram[i] * inv;
  __syncthreads();
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) s_cov[i] = gram[T + i] * inv - s_mean[i / T] * s_mean[i % T];
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cout) return;
  double mean = 0.0;
  for (int t = 0; t < T; ++t) mean += static_cast<double>(w[c * T + t]) * s_mean[t];
  double var = 0.0;
  for (int a = 0; a < T; ++a) {
    double row = 0.0;
    for (int b = 0; b < T; ++b) row += static_cast<double>(w[c * T + b]) * s_cov[a * T + b];
    var += static_cast<double>(w[c * T + a]) * row;
  }
  moments[c] = static_cast<float>(mean);
  moments[cout + c] = static_cast<float>(var > 0.0 ? var : 0.0);
}

// ------------------------------------------------------------------------------------------------ T1 = sum_w dP f'(z*) x[p* + t]
// The forward pass (ctk_conv_first_pool_codes) left a 4-bit code per pooled element: arg-max position p* inside the 2x2
// window and the sign of the pre-activation there.  The data term of the weight gradient is then a gather:
// 9*CIN shared-memory loads and FMAs per (window, channel), no recomputation of the convolution.
//   t1[c][t] = sum_w g[w,c] x[p*(w,c) + t],   s1[c] = sum_w g[w,c],   g = dP * (sign ? slope : 1)
// (sum_w g * xhat, the BatchNorm weight gradient, follows from t1 and s1 in the finalize kernel:
//  sum_w g conv(x)[p*] = sum_t w[c][t] t1[c][t].)
template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
first_wgrad_codes_kernel(const float* __restrict__ x, int n_img, int c_total, int c_offset, int H, int W,
                         const uint32_t* __restrict__ codes, const __nv_bfloat16* __restrict__ dp, float slope,
                         float* __restrict__ part) {
  constexpr int T = 9 * CIN;
  constexpr int G = COUT / 4;                  // channel groups of 4
  constexpr int SLOTS = 256 / G;               // windows processed concurrently
  constexpr int TWW = 16, TWH = 16;            // pooled-pixel (window) tile per iteration
  constexpr int PITCH = 36;                    // == 4 (mod 32): the <= 8 distinct addresses of a warp hit distinct banks
  constexpr int ROWS = 2 * TWH + 2, COLS = 2 * TWW + 2;
  constexpr int PRE = (ROWS * COLS + 255) / 256;
  __shared__ float s_in[CIN][ROWS][PITCH];
  __shared__ float s_red[256];
  const int cg = threadIdx.x % G, slot = threadIdx.x / G;
  float acc[4][T], s1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s1[j] = 0.f;
#pragma unroll
    for (int k = 0; k < T; ++k) acc[j][k] = 0.f;
  }
  const int Hp = H >> 1, Wp = W >> 1;
  const int tiles_x = (Wp + TWW - 1) / TWW, tiles_y = (Hp + TWH - 1) / TWH;
  const long long total = static_cast<long long>(n_img) * tiles_x * tiles_y;
  int sr[PRE], sq[PRE];
#pragma unroll
  for (int k = 0; k < PRE; ++k) {
    const int i = threadIdx.x + 256 * k;
    sr[k] = i < ROWS * COLS ? i / COLS : -1000000;
    sq[k] = i - (i / COLS) * COLS;
  }
  float pf[CIN][PRE];
  auto fetch = [&](long long tile) {
    const int tx = static_cast<int>(tile % tiles_x);
    const int ty = static_cast<int>((tile / tiles_x) % tiles_y);
    const int img = static_cast<int>(tile / (static_cast<long long>(tiles_x) * tiles_y));
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float* plane = x + (static_cast<size_t>(img) * c_total + c_offset + c) * H * W;
#pragma unroll
      for (int k = 0; k < PRE; ++k) {
        const int gy = 2 * ty * TWH - 1 + sr[k], gx = 2 * tx * TWW - 1 + sq[k];
        pf[c][k] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(plane + static_cast<size_t>(gy) * W + gx) : 0.f;
      }
    }
  };
  if (blockIdx.x < total) fetch(blockIdx.x);
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tx = static_cast<int>(tile % tiles_x);
    const int ty = static_cast<int>((tile / tiles_x) % tiles_y);
    const int img = static_cast<int>(tile / (static_cast<long long>(tiles_x) * tiles_y));
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CIN; ++c)
#pragma unroll
      for (int k = 0; k < PRE; ++k)
        if (sr[k] >= 0) s_in[c][sr[k]][sq[k]] = pf[c][k];
    __syncthreads();
    if (tile + gridDim.x < total) fetch(tile + gridDim.x);
    // this thread's windows of the tile, eight at a time: all their gradient / code loads are issued before the first
    // one is used (the gather itself only touches shared memory, so these global loads are the kernel's latency)
    constexpr int PER_THREAD = TWW * TWH / SLOTS;
    constexpr int BATCH = PER_THREAD < 8 ? PER_THREAD : 8;
    static_assert(PER_THREAD % BATCH == 0, "window batches");
#pragma unroll 1
    for (int w0 = 0; w0 < PER_THREAD; w0 += BATCH) {
      uint2 raws[BATCH];
      uint32_t cws[BATCH];
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        const int win = slot + (w0 + b) * SLOTS;
        const int wy = win / TWW, wx = win % TWW;
        const int py = ty * TWH + wy, px = tx * TWW + wx;
        const bool ok = py < Hp && px < Wp;
        const size_t pix = ok ? (static_cast<size_t>(img) * Hp + py) * Wp + px : 0;
        raws[b] = ok ? __ldcs(reinterpret_cast<const uint2*>(dp + pix * COUT + cg * 4)) : make_uint2(0u, 0u);
        cws[b] = ok ? __ldcs(codes + pix * (COUT / 8) + (cg >> 1)) : 0u;
      }
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        const int win = slot + (w0 + b) * SLOTS;
        const int wy = win / TWW, wx = win % TWW;
        const uint2 raw = raws[b];
        const uint32_t cw = cws[b] >> ((cg & 1) * 16);
        const float g[4] = {__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u),
                            __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t nib = (cw >> (4 * j)) & 0xfu;
          const float gg = (nib & 4u) ? g[j] * slope : g[j];     // out-of-range windows carry g = 0
          s1[j] += gg;
          const float* base = &s_in[0][2 * wy + ((nib >> 1) & 1u)][2 * wx + (nib & 1u)];
#pragma unroll
          for (int c = 0; c < CIN; ++c)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx)
                acc[j][c * 9 + ky * 3 + kx] = fmaf(gg, base[(c * ROWS + ky) * PITCH + kx], acc[j][c * 9 + ky * 3 + kx]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k <= T; ++k) {
      __syncthreads();
      s_red[threadIdx.x] = k < T ? acc[j][k < T ? k : 0] : s1[j];
      __syncthreads();
      if (slot == 0) {
        float s = 0.f;
        for (int t = 0; t < SLOTS; ++t) s += s_red[t * G + cg];
        // one row [COUT * T (t1) | COUT (s1)] of partial sums per CTA
        float* row = part + static_cast<size_t>(blockIdx.x) * (COUT * T + COUT);
        if (k < T) row[(cg * 4 + j) * T + k] = s;
        else row[COUT * T + cg * 4 + j] = s;
      }
    }
}

// ------------------------------------------------------------------------------------------------ the same gather on tcgen05
// t1[c][t] = sum_w g[w,c] x[p*(w,c) + t] is a GEMM over K = windows once the arg-max position is folded into the A operand:
//   D[c][n] = sum_q sum_w A_q[c][w] * B_q[w][n],   A_q[c][w] = g[w,c] if position(w,c) == q else 0   (bf16: dP is bf16),
//   B_q[w][n] = the 9*CIN taps of window w's patch shifted by q, every fp32 value carried as three bf16 terms
//   (hi, mid, lo planes: exact to fp32), plus a column of ones that yields s1[c] = sum_w g[w,c].
// Both operands are MN-major (the K index -- the window -- selects a 128-byte row), built by hand in the 128B-swizzled
// layout TMA would produce: A builders mask one 16-byte chunk (8 channels of one window) per variant with two PRMT look-ups
// of the packed 2-bit positions, B builders (one thread per window) expand the 4x4 patch.  One (region of 16x8 windows,
// position) pair is one pipeline stage of 8 MMAs (M = 128 channels, N = 64 columns, K = 16 windows each).  tcgen05
// accumulates with truncation (-2^-24 per step), so the TMEM accumulator is drained into fp32 shared-memory totals every
// four regions (128 steps) instead of once per kernel (~28 000 steps).  Per-CTA partial rows as in the CUDA-core kernel.
namespace wtc {
constexpr int kWinH = 16, kWinW = 8;                 // windows per region: K = 128 per stage
constexpr int kStages = 3;
constexpr int kABytes = 32768, kBBytes = 16384, kStageBytes = kABytes + kBBytes;
constexpr int kAccPitch = 65;                        // fp32 totals [128 channels][64 columns], conflict-free rows
constexpr int kInRows = 2 * kWinH + 2, kInCols = 2 * kWinW + 2, kInPitch = 24;
// warps 0-3 B builders + drain, warp 4 MMA, then 8 A-builder warps.  (16 A-builder warps for Cout = 64 -- two chunks per
// thread, 80 registers -- were measured SLOWER: 1.49 ms against 1.02 ms per step; the builders are bound by issue slots, not
// by the number of warps hiding latency.)
constexpr int a_warps(int) { return 8; }
constexpr int threads(int cout) { return (5 + a_warps(cout)) * 32; }
constexpr int kDrainEvery = 4;                       // regions per accumulation period
constexpr int smem_bytes(int cin) {
  return 1024 + kStages * kStageBytes + 128 * kAccPitch * 4 + cin * kInRows * kInPitch * 4 + 256;
}
__device__ __forceinline__ uint64_t mn_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {   // MN-major, SWIZZLE_128B
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3fff) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t mn_idesc(uint32_t M, uint32_t N) {                    // bf16 x bf16 -> f32, A and B MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ uint32_t bf16_bits(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
// Bounded wait without the %globaltimer read of ctk::mbar_wait: the roles of this kernel hand stages to one another four
// times per region and usually find the barrier not yet flipped; a protocol bug still traps (try_wait itself suspends the
// warp for a hardware-bounded time, so 2^26 failed polls are seconds, not an endless spin).
__device__ __forceinline__ void wait(uint32_t bar_addr, uint32_t parity) {
  uint32_t n = 0;
  for (;;) {
    uint32_t ok;
    // the suspend-time hint keeps a waiting warp parked instead of polling: the builder and MMA warps share schedulers
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar_addr), "r"(parity), "r"(0x989680u)
        : "memory");
    if (ok) return;
    if (++n > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void arrive(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
}  // namespace wtc

template <int CIN, int COUT>
__global__ void __launch_bounds__(wtc::threads(COUT), 1)
first_wgrad_tc_kernel(const float* __restrict__ x, int n_img, int c_total, int c_offset, int H, int W,
                      const uint32_t* __restrict__ codes, const __nv_bfloat16* __restrict__ dp, float slope,
                      float* __restrict__ part) {
  using namespace wtc;
  constexpr int T = 9 * CIN;
  constexpr int TP = CIN == 1 ? 10 : 18;              // plane stride inside a B row (even: planes start on a word)
  constexpr int ONES = 3 * TP;                        // the column of ones
  constexpr int BWORDS = ONES / 2 + 1;                // 32-bit words of a B row that are ever non-zero
  constexpr int BCHUNKS = (BWORDS + 3) / 4;
  constexpr int CH = COUT / 8;                        // 16-byte chunks (8 channels) per window
  constexpr int NAT = a_warps(COUT) * 32;             // A-builder threads
  constexpr int kThreads = threads(COUT);
  constexpr int AIT = 128 * CH / NAT;                 // chunks per A-builder thread and variant
  constexpr int WSTEP = NAT / CH;                     // windows between a thread's chunks (a multiple of 8: whole rows)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stages = smem;
  float* acc_s = reinterpret_cast<float*>(smem + kStages * kStageBytes);
  float* s_in = acc_s + 128 * kAccPitch;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_in + CIN * kInRows * kInPitch);
  uint64_t* full = bars;                    // [kStages]: 12 builder-warp arrivals
  uint64_t* empty = bars + kStages;         // [kStages]: one MMA commit
  uint64_t* acc_full = empty + kStages;     // [2]: one MMA commit
  uint64_t* acc_empty = acc_full + 2;       // [2]: 4 drain-warp arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], a_warps(COUT) + 4); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<1>(tmem_slot, 128);
  // stage memory starts zeroed: channel block 1 of A (COUT == 64) and the unused tail of every B row stay zero for good
  for (int i = threadIdx.x; i < kStages * kStageBytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(stages)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < 128 * kAccPitch; i += kThreads) acc_s[i] = 0.f;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t stages_a = smem_u32(stages);
  const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty), acc_full_a = smem_u32(acc_full),
                 acc_empty_a = smem_u32(acc_empty);

  const int Hp = H >> 1, Wp = W >> 1;
  const int regions_x = (Wp + kWinW - 1) / kWinW, regions_y = (Hp + kWinH - 1) / kWinH;
  const long long total = static_cast<long long>(n_img) * regions_x * regions_y;
  const int my_regions = blockIdx.x < total ? static_cast<int>((total - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  auto decode = [&](long long region, int& img, int& ry, int& rx) {
    rx = static_cast<int>(region % regions_x);
    ry = static_cast<int>((region / regions_x) % regions_y);
    img = static_cast<int>(region / (static_cast<long long>(regions_x) * regions_y));
  };

  if (warp >= 5) {
    // ------------------------------------------------------------------ A builders: masked gradient tiles
    const int tp = threadIdx.x - 160;
    const int cc = tp % CH, ws = tp / CH;
    const uint32_t one_b = 0x3f80u, slope_b = bf16_bits(slope);
    // chunk j of this thread: window ws + j * WSTEP; its byte offset inside an A tile never changes
    uint32_t a_off[AIT];
#pragma unroll
    for (int j = 0; j < AIT; ++j) {
      const int w = ws + j * WSTEP, row = w & 7;
      a_off[j] = static_cast<uint32_t>((cc >> 3) * 16384 + (w >> 3) * 1024 + row * 128 + (((cc & 7) ^ row) << 4));
    }
    uint4 raw[AIT];
    uint32_t cw[AIT];
    auto load_region = [&](long long region) {
      int img, ry, rx;
      decode(region, img, ry, rx);
      const int py0 = ry * kWinH + (ws >> 3), px = rx * kWinW + (ws & 7);     // WSTEP is a multiple of 8: px is shared
      const size_t pix0 = (static_cast<size_t>(img) * Hp + py0) * Wp + px;
      const uint4* dp_j = reinterpret_cast<const uint4*>(dp + pix0 * COUT + cc * 8);
      const uint32_t* cd_j = codes + pix0 * CH + cc;
      const size_t row_step = static_cast<size_t>(WSTEP / 8) * Wp;                // pooled pixels between this thread's windows
#pragma unroll
      for (int j = 0; j < AIT; ++j) {
        const bool ok = py0 + j * (WSTEP / 8) < Hp && px < Wp;
        raw[j] = ok ? __ldcs(dp_j + j * row_step * (COUT / 8)) : make_uint4(0u, 0u, 0u, 0u);
        cw[j] = ok ? __ldcs(cd_j + j * row_step * CH) : 0u;
      }
    };
    if (my_regions > 0) load_region(blockIdx.x);
    int slot = 0;
    uint32_t phase = 1;                     // parity the empty barrier of `slot` must have completed
    for (int k = 0; k < my_regions; ++k) {
      // g = dP * (slope where the pre-activation was negative), in bf16 like dP itself; positions as PRMT selectors
      uint32_t g[AIT][4], sel_lo[AIT], sel_hi[AIT];
#pragma unroll
      for (int j = 0; j < AIT; ++j) {
        const uint32_t rw[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          // elements 2i / 2i+1 of the chunk: sign bits at 8i + 2 / 8i + 6 of the code word
          const uint32_t f = (((cw[j] >> (8 * i + 2)) & 1u) ? slope_b : one_b) |
                             ((((cw[j] >> (8 * i + 6)) & 1u) ? slope_b : one_b) << 16);
          const __nv_bfloat162 v = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&rw[i]),
                                           *reinterpret_cast<const __nv_bfloat162*>(&f));
          g[j][i] = *reinterpret_cast<const uint32_t*>(&v);
        }
        sel_lo[j] = cw[j] & 0x3333u;
        sel_hi[j] = (cw[j] >> 16) & 0x3333u;
      }
      if (k + 1 < my_regions) load_region(blockIdx.x + static_cast<long long>(k + 1) * gridDim.x);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        wait(empty_a + slot * 8, phase);
        const uint32_t a_tile = stages_a + slot * kStageBytes;
        const uint32_t lut = 0xffu << (8 * q);
#pragma unroll
        for (int j = 0; j < AIT; ++j) {
          const uint32_t m_lo = __byte_perm(lut, 0u, sel_lo[j]);      // byte e = 0xff where element e sits at position q
          const uint32_t m_hi = __byte_perm(lut, 0u, sel_hi[j]);
          sts128(a_tile + a_off[j], g[j][0] & __byte_perm(m_lo, 0u, 0x1100), g[j][1] & __byte_perm(m_lo, 0u, 0x3322),
                 g[j][2] & __byte_perm(m_hi, 0u, 0x1100), g[j][3] & __byte_perm(m_hi, 0u, 0x3322));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) arrive(full_a + slot * 8);
        if (++slot == kStages) { slot = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = mn_idesc(128, 64);
    uint64_t adesc[kStages], bdesc[kStages];
#pragma unroll
    for (int i = 0; i < kStages; ++i) {
      adesc[i] = mn_desc(stages_a + i * kStageBytes, 16384, 1024);
      bdesc[i] = mn_desc(stages_a + i * kStageBytes + kABytes, 16384, 1024);
    }
    int slot = 0, periods = 0, in_period = 0;
    uint32_t phase = 0;
    for (int k = 0; k < my_regions; ++k) {
      const int buf = periods & 1;
      if (in_period == 0) {
        wait(acc_empty_a + buf * 8, ((periods >> 1) & 1) ^ 1);
        tc_fence_after();
      }
      const bool last_of_period = in_period == kDrainEvery - 1 || k + 1 == my_regions;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        wait(full_a + slot * 8, phase);
        tc_fence_after();
        const uint64_t ad = slot == 0 ? adesc[0] : slot == 1 ? adesc[1] : adesc[2];
        const uint64_t bd = slot == 0 ? bdesc[0] : slot == 1 ? bdesc[1] : bdesc[2];
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_bf16(tmem_base + buf * 64, ad + (ks * 2048) / 16, bd + (ks * 2048) / 16, idesc,
                      (in_period == 0 && q == 0 && ks == 0) ? 0u : 1u);
          umma_commit(&empty[slot]);
          if (q == 3 && last_of_period) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++slot == kStages) { slot = 0; phase ^= 1u; }
      }
      if (++in_period == kDrainEvery) { in_period = 0; ++periods; }
    }
  } else {
    // ------------------------------------------------------------------ B builders (thread = window) + accumulator drain
    const int w = threadIdx.x, wy = w >> 3, wx = w & 7;
    constexpr int PRE = (kInRows * kInCols + 127) / 128;
    int sr[PRE], sq[PRE];
#pragma unroll
    for (int j = 0; j < PRE; ++j) {
      const int i = w + 128 * j;
      sr[j] = i < kInRows * kInCols ? i / kInCols : -1000000;
      sq[j] = i - (i / kInCols) * kInCols;
    }
    float pf[CIN][PRE];
    auto fetch = [&](long long region) {
      int img, ry, rx;
      decode(region, img, ry, rx);
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        const float* plane = x + (static_cast<size_t>(img) * c_total + c_offset + c) * H * W;
#pragma unroll
        for (int j = 0; j < PRE; ++j) {
          const int gy = 2 * ry * kWinH - 1 + sr[j], gx = 2 * rx * kWinW - 1 + sq[j];
          pf[c][j] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(plane + static_cast<size_t>(gy) * W + gx) : 0.f;
        }
      }
    };
    auto drain = [&](int period) {       // TMEM accumulator of `period` -> fp32 totals owned by this thread's channel
      const int buf = period & 1;
      wait(acc_full_a + buf * 8, (period >> 1) & 1);
      tc_fence_after();
      float* mine = acc_s + w * kAccPitch;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + buf * 64 + h * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) mine[h * 32 + i] += __uint_as_float(v[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive(acc_empty_a + buf * 8);
    };
    if (my_regions > 0) fetch(blockIdx.x);
    int slot = 0, drained = 0, in_period = 0, periods_done = 0;
    uint32_t phase = 1;
    const uint32_t b_row_off = static_cast<uint32_t>(kABytes + (w >> 3) * 1024 + (w & 7) * 128);
    const uint32_t s_in_a = smem_u32(s_in);
    for (int k = 0; k < my_regions; ++k) {
      asm volatile("bar.sync 1, 128;" ::: "memory");                 // everybody has read the previous region's patches
#pragma unroll
      for (int c = 0; c < CIN; ++c)
#pragma unroll
        for (int j = 0; j < PRE; ++j)
          if (sr[j] >= 0)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_in_a + static_cast<uint32_t>(((c * kInRows + sr[j]) * kInPitch + sq[j]) * 4)),
                         "f"(pf[c][j]) : "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (k + 1 < my_regions) fetch(blockIdx.x + static_cast<long long>(k + 1) * gridDim.x);
      // this window's 4 x 4 patch per channel as three bf16 terms (bit patterns in the UPPER halves): hi = the top 8
      // significant bits (truncation), mid = the next 8, lo = the last 8 -- x = hi + mid + lo exactly, four ALU ops per value
      uint32_t ph[CIN][4][4], pm[CIN][4][4], pl[CIN][4][4];
#pragma unroll
      for (int c = 0; c < CIN; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          float vals[4];
          const uint32_t ra = s_in_a + static_cast<uint32_t>(((c * kInRows + 2 * wy + r) * kInPitch + 2 * wx) * 4);
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(vals[0]), "=f"(vals[1]) : "r"(ra));
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(vals[2]), "=f"(vals[3]) : "r"(ra + 8));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t hb = __float_as_uint(vals[i]) & 0xffff0000u;
            const float r1 = vals[i] - __uint_as_float(hb);
            const uint32_t mb = __float_as_uint(r1) & 0xffff0000u;
            ph[c][r][i] = hb; pm[c][r][i] = mb; pl[c][r][i] = __float_as_uint(r1 - __uint_as_float(mb));
          }
        }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int dy = q >> 1, dx = q & 1;
        // row of window w for position q: three planes of T taps (plane stride TP, two taps per 32-bit word), then 1.0
        uint32_t words[BCHUNKS * 4];
#pragma unroll
        for (int i = 0; i < BCHUNKS * 4; ++i) words[i] = 0u;
#pragma unroll
        for (int pln = 0; pln < 3; ++pln)
#pragma unroll
          for (int t = 0; t < T; t += 2) {
            auto tap = [&](int tt) -> uint32_t {
              const int c = tt / 9, ky = (tt % 9) / 3, kx = tt % 3;
              return pln == 0 ? ph[c][dy + ky][dx + kx] : pln == 1 ? pm[c][dy + ky][dx + kx] : pl[c][dy + ky][dx + kx];
            };
            words[(pln * TP + t) >> 1] = t + 1 < T ? __byte_perm(tap(t), tap(t + 1), 0x7632) : (tap(t) >> 16);
          }
        words[ONES >> 1] = 0x3f80u;                                      // bf16 1.0 at column ONES (even)
        wait(empty_a + slot * 8, phase);
        const uint32_t b_row = stages_a + slot * kStageBytes + b_row_off;
#pragma unroll
        for (int ch = 0; ch < BCHUNKS; ++ch)
          sts128(b_row + ((ch ^ (w & 7)) << 4), words[4 * ch], words[4 * ch + 1], words[4 * ch + 2], words[4 * ch + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) arrive(full_a + slot * 8);
        if (++slot == kStages) { slot = 0; phase ^= 1u; }
      }
      // one period behind the MMA warp: the accumulator drained here was committed four regions ago
      if (++in_period == kDrainEvery) {
        in_period = 0;
        if (++periods_done >= 2) drain(drained++);
      }
    }
    const int periods_total = (my_regions + kDrainEvery - 1) / kDrainEvery;
    while (drained < periods_total) drain(drained++);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    // hi + mid + lo planes -> t1, the ones column -> s1; this CTA's row [COUT * T | COUT] of partial sums
    if (w < COUT) {
      float* row = part + static_cast<size_t>(blockIdx.x) * (COUT * T + COUT);
      const float* mine = acc_s + w * kAccPitch;
#pragma unroll
      for (int t = 0; t < T; ++t) row[w * T + t] = (mine[t] + mine[TP + t]) + mine[2 * TP + t];
      row[COUT * T + w] = mine[ONES];
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 128);
  }
}

// dW[c,t] = scale_c [ T1 - m1 S_t - m2 invstd ((G w_c)[t] - mu S_t) ],  m1 = s1 / count,  m2 = s2 / count with
// s2[c] = sum dA * xhat = invstd_c (sum_t w[c][t] T1[c][t] - mu_c s1[c]);  sums = [s1 (in) | s2 (out)] = [d beta | d gamma]
__global__ void first_wgrad_finalize_kernel(const float* __restrict__ t1, const double* __restrict__ gram,
                                            const float* __restrict__ w, const float* __restrict__ scale,
                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                            float* __restrict__ sums, double count, int cout, int T,
                                            float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * T) return;
  const int c = i / T, t = i % T;
  const double* S = gram;
  const double* G = gram + T;
  double gw = 0.0, wt1 = 0.0;
  for (int b = 0; b < T; ++b) {
    gw += G[t * T + b] * static_cast<double>(w[c * T + b]);
    wt1 += static_cast<double>(w[c * T + b]) * static_cast<double>(t1[c * T + b]);
  }
  const double s1 = static_cast<double>(sums[c]);
  const double s2 = static_cast<double>(invstd[c]) * (wt1 - static_cast<double>(mean[c]) * s1);
  if (t == 0) sums[cout + c] = static_cast<float>(s2);
  const double m1 = s1 / count, m2 = s2 / count;
  const double v = static_cast<double>(t1[i]) - m1 * S[t] -
                   m2 * static_cast<double>(invstd[c]) * (gw - static_cast<double>(mean[c]) * S[t]);
  dw[i] = static_cast<float>(static_cast<double>(scale[c]) * v);
}

}  // namespace

extern "C" {

size_t ctk_first_patch_gram_workspace_bytes(int cin) {
  const int T = 9 * cin;
  return cin > 0 ? static_cast<size_t>(ctk::num_sms()) * 4 * (T + T * (T + 1) / 2) * sizeof(double) : 0;   // <= 4 CTAs per SM
}

int ctk_first_patch_gram(const float* x, int n, int c_total, int c_offset, int cin, int H, int W, double* gram,
                         void* workspace, size_t workspace_bytes, void* stream) {
  CTK_REQUIRE(x && gram && n > 0 && H > 0 && W > 0 && c_offset >= 0 && c_offset + cin <= c_total);
  CTK_REQUIRE(cin == 1 || cin == 2);
  const int T = 9 * cin;
  const int items = T + T * (T + 1) / 2;
  cudaStream_t s = ctk::as_stream(stream);
  // Cin = 1 variants (CTK_GRAM_VARIANT): 0 = one role, one CTA per SM; 1 = two roles, two CTAs per SM; 2 = four roles, three
  static const int variant = [] { const char* e = getenv("CTK_GRAM_VARIANT"); return e ? atoi(e) : 1; }();
  const int per_sm = cin == 1 ? (variant == 2 ? 3 : variant == 1 ? 2 : 1) : 1;
  const int grid = ctk::num_sms() * per_sm;
  CTK_REQUIRE_WORKSPACE(workspace, workspace_bytes, static_cast<size_t>(grid) * items * sizeof(double));
  double* part = static_cast<double*>(workspace);
  if (cin == 1) {
    if (variant == 2) patch_gram_kernel<1, 4, 3><<<grid, 256, 0, s>>>(x, n, c_total, c_offset, H, W, part);
    else if (variant == 1) patch_gram_kernel<1, 2, 2><<<grid, 256, 0, s>>>(x, n, c_total, c_offset, H, W, part);
    else patch_gram_kernel<1, 1, 1><<<grid, 256, 0, s>>>(x, n, c_total, c_offset, H, W, part);
  } else {
    patch_gram_kernel<2, 4, 1><<<grid, 256, 0, s>>>(x, n, c_total, c_offset, H, W, part);
  }
  int st = ctk::check_launch();
  if (st != CTK_OK) return st;
  gram_reduce_kernel<<<(items + 7) / 8, 256, 0, s>>>(part, grid, T, gram);
  return ctk::check_launch();
}

int ctk_first_moments(const double* gram, const float* w, int cout, int cin, double count, float* moments,
                      void* stream) {
  CTK_REQUIRE(gram && w && moments && cout > 0 && cin > 0 && cin <= 2 && count >= 1.0);
  first_moments_kernel<<<(cout + 63) / 64, 64, 0, ctk::as_stream(stream)>>>(gram, w, cout, 9 * cin, count, moments);
  return ctk::check_launch();
}

size_t ctk_first_wgrad_codes_workspace_bytes(int cin, int cout) {
  return cin > 0 && cout > 0 ? static_cast<size_t>(ctk::num_sms()) * 2 * (9 * cin + 1) * cout * sizeof(float) : 0;
}

int ctk_first_wgrad_codes(const float* x, int n, int c_total, int c_offset, int cin, int H, int W, const void* codes_u32,
                          const void* dp_bf16, int cout, float slope, float* t1, float* sums, void* workspace,
                          size_t workspace_bytes, void* stream) {
  CTK_REQUIRE(x && codes_u32 && dp_bf16 && t1 && sums && n > 0 && H % 2 == 0 && W % 2 == 0);
  CTK_REQUIRE(c_offset >= 0 && c_offset + cin <= c_total);
  cudaStream_t s = ctk::as_stream(stream);
  // CTK_FIRST_WGRAD=cuda selects the CUDA-core gather (one shared-memory load and one FMA per window, channel and tap);
  // the default is the tcgen05 formulation (masked-gradient tiles x patch tiles over K = windows)
  static const bool use_tc = [] { const char* e = getenv("CTK_FIRST_WGRAD"); return !(e && e[0] == 'c'); }();
  const int grid = use_tc ? ctk::num_sms() : ctk::num_sms() * 2;
  const int T = 9 * cin;
  const int cols = (T + 1) * cout;
  CTK_REQUIRE_WORKSPACE(workspace, workspace_bytes, static_cast<size_t>(grid) * cols * sizeof(float));
  CTK_REQUIRE((reinterpret_cast<uintptr_t>(dp_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(codes_u32) & 3) == 0);
  float* part = static_cast<float*>(workspace);
  const __nv_bfloat16* dp = static_cast<const __nv_bfloat16*>(dp_bf16);
  const uint32_t* cd = static_cast<const uint32_t*>(codes_u32);
  if (cin == 1 && cout == 64) {
    if (use_tc) {
      auto kernel = first_wgrad_tc_kernel<1, 64>;
      CTK_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wtc::smem_bytes(1)));
      kernel<<<grid, wtc::threads(64), wtc::smem_bytes(1), s>>>(x, n, c_total, c_offset, H, W, cd, dp, slope, part);
    } else {
      first_wgrad_codes_kernel<1, 64><<<grid, 256, 0, s>>>(x, n, c_total, c_offset, H, W, cd, dp, slope, part);
    }
  } else if (cin == 2 && cout == 128) {
    if (use_tc) {
      auto kernel = first_wgrad_tc_kernel<2, 128>;
      CTK_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wtc::smem_bytes(2)));
      kernel<<<grid, wtc::threads(128), wtc::smem_bytes(2), s>>>(x, n, c_total, c_offset, H, W, cd, dp, slope, part);
    } else {
      first_wgrad_codes_kernel<2, 128><<<grid, 256, 0, s>>>(x, n, c_total, c_offset, H, W, cd, dp, slope, part);
    }
  } else {
    return CTK_ERR_UNSUPPORTED;
  }
  int st = ctk::check_launch();
  if (st != CTK_OK) return st;
  st = ctk::reduce_rows_f32(part, grid, cols, T * cout, t1, s);
  if (st != CTK_OK) return st;
  // sums = [d beta (this pass) | d gamma (ctk_first_wgrad_finalize)]
  CTK_CUDA_TRY(cudaMemsetAsync(sums + cout, 0, sizeof(float) * cout, s));
  return ctk::reduce_rows_f32(part + T * cout, grid, cols, cout, sums, s);
}

int ctk_first_wgrad_finalize(const float* t1, const double* gram, const float* w, const float* scale, const float* mean,
                             const float* invstd, float* sums, double count, int cout, int cin, float* dw,
                             void* stream) {
  CTK_REQUIRE(t1 && gram && w && scale && mean && invstd && sums && dw && cout > 0 && cin > 0 && count >= 1.0);
  const int total = cout * 9 * cin;
  first_wgrad_finalize_kernel<<<(total + 127) / 128, 128, 0, ctk::as_stream(stream)>>>(t1, gram, w, scale, mean, invstd,
                                                                                       sums, count, cout, 9 * cin, dw);
  return ctk::check_launch();
}

}  // extern "C"
