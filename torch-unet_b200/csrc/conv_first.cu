// First conv block: Conv2d(cin in {1,2}, cout, 3, 1, 1) + BN + LeakyReLU + MaxPool2d(2,2), fp32 NCHW planes in, NHWC
// bf16 out.  Replaces /root/reference/regression_model.py:14-17 (cin=2, cout=128) and two_branch_regression.py:10-13
// (cin=1, cout=64).
//
// K = 9*cin is tiny, so a direct convolution is bound by the fp32 pipe (576 FMA per pixel for 64 channels).
// Instead the block runs on the tensor cores as a GEMM whose A operand is built in shared memory by the
// threads themselves, with split-bf16 operands so the result is fp32-class despite bf16 multiplicands:
//   x = x_hi + x_lo,  w*bn_scale = w_hi + w_lo   (each half a bf16)
//   acc = sum_taps  x_hi*w_hi + x_lo*w_hi + x_hi*w_lo                      (dropped term x_lo*w_lo ~ 2^-18)
// Per input channel that is 14 32-bit K-words: 9 words (x_hi, x_lo)[tap] against (w_hi, w_hi)[tap], then
// 5 words (x_hi[2u], x_hi[2u+1]) against (w_lo[2u], w_lo[2u+1]).  K is padded to 32 (cin=1) or 64 (cin=2).
//
// Two kernels:
//  * conv_first_tc_kernel  -- raw, full-resolution conv output + batch statistics (one TMEM lane per PIXEL; the
//    "stored" training path and ctk_conv_first_raw)
//  * conv_first_ws_kernel  -- the pooled block as the models use it (one TMEM lane per 2x2 WINDOW, warp specialised,
//    TMA-store epilogue): eval, train (arg-max / sign codes) and fp32-class (hi, lo) outputs
#include "ctk_common.h"
#include "ctk_ptx.cuh"

#include <algorithm>

namespace {

using namespace ctk;

constexpr int kTileH = 16;
constexpr int kTileW = 8;
constexpr int kGroups = 4;
constexpr int kThreads = kGroups * 128;
constexpr int kTmemCols = 512;

template <int CIN, int COUT>
struct FirstCfg {
  static constexpr int kK = CIN == 1 ? 32 : 64;                 // padded K (bf16 elements)
  static constexpr int kKWords = kK / 2;
  static constexpr int kSub = kTmemCols / (kGroups * COUT);     // 16x8 sub-tiles per group iteration
  static constexpr int kRegionW = kTileW * kSub;
  static constexpr int kInW = kRegionW + 2;
  static constexpr int kInPitch = kInW + 1;
  static constexpr int kInH = kTileH + 2;
  static constexpr int kLbo = 128;                              // bytes between core matrices along K
  static constexpr int kSbo = (kK / 8) * 128;                   // bytes between 8-row groups along M / N
  static constexpr int kATileBytes = 16 * kSbo;                 // 128 rows
  static constexpr int kBBytes = (COUT / 8) * kSbo;
  static constexpr int kGroupBytes = (kSub * kATileBytes + CIN * kInH * kInPitch * 4 + 127) / 128 * 128;
  static constexpr int kSmemBytes = 1024 + kBBytes + kGroups * kGroupBytes + 256;
  static_assert(kSub >= 1, "too many output channels for the TMEM split");
};

__device__ __forceinline__ uint32_t split_hi_lo(float x) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
  return static_cast<uint32_t>(__bfloat16_as_ushort(hi)) | (static_cast<uint32_t>(__bfloat16_as_ushort(lo)) << 16);
}

// no-swizzle K-major operand descriptor (layout type 0): LBO = K-direction core-matrix stride, SBO = 8-row group stride
__device__ __forceinline__ uint64_t umma_smem_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// kTrain = false: folded BN shift + LeakyReLU + 2x2 max-pool, pooled NHWC output (eval mode).
// kTrain = true : raw conv output (no bias, no BN) stored at full resolution plus per-channel sum / sum of squares
//                 of the fp32 accumulators into stats[2*COUT] (train-mode BatchNorm needs batch statistics first).
template <int CIN, int COUT, bool kTrain>
__global__ void __launch_bounds__(kThreads, 1)
conv_first_tc_kernel(const float* __restrict__ x, int n_img, int c_total, int c_offset, int H, int W,
                     const float* __restrict__ w_folded, const float* __restrict__ shift, float slope,
                     __nv_bfloat16* __restrict__ out, int out_cstride, int out_coffset, int regions_x, int regions_y,
                     int total_regions, float* __restrict__ stats) {
  using C = FirstCfg<CIN, COUT>;
  // statistics: one private row per warp (plain read-modify-write by the owning lane), combined in warp order at the end
  // and stored as this CTA's row of partial sums -- no shared or global floating-point atomics (deterministic)
  constexpr int kWarps = kThreads / 32;
  __shared__ float s_stat[kTrain ? kWarps * 2 * COUT : 1];
  if constexpr (kTrain)
    for (int i = threadIdx.x; i < kWarps * 2 * COUT; i += kThreads) s_stat[i] = 0.f;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_smem = smem;
  uint8_t* groups_smem = smem + C::kBBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(groups_smem + kGroups * C::kGroupBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kGroups);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = warp >> 2;
  const int gt = threadIdx.x & 127;          // thread within group = TMEM lane = pixel of the 16x8 sub-tile
  const int ew = warp & 3;

  // ---- one-time setup: barriers, TMEM, the B operand (folded weights split into hi/lo bf16)
  if (threadIdx.x == 0) {
    for (int g = 0; g < kGroups; ++g) mbar_init(&bars[g], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<1>(tmem_slot, kTmemCols);
  for (int i = threadIdx.x; i < COUT * C::kKWords; i += kThreads) {
    const int n = i / C::kKWords, kw = i % C::kKWords;
    uint32_t word = 0;
    const int ch = kw / 14, j = kw % 14;
    if (kw == 14 * CIN) {
      word = shift ? split_hi_lo(__ldg(shift + n)) : 0u;                  // (shift_hi, shift_lo) against A's (1, 1)
    } else if (ch < CIN) {
      const float* wr = w_folded + (n * CIN + ch) * 9;
      if (j < 9) {
        const uint32_t hl = split_hi_lo(__ldg(wr + j));
        word = (hl & 0xffffu) | (hl << 16);                               // (w_hi, w_hi)
      } else {
        const int t0 = 2 * (j - 9);
        const uint32_t a = split_hi_lo(__ldg(wr + t0)) >> 16;             // w_lo[t0]
        const uint32_t b = t0 + 1 < 9 ? (split_hi_lo(__ldg(wr + t0 + 1)) >> 16) : 0u;
        word = a | (b << 16);
      }
    }
    const int k = 2 * kw;                                                 // first bf16 element of this word
    const uint32_t off = (n >> 3) * C::kSbo + (k >> 3) * C::kLbo + (n & 7) * 16 + (k & 7) * 2;
    *reinterpret_cast<uint32_t*>(b_smem + off) = word;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  uint8_t* a_smem = groups_smem + group * C::kGroupBytes;
  uint32_t* in_smem = reinterpret_cast<uint32_t*>(a_smem + C::kSub * C::kATileBytes);
  const uint32_t tmem_group = tmem_base + static_cast<uint32_t>(group * C::kSub * COUT);
  const int r = gt >> 3, cpx = gt & 7;
  constexpr uint32_t idesc = umma_idesc_bf16_f32(128, COUT);
  uint32_t parity = 0;

  // input prefetch: the halo of the NEXT region is loaded into registers while this one is being processed.
  // Element i = gt + 128*j of the halo is the same (row, column) every iteration, so the index math is hoisted.
  constexpr int kPref = (C::kInH * C::kInW + 127) / 128;
  int pre_rr[kPref], pre_q[kPref];
#pragma unroll
  for (int j = 0; j < kPref; ++j) {
    const int i = gt + 128 * j;
    pre_rr[j] = i < C::kInH * C::kInW ? i / C::kInW : -1000000;     // out-of-range slots never pass the bounds test
    pre_q[j] = i - (i / C::kInW) * C::kInW;
  }
  float pref[CIN][kPref];
  auto load_region = [&](int region) {
    const int rx = region % regions_x;
    const int ry = (region / regions_x) % regions_y;
    const int img = region / (regions_x * regions_y);
    const int y0 = ry * kTileH - 1, x0 = rx * C::kRegionW - 1;
    const float* plane0 = x + (static_cast<size_t>(img) * c_total + c_offset) * H * W;
#pragma unroll
    for (int j = 0; j < kPref; ++j) {
      const int gy = y0 + pre_rr[j], gx = x0 + pre_q[j];
      const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      const size_t off = static_cast<size_t>(gy) * W + gx;
#pragma unroll
      for (int c = 0; c < CIN; ++c) pref[c][j] = ok ? __ldg(plane0 + static_cast<size_t>(c) * H * W + off) : 0.f;
    }
  };
  const int region_first = blockIdx.x * kGroups + group;
  const int region_step = gridDim.x * kGroups;
  if (region_first < total_regions) load_region(region_first);

  for (int region = region_first; region < total_regions; region += region_step) {
    const int rx = region % regions_x;
    const int ry = (region / regions_x) % regions_y;
    const int img = region / (regions_x * regions_y);
    const int y0 = ry * kTileH, x0 = rx * C::kRegionW;

    // ---- stage the prefetched input halo as packed (hi | lo << 16) words (zero padded), then prefetch the next one
#pragma unroll
    for (int c = 0; c < CIN; ++c)
#pragma unroll
      for (int j = 0; j < kPref; ++j)
        if (pre_rr[j] >= 0) in_smem[(c * C::kInH + pre_rr[j]) * C::kInPitch + pre_q[j]] = split_hi_lo(pref[c][j]);
    asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
    if (region + region_step < total_regions) load_region(region + region_step);

    // ---- every thread writes the K-words of its pixel for each sub-tile
#pragma unroll
    for (int s = 0; s < C::kSub; ++s) {
      uint32_t kw[C::kKWords];
#pragma unroll
      for (int i = 0; i < C::kKWords; ++i) kw[i] = 0;
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        uint32_t t[9];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            t[ky * 3 + kx] = in_smem[(c * C::kInH + r + ky) * C::kInPitch + s * kTileW + cpx + kx];
#pragma unroll
        for (int j = 0; j < 9; ++j) kw[c * 14 + j] = t[j];
#pragma unroll
        for (int u = 0; u < 4; ++u) kw[c * 14 + 9 + u] = __byte_perm(t[2 * u], t[2 * u + 1], 0x5410);   // (hi, hi)
        kw[c * 14 + 13] = t[8] & 0xffffu;
      }
      kw[14 * CIN] = 0x3f803f80u;                                         // bf16 (1, 1): adds the folded BN shift inside the MMA
      uint8_t* tile = a_smem + s * C::kATileBytes + (gt >> 3) * C::kSbo + (gt & 7) * 16;
#pragma unroll
      for (int j = 0; j < C::kK / 8; ++j)
        *reinterpret_cast<uint4*>(tile + j * C::kLbo) = make_uint4(kw[4 * j], kw[4 * j + 1], kw[4 * j + 2], kw[4 * j + 3]);
    }
    fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
    tc_fence_before();
    asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");

    // ---- one thread issues the MMAs of this region and commits to the group's barrier
    if (ew == 0) {                      // warp-uniform branch; the single issuing lane is elected inside
      tc_fence_after();
      const uint64_t desc0 = umma_smem_desc_nosw(0, C::kLbo, C::kSbo);
      const uint64_t bdesc = desc0 | static_cast<uint64_t>(smem_u32(b_smem) >> 4);
      const uint64_t adesc = desc0 | static_cast<uint64_t>(smem_u32(a_smem) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < C::kSub; ++s) {
#pragma unroll
          for (int ks = 0; ks < C::kK / 16; ++ks) {
            umma_bf16(tmem_group + s * COUT, adesc + ((s * C::kATileBytes + ks * 2 * C::kLbo) >> 4),
                      bdesc + ((ks * 2 * C::kLbo) >> 4), idesc, ks != 0 ? 1u : 0u);
          }
        }
        umma_commit(&bars[group]);
      }
      __syncwarp();
    }
    mbar_wait(&bars[group], parity);
    parity ^= 1;
    tc_fence_after();

    // ---- epilogue per sub-tile: raw bf16 store + per-channel statistics
    const int y = y0 + r;
#pragma unroll 1
    for (int s = 0; s < C::kSub; ++s) {
      const int xg = x0 + s * kTileW + cpx;
      const bool valid = y < H && xg < W;
#pragma unroll 1
      for (int cb = 0; cb < COUT / 32; ++cb) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_group + (static_cast<uint32_t>(ew * 32) << 16) + s * COUT + cb * 32, v);
        tmem_ld_wait();
        if constexpr (kTrain) {
          float f[32], sq[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            f[i] = valid ? __uint_as_float(v[i]) : 0.f;
            sq[i] = f[i] * f[i];
          }
          const float tot = warp_transpose_sum32(f, lane);
          const float tot2 = warp_transpose_sum32(sq, lane);
          s_stat[warp * 2 * COUT + cb * 32 + lane] += tot;
          s_stat[warp * 2 * COUT + COUT + cb * 32 + lane] += tot2;
          if (valid) {
            __nv_bfloat16* dst = out + (static_cast<size_t>(img) * H * W + static_cast<size_t>(y) * W + xg) * out_cstride +
                                 out_coffset + cb * 32;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<uint4*>(dst)[i] =
                  make_uint4(pack_bf16x2(__uint_as_float(v[8 * i]), __uint_as_float(v[8 * i + 1])),
                             pack_bf16x2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3])),
                             pack_bf16x2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5])),
                             pack_bf16x2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7])));
          }
          continue;
        }
      }
    }
    tc_fence_before();   // TMEM reads of this iteration are ordered before the barrier the next MMA issue follows
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if constexpr (kTrain)
    for (int i = threadIdx.x; i < 2 * COUT; i += kThreads) {
      float tot = 0.f;
      for (int w = 0; w < kWarps; ++w) tot += s_stat[w * 2 * COUT + i];
      stats[static_cast<size_t>(blockIdx.x) * 2 * COUT + i] = tot;
    }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, kTmemCols);
  }
}

// ======================================================================================================================
// Pooled variant (the block as the models use it): one TMEM lane per 2x2 WINDOW, warp specialised.
//
// The four pixels of a pooling window are four separate A tiles (same 128 windows, window position q = (dy, dx)); the MMA
// of position q accumulates into its own 64 TMEM columns.  The thread that owns a window then reads its four positions
// from its own lane and takes the max in registers: no cross-lane shuffles, no selects, the bf16 pack and the LeakyReLU
// run on the pooled quarter of the data, and the window's patch (4 x 4 inputs per channel) is read from shared memory
// once for all four A rows (~430 instructions per window and 64 channels against ~1600 for a pixel-per-lane epilogue).
// kMode 1 also stores, per pooled element, the arg-max position and the sign of the pre-activation (4-bit code, eight
// channels per 32-bit word) -- what the training backward needs to route gradients without recomputing the conv; the
// position rides in the two low mantissa bits of the fp32 accumulators through the max (<= 3 fp32 ulp).
// kMode 2 stores fp32-class results as bf16 (hi, lo) pairs.
// Three concurrent roles, so that each one's latencies hide behind the others':
//   warps 8-15  two producer groups (128 threads = 128 windows each) take alternate regions: input halo -> (hi|lo) words
//               in shared memory -> the four A tiles of the region into a ring slot            [a_empty -> a_full]
//   warp 16     one thread issues, per 64-channel pass, 4 positions x K/16 MMAs (M = 128 windows, N = 64) into one of
//               two 256-column TMEM buffers, commits acc_full, and a_empty after the region's last pass
//   warps 0-7   two epilogue groups (one per TMEM buffer, warp % 4 = lane quadrant): max over the four positions in
//               registers -> LeakyReLU -> bf16 -> 32-byte stores (+ codes / (hi, lo) pairs)     [acc_full -> acc_empty]
template <int CIN>
struct WsCfg {
  static constexpr int kK = CIN == 1 ? 32 : 64;
  static constexpr int kKWords = kK / 2;
  static constexpr int kWinH = 16, kWinW = 8;
  static constexpr int kInH = 2 * kWinH + 2, kInW = 2 * kWinW + 2;
  static constexpr int kInPitch = 24;
  static constexpr int kLbo = 128;
  static constexpr int kSbo = (kK / 8) * 128;
  static constexpr int kATileBytes = 16 * kSbo;
  static constexpr int kSlotBytes = 4 * kATileBytes;             // 32 KB (CIN 1) / 64 KB (CIN 2)
  static constexpr int kSlots = CIN == 1 ? 4 : 2;
  static constexpr int kInBytes = CIN * kInH * kInPitch * 4;
  static constexpr int kPref = (kInH * kInW + 127) / 128;
  static constexpr int kThreads = 17 * 32;
  static constexpr int kStageBytes = 8 * 2 * 4096;              // output staging: 8 epilogue warps x 2 x (32 windows x 128 B)
  static constexpr int smem_bytes(int cout) {
    return 1024 + ((cout / 8) * kSbo + 1023) / 1024 * 1024 + kSlots * kSlotBytes + kStageBytes + 2 * kInBytes + 256;
  }
};

template <int CIN, int kMode>
__global__ void __launch_bounds__(WsCfg<CIN>::kThreads, 1)
conv_first_ws_kernel(const __grid_constant__ CUtensorMap tm_out, const float* __restrict__ x, int n_img, int c_total,
                     int c_offset, int H, int W, const float* __restrict__ w_folded, const float* __restrict__ shift,
                     float slope, int cout,
                     __nv_bfloat16* __restrict__ out, int out_cstride, int out_coffset, uint32_t* __restrict__ codes,
                     __nv_bfloat16* __restrict__ out_lo, FastDiv div_rx, FastDiv div_ry, int total_regions) {
  using C = WsCfg<CIN>;
  constexpr bool kCodes = kMode == 1;
  constexpr int kPosCols = 64;                 // TMEM columns per window position (64 channels)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_smem = smem;
  uint8_t* slots = smem + ((cout / 8) * C::kSbo + 1023) / 1024 * 1024;
  uint8_t* stage_base = slots + C::kSlots * C::kSlotBytes;       // 1024-byte aligned: SWIZZLE_128B tiles for the TMA stores
  uint8_t* in_base = stage_base + C::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(in_base + 2 * C::kInBytes);
  uint64_t* a_full = bars;                       // [kSlots], 4 producer-warp arrivals
  uint64_t* a_empty = bars + C::kSlots;          // [kSlots], one MMA commit
  uint64_t* acc_full = bars + 2 * C::kSlots;     // [2], one MMA commit
  uint64_t* acc_empty = acc_full + 2;            // [2], 4 epilogue-warp arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int passes = cout >> 6;                  // 64 output channels per TMEM buffer

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::kSlots; ++i) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    fence_mbar_init();
    tma_prefetch_desc(&tm_out);
  }
  if (warp == 16) tmem_alloc<1>(tmem_slot, 512);
  // B operand (folded weights split into hi/lo bf16), built once per CTA by everybody
  for (int i = threadIdx.x; i < cout * C::kKWords; i += C::kThreads) {
    const int n = i / C::kKWords, kw = i % C::kKWords;
    uint32_t word = 0;
    const int ch = kw / 14, j = kw % 14;
    if (kw == 14 * CIN) {
      word = split_hi_lo(__ldg(shift + n));                          // (shift_hi, shift_lo) against A's (1, 1)
    } else if (ch < CIN) {
      const float* wr = w_folded + (n * CIN + ch) * 9;
      if (j < 9) {
        const uint32_t hl = split_hi_lo(__ldg(wr + j));
        word = (hl & 0xffffu) | (hl << 16);                               // (w_hi, w_hi)
      } else {
        const int t0 = 2 * (j - 9);
        const uint32_t a = split_hi_lo(__ldg(wr + t0)) >> 16;        // w_lo[t0]
        const uint32_t b = t0 + 1 < 9 ? (split_hi_lo(__ldg(wr + t0 + 1)) >> 16) : 0u;
        word = a | (b << 16);
      }
    }
    const int k = 2 * kw;
    const uint32_t off = (n >> 3) * C::kSbo + (k >> 3) * C::kLbo + (n & 7) * 16 + (k & 7) * 2;
    *reinterpret_cast<uint32_t*>(b_smem + off) = word;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int Hp = H >> 1, Wp = W >> 1;

  auto decode = [&](int region, int& img, int& ry, int& rx) {
    const uint32_t q1 = fdiv(static_cast<uint32_t>(region), div_rx);
    rx = region - static_cast<int>(q1 * div_rx.d);
    const uint32_t q2 = fdiv(q1, div_ry);
    ry = static_cast<int>(q1 - q2 * div_ry.d);
    img = static_cast<int>(q2);
  };

  if (warp >= 8 && warp < 16) {
    // ------------------------------------------------------------------ producers
    const int pg = (warp - 8) >> 2;
    const int gt = threadIdx.x & 127;            // window of the region this thread builds
    const int wy = gt >> 3, wx = gt & 7;
    uint32_t* in_smem = reinterpret_cast<uint32_t*>(in_base + pg * C::kInBytes);
    int pre_rr[C::kPref], pre_q[C::kPref];
#pragma unroll
    for (int j = 0; j < C::kPref; ++j) {
      const int i = gt + 128 * j;
      pre_rr[j] = i < C::kInH * C::kInW ? i / C::kInW : -1000000;
      pre_q[j] = i - (i / C::kInW) * C::kInW;
    }
    float pref[CIN][C::kPref];
    auto load_region = [&](int region) {
      int img, ry, rx;
      decode(region, img, ry, rx);
      const int y0 = ry * 2 * C::kWinH - 1, x0 = rx * 2 * C::kWinW - 1;
      const float* plane0 = x + (static_cast<size_t>(img) * c_total + c_offset) * H * W;
#pragma unroll
      for (int j = 0; j < C::kPref; ++j) {
        const int gy = y0 + pre_rr[j], gx = x0 + pre_q[j];
        const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
        const size_t off = static_cast<size_t>(gy) * W + gx;
#pragma unroll
        for (int c = 0; c < CIN; ++c) pref[c][j] = ok ? __ldg(plane0 + static_cast<size_t>(c) * H * W + off) : 0.f;
      }
    };
    const int k_step = 2;
    int k = pg;
    int region = blockIdx.x + k * gridDim.x;
    if (region < total_regions) load_region(region);
    for (; region < total_regions; k += k_step, region = blockIdx.x + k * gridDim.x) {
      // halo -> packed (hi | lo << 16) words; the next region of this group is prefetched into registers meanwhile
#pragma unroll
      for (int c = 0; c < CIN; ++c)
#pragma unroll
        for (int j = 0; j < C::kPref; ++j)
          if (pre_rr[j] >= 0) in_smem[(c * C::kInH + pre_rr[j]) * C::kInPitch + pre_q[j]] = split_hi_lo(pref[c][j]);
      asm volatile("bar.sync %0, 128;" ::"r"(pg + 1) : "memory");
      const int next = blockIdx.x + (k + k_step) * gridDim.x;
      if (next < total_regions) load_region(next);
      uint32_t pt[CIN][4][4];
#pragma unroll
      for (int c = 0; c < CIN; ++c)
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const uint2* row = reinterpret_cast<const uint2*>(in_smem + (c * C::kInH + 2 * wy + rr) * C::kInPitch + 2 * wx);
          const uint2 a = row[0], b = row[1];
          pt[c][rr][0] = a.x; pt[c][rr][1] = a.y; pt[c][rr][2] = b.x; pt[c][rr][3] = b.y;
        }
      asm volatile("bar.sync %0, 128;" ::"r"(pg + 1) : "memory");      // everybody has its patch: in_smem may be restaged
      const int slot = k % C::kSlots;
      mbar_wait(&a_empty[slot], ((k / C::kSlots) & 1) ^ 1);
      uint8_t* a_smem = slots + slot * C::kSlotBytes;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int dy = q >> 1, dx = q & 1;
        uint32_t kw[C::kKWords];
#pragma unroll
        for (int i = 0; i < C::kKWords; ++i) kw[i] = 0;
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          uint32_t t[9];
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) t[ky * 3 + kx] = pt[c][dy + ky][dx + kx];
#pragma unroll
          for (int j = 0; j < 9; ++j) kw[c * 14 + j] = t[j];
#pragma unroll
          for (int u = 0; u < 4; ++u) kw[c * 14 + 9 + u] = __byte_perm(t[2 * u], t[2 * u + 1], 0x5410);   // (hi, hi)
          kw[c * 14 + 13] = t[8] & 0xffffu;
        }
        kw[14 * CIN] = 0x3f803f80u;                                       // bf16 (1, 1): adds the folded BN shift inside the MMA
        uint8_t* tile = a_smem + q * C::kATileBytes + (gt >> 3) * C::kSbo + (gt & 7) * 16;
#pragma unroll
        for (int j = 0; j < C::kK / 8; ++j)
          *reinterpret_cast<uint4*>(tile + j * C::kLbo) = make_uint4(kw[4 * j], kw[4 * j + 1], kw[4 * j + 2], kw[4 * j + 3]);
      }
      fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[slot]);
    }
  } else if (warp == 16) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, 64);
    const uint64_t desc0 = umma_smem_desc_nosw(0, C::kLbo, C::kSbo);
    int u = 0;
    for (int k = 0, region = blockIdx.x; region < total_regions; ++k, region += gridDim.x) {
      const int slot = k % C::kSlots;
      mbar_wait(&a_full[slot], (k / C::kSlots) & 1);
      tc_fence_after();
      const uint64_t adesc = desc0 | static_cast<uint64_t>(smem_u32(slots + slot * C::kSlotBytes) >> 4);
      for (int p = 0; p < passes; ++p, ++u) {
        const int buf = u & 1;
        mbar_wait(&acc_empty[buf], ((u >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t bdesc = desc0 | static_cast<uint64_t>((smem_u32(b_smem) + p * 8 * C::kSbo) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int ks = 0; ks < C::kK / 16; ++ks)
              umma_bf16(tmem_base + buf * 256 + q * kPosCols, adesc + ((q * C::kATileBytes + ks * 2 * C::kLbo) >> 4),
                        bdesc + ((ks * 2 * C::kLbo) >> 4), idesc, ks != 0 ? 1u : 0u);
          umma_commit(&acc_full[buf]);
          if (p == passes - 1) umma_commit(&a_empty[slot]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: group eg drains TMEM buffer eg
    const int eg = warp >> 2, ew = warp & 3;
    const int gt = ew * 32 + lane;               // TMEM lane = window
    const int wy = gt >> 3, wx = gt & 7;
    const __nv_bfloat162 slope2 = __float2bfloat162_rn(slope);
    int u = 0;
    int stage_turn = 0;
    for (int k = 0, region = blockIdx.x; region < total_regions; ++k, region += gridDim.x) {
      int img, ry, rx;
      decode(region, img, ry, rx);
      const int py = ry * C::kWinH + wy, px = rx * C::kWinW + wx;
      const bool valid = py < Hp && px < Wp;
      const size_t pooled_pix = (static_cast<size_t>(img) * Hp + py) * Wp + px;
      for (int p = 0; p < passes; ++p, ++u) {
        if ((u & 1) != eg) continue;
        uint32_t code_words[8];
        uint32_t stage_addr = 0;
        if constexpr (kMode != 2) {
          stage_addr = smem_u32(stage_base) + static_cast<uint32_t>((warp * 2 + stage_turn) * 4096);
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that last read this tile
          __syncwarp();
        }
        mbar_wait(&acc_full[eg], (u >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + eg * 256;
        // Eight channels (x the four window positions) per step, TMEM loads one step ahead of the arithmetic: the loads of
        // step h + 1 are in flight while step h is reduced, packed and staged (ncu, round 2: the epilogue warps sat on the
        // TMEM scoreboard for most of a unit when every 16-channel load was waited for before anything else was issued).
        uint32_t va[4][8], vb[4][8];
        auto load8 = [&](uint32_t (&v)[4][8], int h8) {
#pragma unroll
          for (int q = 0; q < 4; ++q) tmem_ld_32x8(taddr + q * 64 + h8 * 8, v[q]);
        };
        auto release = [&]() {                    // all TMEM reads of this unit are done: hand the buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[eg]);
        };
        auto step8 = [&](const uint32_t (&v)[4][8], int h8) {
          float best[8];
          const int ch = p * 64 + h8 * 8;
          if constexpr (kCodes) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float m = __uint_as_float((v[0][i] & ~3u) | 3u);
              m = fmaxf(m, __uint_as_float((v[1][i] & ~3u) | 2u));
              m = fmaxf(m, __uint_as_float((v[2][i] & ~3u) | 1u));
              best[i] = fmaxf(m, __uint_as_float(v[3][i] & ~3u));
            }
            uint32_t cw = 0u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t b = __float_as_uint(best[i]);
              const uint32_t nib = (3u - (b & 3u)) | ((b >> 29) & 4u);      // arg-max position | sign << 2
              cw |= nib << (4 * i);
            }
            code_words[h8] = cw;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              best[i] = fmaxf(fmaxf(__uint_as_float(v[0][i]), __uint_as_float(v[1][i])),
                              fmaxf(__uint_as_float(v[2][i]), __uint_as_float(v[3][i])));
          }
          if constexpr (kMode == 2) {
            uint32_t oh[4], ol[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t s0 = split_hi_lo(leaky(best[2 * i], slope)), s1 = split_hi_lo(leaky(best[2 * i + 1], slope));
              oh[i] = (s0 & 0xffffu) | (s1 << 16);
              ol[i] = (s0 >> 16) | (s1 & 0xffff0000u);
            }
            if (valid) {
              const size_t off = pooled_pix * out_cstride + out_coffset + ch;
              *reinterpret_cast<uint4*>(out + off) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
              *reinterpret_cast<uint4*>(out_lo + off) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
            }
          } else {
            // pooled bf16 output: staged as a 128B-swizzled [32 windows][64 channels] tile, stored by one TMA instruction
            // per warp and pass -- registers are free again after the shared-memory store, every global write is a full
            // 128-byte line, and ragged edges are clipped by the tensor map
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = leaky_bf16x2(pack_bf16x2(best[2 * i], best[2 * i + 1]), slope2);
            const uint32_t row = stage_addr + static_cast<uint32_t>(lane * 128);
            const uint32_t c0 = static_cast<uint32_t>((h8 ^ (lane & 7)) * 16);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + c0), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
          }
        };
        load8(va, 0);
#pragma unroll
        for (int h8 = 0; h8 < 8; h8 += 2) {
          tmem_ld_wait();
          load8(vb, h8 + 1);
          step8(va, h8);
          tmem_ld_wait();
          if (h8 + 2 < 8) load8(va, h8 + 2);
          else release();
          step8(vb, h8 + 1);
        }
        if constexpr (kCodes) {
          if (valid) {                                      // the window's 64 codes of this pass: one full 32-byte sector
            uint4* cdst = reinterpret_cast<uint4*>(codes + pooled_pix * (cout >> 3) + p * 8);
            cdst[0] = make_uint4(code_words[0], code_words[1], code_words[2], code_words[3]);
            cdst[1] = make_uint4(code_words[4], code_words[5], code_words[6], code_words[7]);
          }
        }
        if constexpr (kMode != 2) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            asm volatile(
                "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                    reinterpret_cast<uint64_t>(&tm_out)),
                "r"(stage_addr), "r"(out_coffset + p * 64), "r"(rx * C::kWinW), "r"(ry * C::kWinH + ew * 4), "r"(img)
                : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          stage_turn ^= 1;
        }
      }
    }
    if constexpr (kMode != 2) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores must land before the CTA exits
      __syncwarp();
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

template <int CIN, int kMode>
int launch_first_win(const float* x, int n, int c_total, int c_offset, int H, int W, const float* w_folded,
                     const float* shift, float slope, int cout, __nv_bfloat16* out, int out_cstride, int out_coffset,
                     uint32_t* codes, __nv_bfloat16* out_lo, cudaStream_t stream) {
  using WS = WsCfg<CIN>;
  const int regions_x = (W / 2 + WS::kWinW - 1) / WS::kWinW;
  const int regions_y = (H / 2 + WS::kWinH - 1) / WS::kWinH;
  const long long total = static_cast<long long>(n) * regions_x * regions_y;
  if (total >= (1ll << 30)) return CTK_ERR_BAD_ARG;
  auto ws = conv_first_ws_kernel<CIN, kMode>;
  const int ws_smem = WS::smem_bytes(cout);
  CTK_CUDA_TRY(cudaFuncSetAttribute(ws, cudaFuncAttributeMaxDynamicSharedMemorySize, ws_smem));
  const int ws_grid = static_cast<int>(std::min<long long>(total, ctk::num_sms()));
  CUtensorMap tm_out;
  {
    const uint64_t cs = static_cast<uint64_t>(out_cstride);
    const uint64_t dims[4] = {cs, static_cast<uint64_t>(W / 2), static_cast<uint64_t>(H / 2), static_cast<uint64_t>(n)};
    const uint64_t strides[3] = {cs * 2, static_cast<uint64_t>(W / 2) * cs * 2, static_cast<uint64_t>(H / 2) * (W / 2) * cs * 2};
    const uint32_t box[4] = {64, 8, 4, 1};
    int st = ctk::encode_tmap_bf16_sw128(&tm_out, out, 4, dims, strides, box);
    if (st != CTK_OK) return st;
  }
  ws<<<ws_grid, WS::kThreads, ws_smem, stream>>>(tm_out, x, n, c_total, c_offset, H, W, w_folded, shift, slope, cout, out,
                                                 out_cstride, out_coffset, codes, out_lo, make_fastdiv(regions_x),
                                                 make_fastdiv(regions_y), static_cast<int>(total));
  return ctk::check_launch();
}

template <int CIN, int COUT, bool kTrain>
int launch_first(const float* x, int n, int c_total, int c_offset, int H, int W, const float* w_folded,
                 const float* shift, float slope, __nv_bfloat16* out, int out_cstride, int out_coffset, float* stats,
                 cudaStream_t stream, void* workspace = nullptr, size_t workspace_bytes = 0) {
  using C = FirstCfg<CIN, COUT>;
  const int regions_x = (W + C::kRegionW - 1) / C::kRegionW;
  const int regions_y = (H + kTileH - 1) / kTileH;
  const long long total = static_cast<long long>(n) * regions_x * regions_y;
  if (total >= (1ll << 31)) return CTK_ERR_BAD_ARG;
  auto kernel = conv_first_tc_kernel<CIN, COUT, kTrain>;
  CTK_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
  const int grid = static_cast<int>(std::min<long long>((total + kGroups - 1) / kGroups, ctk::num_sms()));
  float* part = nullptr;
  if (kTrain) {
    CTK_REQUIRE_WORKSPACE(workspace, workspace_bytes, static_cast<size_t>(grid) * 2 * COUT * sizeof(float));
    part = static_cast<float*>(workspace);
  }
  kernel<<<grid, kThreads, C::kSmemBytes, stream>>>(x, n, c_total, c_offset, H, W, w_folded, shift, slope, out,
                                                    out_cstride, out_coffset, regions_x, regions_y,
                                                    static_cast<int>(total), kTrain ? part : stats);
  int st = ctk::check_launch();
  if (st != CTK_OK || !kTrain) return st;
  return ctk::reduce_rows_f32(part, grid, 2 * COUT, 2 * COUT, stats, stream);
}

}  // namespace

static int first_pool_dispatch(const float* x, int n, int c_total, int c_offset, int cin, int H, int W,
                               const float* w_folded, const float* shift, int cout, float slope, void* out_bf16,
                               int out_cstride, int out_coffset, void* codes, void* out_lo_bf16, void* stream) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(x && w_folded && shift && out_bf16 && n > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0);
  CTK_REQUIRE(c_offset >= 0 && c_offset + cin <= c_total && out_coffset >= 0 && out_coffset + cout <= out_cstride);
  CTK_REQUIRE(out_cstride % 8 == 0 && out_coffset % 8 == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0);
  CTK_REQUIRE(cout > 0 && cout % 64 == 0 && cout <= 256 && (reinterpret_cast<uintptr_t>(codes) & 15) == 0);
  CTK_REQUIRE((reinterpret_cast<uintptr_t>(out_lo_bf16) & 15) == 0 && !(codes && out_lo_bf16));
  cudaStream_t s = ctk::as_stream(stream);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(out_bf16);
  __nv_bfloat16* out_lo = static_cast<__nv_bfloat16*>(out_lo_bf16);
  uint32_t* cd = static_cast<uint32_t*>(codes);
  const int mode = cd ? 1 : (out_lo ? 2 : 0);
#define CTK_FIRST_LAUNCH(CIN, MODE)                                                                                  \
  return launch_first_win<CIN, MODE>(x, n, c_total, c_offset, H, W, w_folded, shift, slope, cout, out, out_cstride, \
                                     out_coffset, cd, out_lo, s)
  if (cin == 1) {
    if (mode == 0) { CTK_FIRST_LAUNCH(1, 0); }
    if (mode == 1) { CTK_FIRST_LAUNCH(1, 1); }
    CTK_FIRST_LAUNCH(1, 2);
  }
  if (cin == 2) {
    if (mode == 0) { CTK_FIRST_LAUNCH(2, 0); }
    if (mode == 1) { CTK_FIRST_LAUNCH(2, 1); }
    CTK_FIRST_LAUNCH(2, 2);
  }
#undef CTK_FIRST_LAUNCH
  return CTK_ERR_UNSUPPORTED;
}

extern "C" int ctk_conv_first_eval(const float* x, int n, int c_total, int c_offset, int cin, int H, int W,
                                   const float* w_folded, const float* shift, int cout, float slope, void* out_bf16,
                                   int out_cstride, int out_coffset, void* stream) {
  return first_pool_dispatch(x, n, c_total, c_offset, cin, H, W, w_folded, shift, cout, slope, out_bf16, out_cstride,
                             out_coffset, nullptr, nullptr, stream);
}

extern "C" int ctk_conv_first_eval_split(const float* x, int n, int c_total, int c_offset, int cin, int H, int W,
                                         const float* w_folded, const float* shift, int cout, float slope,
                                         void* out_hi_bf16, void* out_lo_bf16, int out_cstride, int out_coffset,
                                         void* stream) {
  CTK_REQUIRE(out_lo_bf16 != nullptr);
  return first_pool_dispatch(x, n, c_total, c_offset, cin, H, W, w_folded, shift, cout, slope, out_hi_bf16, out_cstride,
                             out_coffset, nullptr, out_lo_bf16, stream);
}

extern "C" int ctk_conv_first_pool_codes(const float* x, int n, int c_total, int c_offset, int cin, int H, int W,
                                         const float* w_folded, const float* shift, int cout, float slope,
                                         void* out_bf16, int out_cstride, int out_coffset, void* codes_u32,
                                         void* stream) {
  CTK_REQUIRE(codes_u32 != nullptr);
  return first_pool_dispatch(x, n, c_total, c_offset, cin, H, W, w_folded, shift, cout, slope, out_bf16, out_cstride,
                             out_coffset, codes_u32, nullptr, stream);
}

extern "C" size_t ctk_conv_first_raw_workspace_bytes(int cout) {
  return cout > 0 ? static_cast<size_t>(ctk::num_sms()) * 2 * cout * sizeof(float) : 0;
}

extern "C" int ctk_conv_first_raw(const float* x, int n, int c_total, int c_offset, int cin, int H, int W,
                                  const float* w, int cout, void* y_bf16, float* stats, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(x && w && y_bf16 && stats && n > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0);
  CTK_REQUIRE(c_offset >= 0 && c_offset + cin <= c_total && (reinterpret_cast<uintptr_t>(y_bf16) & 15) == 0);
  cudaStream_t s = ctk::as_stream(stream);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(y_bf16);
  if (cin == 1 && cout == 64)
    return launch_first<1, 64, true>(x, n, c_total, c_offset, H, W, w, nullptr, 0.f, out, cout, 0, stats, s, workspace,
                                     workspace_bytes);
  if (cin == 2 && cout == 128)
    return launch_first<2, 128, true>(x, n, c_total, c_offset, H, W, w, nullptr, 0.f, out, cout, 0, stats, s, workspace,
                                      workspace_bytes);
  return CTK_ERR_UNSUPPORTED;
}
