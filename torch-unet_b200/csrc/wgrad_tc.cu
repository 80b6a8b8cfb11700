// Weight gradient of the 3x3 convolutions on tcgen05 (replaces the wgrad half of aten::convolution_backward for
// nn.Conv2d at /root/reference/regression_model.py:23 and two_branch_regression.py:16,22,28):
//   dW[co][ci][ky][kx] = sum_{n,y,x} dY[n,y,x,co] * X[n, y+ky-1, x+kx-1, ci]
// GEMM view per CTA: D[co (M = 128), (ky, ci) (N = 3 x 64 = 192)] for one or two kernel columns kx, reduced over
// K = pixels.  Both operands are NHWC, i.e. contiguous along their M/N dimension: they are consumed as MN-major
// 128B-swizzled tiles exactly as TMA lands them (one 128-byte row of 64 channels per pixel, 8-pixel swizzle atoms):
//   A = dY patch  : 16x8 pixels x 128 co  -> two boxes {64 co, 8, 16}; K atom = one image row of the patch (SBO 1024),
//                   second 64-co block at LBO = 16 KiB
//   B = X halo    : 18 rows x 10 columns x 64 ci, loaded ONCE per patch.  The tap (ky, kx) is the view that starts
//                   (ky*10 + kx) halo pixels in; consecutive K atoms (image rows) are SBO = 1280 B apart, and so are the
//                   three ky views -- which makes them the three 64-element "N blocks" of one operand with LBO = 1280:
//                   a single MMA covers all three kernel rows (N = 192).
// With N = 192 an MMA reads 4 KB of A and 6 KB of B from shared memory per 96 clocks (107 B/clk against the 128 B/clk
// port; the previous N = 64 / 128 tiles needed 192 / 128 B/clk plus the TMA writes and sat at 44 / 64 % tensor activity),
// and a patch costs 55 KB of L2 traffic for 9 taps instead of 52-72 KB for 3.
// A CTA owns (co block, 64-ci block, kx group, slice): kx group 0 = columns {0, 1} (384 TMEM columns), group 1 = column 2
// (192 columns, half the work, so it gets half as many slices); it walks every `slices`-th patch, accumulating in TMEM
// with no epilogue in between; at the end every CTA stores its accumulators as one slice of partial sums in the workspace
// (plain coalesced stores) and wgrad_reduce_kernel adds the slices of each (co block, ci block) in slice order into dW
// (reference layout).  No floating-point atomics: two identical launches give bit-identical gradients.
#include "ctk_common.h"
#include "ctk_ptx.cuh"

#include <algorithm>
#include <cstdlib>

namespace {

using namespace ctk;

constexpr int kTileH = 16, kTileW = 8, kHaloW = 10, kHaloH = kTileH + 2;
constexpr int kThreads = 192;                 // warp 0 = TMA producer, warp 1 = MMA issuer + TMEM, warps 2-5 = epilogue
constexpr int kABlockBytes = 128 * 128;       // one 64-channel block of the dY patch
constexpr int kBBytes = kHaloH * kHaloW * 128;                       // 23040: the 64-channel X halo
constexpr int kStageBytes = 2 * kABlockBytes + (kBBytes + 1023) / 1024 * 1024;   // 56320
constexpr int kStages = 4;                    // 4 x 55 KB: one more patch in flight than round 1 (the 64 -> 128 layer is 66 % DRAM-bound)
constexpr int kTmemCols = 512;
constexpr int kN = 192;                       // (ky, ci) columns per kernel column kx

struct WgradParams {
  int n_img, H, W, cin, cout;
  int tiles_x, tiles_y, total_tiles;
  int co_blocks, ci_blocks, slices_a, slices_b;   // slices of kx group 0 / 1
  float* part;                                // workspace: per pair [slices_a][2][6][32][128] then [slices_b][6][32][128] fp32
  float* dw;                                  // [cout][cin][3][3] fp32
};
constexpr int kColBlock = 32 * 128;           // one 32-column TMEM block of a CTA: [i (ci)][co] floats
// floats of partial sums per (co block, ci block) pair
__host__ __device__ inline size_t pair_part_floats(int slices_a, int slices_b) {
  return static_cast<size_t>(slices_a) * 12 * kColBlock + static_cast<size_t>(slices_b) * 6 * kColBlock;
}

struct WgradSmem {
  uint64_t full[kStages], empty[kStages], done;
  uint32_t tmem_base;
};

// instruction descriptor: bf16 x bf16 -> f32, both operands MN-major
__host__ __device__ constexpr uint32_t idesc_mn(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// MN-major SW128 shared-memory descriptor: LBO = stride between 64-element MN blocks, SBO = stride between 8-row K atoms
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3fff) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x,
                const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  WgradSmem* sl = reinterpret_cast<WgradSmem*>(smem + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // unit decode: blockIdx.x = (co_blk * ci_blocks + ci_blk) * (slices_a + slices_b) + r.  Inside a pair the CTAs come in
  // triplets (A, A, B): group-0 slices 2j and 2j+1 and group-1 slice j walk the SAME patches at the same time (the B CTA
  // does two patches while each A CTA does one), so the second read of every dY / X tile is an L2 hit -- provided the
  // whole grid is resident at once (grid <= SM count, see the launch code).
  const int per_pair = p.slices_a + p.slices_b;
  const int pair = blockIdx.x / per_pair;
  const int r = blockIdx.x - pair * per_pair;
  int group, slice;
  if (p.slices_a == 2 * p.slices_b) {
    group = (r % 3) < 2 ? 0 : 1;
    slice = group == 0 ? (r / 3) * 2 + (r % 3) : r / 3;
  } else {                                            // tiny problems: plain split
    group = r < p.slices_a ? 0 : 1;
    slice = group == 0 ? r : r - p.slices_a;
  }
  const int slices = group == 0 ? p.slices_a : p.slices_b;
  const int kx0 = group == 0 ? 0 : 2, nkx = group == 0 ? 2 : 1;
  const int ci_blk = pair % p.ci_blocks;
  const int co_blk = pair / p.ci_blocks;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&sl->full[i], 1); mbar_init(&sl->empty[i], 1); }
    mbar_init(&sl->done, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_dy);
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 1) tmem_alloc<1>(&sl->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sl->tmem_base;

  if (warp == 0) {
    int stage = 0, phase = 0;
    for (int tile = slice; tile < p.total_tiles; tile += slices) {
      const int tx = tile % p.tiles_x;
      const int ty = (tile / p.tiles_x) % p.tiles_y;
      const int img = tile / (p.tiles_x * p.tiles_y);
      mbar_wait(&sl->empty[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* dst = smem + stage * kStageBytes;
        mbar_arrive_expect_tx(&sl->full[stage], static_cast<uint32_t>(2 * kABlockBytes + kBBytes));
        for (int b = 0; b < 2; ++b)
          tma_load_4d(dst + b * kABlockBytes, &tm_dy, &sl->full[stage], co_blk * 128 + b * 64, tx * kTileW,
                      ty * kTileH, img);
        tma_load_4d(dst + 2 * kABlockBytes, &tm_x, &sl->full[stage], ci_blk * 64, tx * kTileW - 1, ty * kTileH - 1, img);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = idesc_mn(128, kN);
    const uint64_t adesc0 = desc_mn_sw128(0, kABlockBytes, 1024);
    const uint64_t bdesc0 = desc_mn_sw128(0, kHaloW * 128, kHaloW * 128);   // N blocks = ky views, K atoms = image rows
    int stage = 0, phase = 0;
    bool first = true;
    for (int tile = slice; tile < p.total_tiles; tile += slices) {
      mbar_wait(&sl->full[stage], phase);
      tc_fence_after();
      const uint32_t a_base = smem_u32(smem + stage * kStageBytes);
      const uint64_t adesc_s = adesc0 | static_cast<uint64_t>(a_base >> 4);
      const uint64_t bdesc_s = bdesc0 | static_cast<uint64_t>((a_base + 2 * kABlockBytes + kx0 * 128) >> 4);
      if (elect_one()) {
        for (int j = 0; j < nkx; ++j) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {                       // 16 pixels = image rows 2ks, 2ks+1 of the patch
            umma_bf16(tmem_base + static_cast<uint32_t>(j * kN), adesc_s + (ks * 2048) / 16,
                      bdesc_s + ((ks * 2 * kHaloW + j) * 128) / 16, idesc, (first && ks == 0) ? 0u : 1u);
          }
        }
        umma_commit(&sl->empty[stage]);
      }
      __syncwarp();
      first = false;
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(&sl->done);
    __syncwarp();
  } else {
    // epilogue: lane quadrant q of TMEM = co rows q*32 .. q*32+31 of this CTA's co block
    const int q = warp & 3;
    const bool any = slice < p.total_tiles;
    // this CTA's slice of the pair's partial sums: [j (kx)][cb][i][co in block], co fastest -> 128-byte warp stores
    float* mine = p.part + static_cast<size_t>(pair) * pair_part_floats(p.slices_a, p.slices_b) +
                  (group == 0 ? static_cast<size_t>(slice) * 12 * kColBlock
                              : static_cast<size_t>(p.slices_a) * 12 * kColBlock + static_cast<size_t>(slice) * 6 * kColBlock) +
                  q * 32 + lane;
    if (any) {
      mbar_wait(&sl->done, 0);
      tc_fence_after();
    }
    for (int j = 0; j < nkx; ++j) {
      for (int cb = 0; cb < kN / 32; ++cb) {                   // column cb*32 + i = (ky = cb / 2, ci = (cb % 2) * 32 + i)
        uint32_t v[32];
        if (any) {
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + j * kN + cb * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;              // a slice without patches still owns (zero) partial sums
        }
        float* dst = mine + static_cast<size_t>(j * 6 + cb) * kColBlock;
#pragma unroll
        for (int i = 0; i < 32; ++i) dst[i * 128] = __uint_as_float(v[i]);
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, kTmemCols);
  }
}

// Second stage: dW[co][ci][ky][kx] = sum over the slices of its (co block, ci block) pair, in slice order.
// grid = (pairs * 18, 4): one CTA per 32-column block of the pair's 18 = 3 kx x 6 (ky, ci half) and per quarter of its 32
// rows i; thread = (i, co).  (One CTA for all four quarters left the single-pair layers with 18 CTAs on 148 SMs: 19 us.)
__global__ void __launch_bounds__(1024) wgrad_reduce_kernel(const WgradParams p) {
  const int pair = blockIdx.x / 18, blk = blockIdx.x % 18;
  const int kx = blk / 6, cb = blk % 6;
  const int ci_blk = pair % p.ci_blocks, co_blk = pair / p.ci_blocks;
  const int co_l = threadIdx.x & 127, i0 = threadIdx.x >> 7;      // 8 values of i per pass
  const float* base = p.part + static_cast<size_t>(pair) * pair_part_floats(p.slices_a, p.slices_b);
  const int slices = kx < 2 ? p.slices_a : p.slices_b;
  const size_t slice_stride = kx < 2 ? 12 * kColBlock : 6 * kColBlock;
  const float* src = base + (kx < 2 ? static_cast<size_t>(kx * 6 + cb) * kColBlock
                                    : static_cast<size_t>(p.slices_a) * 12 * kColBlock + static_cast<size_t>(cb) * kColBlock);
  const int ky = cb >> 1;
  {
    const int pass = blockIdx.y;
    const int i = pass * 8 + i0;
    const float* col = src + i * 128 + co_l;
    float acc = 0.f;
    int s = 0;
    for (; s + 4 <= slices; s += 4) {
      const float a = col[static_cast<size_t>(s) * slice_stride], b = col[static_cast<size_t>(s + 1) * slice_stride];
      const float c = col[static_cast<size_t>(s + 2) * slice_stride], d = col[static_cast<size_t>(s + 3) * slice_stride];
      acc = (((acc + a) + b) + c) + d;
    }
    for (; s < slices; ++s) acc += col[static_cast<size_t>(s) * slice_stride];
    const int co = co_blk * 128 + co_l, ci = ci_blk * 64 + (cb & 1) * 32 + i;
    p.dw[(static_cast<size_t>(co) * p.cin + ci) * 9 + ky * 3 + kx] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// First layer (Cin = 1 or 2): K = pixels, N = 9*Cin is too narrow for the tensor cores -> fp32 CUDA-core reduction.
// dW[co][ci][tap] = sum dY[n,y,x,co] * x[n,ci,y+ky-1,x+kx-1];  x is the fp32 NCHW input, dY is bf16 NHWC.
// Persistent blocks; thread = (4 output channels, pixel slot); the input window is staged in shared memory.
template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
wgrad_first_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x, int n_img, int c_total,
                   int c_offset, int H, int W, float* __restrict__ part) {
  constexpr int G = COUT / 4;                  // channel groups
  constexpr int SLOTS = 256 / G;               // pixels processed concurrently
  constexpr int TW = 32, TH = 8;               // pixel tile per iteration
  __shared__ float s_in[CIN][TH + 2][TW + 2 + 1];
  __shared__ float s_red[256];
  const int cg = threadIdx.x % G, slot = threadIdx.x / G;
  float acc[4][CIN * 9];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < CIN * 9; ++k) acc[j][k] = 0.f;
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  const long long total = static_cast<long long>(n_img) * tiles_x * tiles_y;
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tx = static_cast<int>(tile % tiles_x);
    const int ty = static_cast<int>((tile / tiles_x) % tiles_y);
    const int img = static_cast<int>(tile / (static_cast<long long>(tiles_x) * tiles_y));
    __syncthreads();
    for (int c = 0; c < CIN; ++c) {
      const float* plane = x + (static_cast<size_t>(img) * c_total + c_offset + c) * H * W;
      for (int i = threadIdx.x; i < (TH + 2) * (TW + 2); i += 256) {
        const int r = i / (TW + 2), q = i % (TW + 2);
        const int gy = ty * TH - 1 + r, gx = tx * TW - 1 + q;
        s_in[c][r][q] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(plane + static_cast<size_t>(gy) * W + gx) : 0.f;
      }
    }
    __syncthreads();
    for (int pix = slot; pix < TW * TH; pix += SLOTS) {
      const int py = pix / TW, px = pix % TW;
      const int gy = ty * TH + py, gx = tx * TW + px;
      if (gy >= H || gx >= W) continue;
      const uint2 raw = __ldg(reinterpret_cast<const uint2*>(
          dy + ((static_cast<size_t>(img) * H + gy) * W + gx) * COUT + cg * 4));
      const float g[4] = {__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u),
                          __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u)};
#pragma unroll
      for (int c = 0; c < CIN; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float xv = s_in[c][py + ky][px + kx];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][c * 9 + ky * 3 + kx] = fmaf(g[j], xv, acc[j][c * 9 + ky * 3 + kx]);
          }
    }
  }
  // reduce the SLOTS partial sums of every (channel, tap) through shared memory, one (j, k) plane at a time
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < CIN * 9; ++k) {
      __syncthreads();
      s_red[threadIdx.x] = acc[j][k];
      __syncthreads();
      if (slot == 0) {
        float s = 0.f;
        for (int t = 0; t < SLOTS; ++t) s += s_red[t * G + cg];
        part[static_cast<size_t>(blockIdx.x) * (COUT * CIN * 9) + (cg * 4 + j) * (CIN * 9) + k] = s;   // one row per CTA
      }
    }
}

}  // namespace

extern "C" {

// the grid is one resident wave (<= one CTA per SM), every CTA stores at most 12 column blocks of 32 x 128 floats
size_t ctk_conv3x3_wgrad_tc_workspace_bytes(int cin, int cout) {
  if (cin <= 0 || cout <= 0) return 0;
  const size_t pairs = static_cast<size_t>((cout + 127) / 128) * ((cin + 63) / 64);
  const size_t ctas = std::max(static_cast<size_t>(ctk::num_sms()), 3 * pairs);
  return ctas * 12 * kColBlock * sizeof(float);
}

int ctk_conv3x3_wgrad_tc(const void* dy_bf16, const void* x_bf16, int n, int H, int W, int cin, int cout, float* dw,
                         void* workspace, size_t workspace_bytes, void* stream) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(dy_bf16 && x_bf16 && dw && n > 0 && H > 0 && W > 0 && W % kTileW == 0 && cin % 64 == 0 && cout % 128 == 0);
  CTK_REQUIRE((reinterpret_cast<uintptr_t>(dy_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(x_bf16) & 15) == 0);
  cudaStream_t s = ctk::as_stream(stream);
  WgradParams p = {};
  p.n_img = n; p.H = H; p.W = W; p.cin = cin; p.cout = cout;
  p.tiles_x = W / kTileW;
  p.tiles_y = (H + kTileH - 1) / kTileH;
  const long long tiles = static_cast<long long>(n) * p.tiles_x * p.tiles_y;
  CTK_REQUIRE(tiles < (1ll << 30));
  p.total_tiles = static_cast<int>(tiles);
  p.co_blocks = cout / 128;
  p.ci_blocks = cin / 64;
  const int pairs = p.co_blocks * p.ci_blocks;
  // kx group 0 does two kernel columns per patch, group 1 one: twice as many slices for group 0 balances the CTAs.
  // One wave (grid <= SMs): all CTAs of a triplet run together, which is what makes their shared tiles L2 hits.
  const int s_unit = std::max(1, std::min(p.total_tiles / 2, ctk::persistent_sms() / (3 * pairs)));
  p.slices_a = std::min(p.total_tiles, 2 * s_unit);
  p.slices_b = std::min(p.total_tiles, s_unit);
  // Share of a pair's CTAs given to kx group 1 (CTK_WGRAD_B_PERMILLE, 0 = the 2 : 1 triplets above).  The B CTAs load the same
  // 55 KB per patch for half the MMA work: they are load-bound (~860 clocks per patch against 1 536 for an A CTA), so the
  // balanced split is 95 : 53, not 98 : 49.  That gives up the triplet lock-step (the second read of a tile is no longer a
  // guaranteed L2 hit); measured on one box, wgrad per step: 300 permille 3.23 ms, 333 (triplets) 2.64 - 2.68, 360 2.56,
  // 400 2.69.  Applied where a pair has at least 12 CTAs to split (pairs <= 12: the 64 -> 128 and 128 -> 256 layers).
  static const int b_permille = [] { const char* e = getenv("CTK_WGRAD_B_PERMILLE"); return e ? atoi(e) : 360; }();
  const int per_pair = ctk::persistent_sms() / pairs;
  if (b_permille > 0 && b_permille < 1000 && per_pair >= 12 && p.total_tiles >= ctk::persistent_sms()) {
    const int sb = std::max(1, std::min(per_pair - 1, per_pair * b_permille / 1000));
    p.slices_a = per_pair - sb;
    p.slices_b = sb;
  }
  p.dw = dw;
  CTK_REQUIRE_WORKSPACE(workspace, workspace_bytes, pairs * pair_part_floats(p.slices_a, p.slices_b) * sizeof(float));
  p.part = static_cast<float*>(workspace);

  CUtensorMap tm_dy, tm_x;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(cout), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                              static_cast<uint64_t>(n)};
    const uint64_t strides[3] = {static_cast<uint64_t>(cout) * 2, static_cast<uint64_t>(W) * cout * 2,
                                 static_cast<uint64_t>(H) * W * cout * 2};
    const uint32_t box[4] = {64, kTileW, kTileH, 1};
    int st = ctk::encode_tmap_bf16_sw128(&tm_dy, dy_bf16, 4, dims, strides, box);
    if (st != CTK_OK) return st;
  }
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                              static_cast<uint64_t>(n)};
    const uint64_t strides[3] = {static_cast<uint64_t>(cin) * 2, static_cast<uint64_t>(W) * cin * 2,
                                 static_cast<uint64_t>(H) * W * cin * 2};
    const uint32_t box[4] = {64, kHaloW, kHaloH, 1};
    int st = ctk::encode_tmap_bf16_sw128(&tm_x, x_bf16, 4, dims, strides, box);
    if (st != CTK_OK) return st;
  }
  const int grid = pairs * (p.slices_a + p.slices_b);
  const int smem_bytes = 1024 + kStages * kStageBytes + 256;
  CTK_CUDA_TRY(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  wgrad_tc_kernel<<<grid, kThreads, smem_bytes, s>>>(tm_dy, tm_x, p);
  int st = ctk::check_launch();
  if (st != CTK_OK) return st;
  wgrad_reduce_kernel<<<dim3(pairs * 18, 4), 1024, 0, s>>>(p);
  return ctk::check_launch();
}

size_t ctk_conv_first_wgrad_workspace_bytes(int cin, int cout) {
  return cin > 0 && cout > 0 ? static_cast<size_t>(ctk::num_sms()) * 4 * 9 * cin * cout * sizeof(float) : 0;
}

int ctk_conv_first_wgrad(const void* dy_bf16, const float* x, int n, int c_total, int c_offset, int cin, int H, int W,
                         int cout, float* dw, void* workspace, size_t workspace_bytes, void* stream) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(dy_bf16 && x && dw && n > 0 && H > 0 && W > 0 && c_offset >= 0 && c_offset + cin <= c_total);
  cudaStream_t s = ctk::as_stream(stream);
  const int grid = ctk::num_sms() * 4;
  const int cols = 9 * cin * cout;
  CTK_REQUIRE_WORKSPACE(workspace, workspace_bytes, static_cast<size_t>(grid) * cols * sizeof(float));
  float* part = static_cast<float*>(workspace);
  const __nv_bfloat16* dy = static_cast<const __nv_bfloat16*>(dy_bf16);
  if (cin == 1 && cout == 64) {
    wgrad_first_kernel<1, 64><<<grid, 256, 0, s>>>(dy, x, n, c_total, c_offset, H, W, part);
  } else if (cin == 2 && cout == 128) {
    wgrad_first_kernel<2, 128><<<grid, 256, 0, s>>>(dy, x, n, c_total, c_offset, H, W, part);
  } else {
    return CTK_ERR_UNSUPPORTED;
  }
  int st = ctk::check_launch();
  if (st != CTK_OK) return st;
  return ctk::reduce_rows_f32(part, grid, cols, cols, dw, s);
}

}  // extern "C"
