// 3x3 / stride 1 / pad 1 convolution block as an implicit GEMM on the 5th-generation tensor cores
// (tcgen05.mma, bf16 operands, fp32 accumulators in TMEM) fed by TMA, with folded BatchNorm, LeakyReLU and
// the 2x2 max-pool fused into the epilogue.  Replaces nn.Conv2d + nn.BatchNorm2d(eval) + nn.LeakyReLU +
// nn.MaxPool2d at /root/reference/regression_model.py:23-26 and two_branch_regression.py:16-19,22-25,28-31.
//
// GEMM view: D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * Wt[tap][cout][cin].
//   M tile  = 128 output pixels per CTA = a 16-row x 8-column patch of one image (one pixel per TMEM lane)
//   N tile  = kBlockN output channels (one fp32 TMEM column each)
//   K loop  = (cin / 64) chunks x 9 taps x 4 UMMA_K=16 steps
// A operand: the (16+2) x (8+2) input halo of the patch is loaded ONCE per 64-channel chunk by a single TMA
// box (out-of-bounds = zero fill gives the conv padding for free) into 128B-swizzled shared memory, one
// 128-byte row per pixel.  The nine taps are nine *views* of that buffer: the UMMA descriptor for tap
// (ky,kx) starts (ky*10 + kx) rows into the halo and strides one halo row (1280 B) per 8-pixel core-matrix
// group.  This cuts L2->SM traffic for A by 9x against a per-tap im2col load.
// B operand: per (chunk, tap) a [kBlockN cout x 64 cin] K-major tile of the tap-major packed weights.
//
// kCtaGroup = 2 runs CTA pairs (cta_group::2): one tcgen05.mma drives both SMs with M = 256 (each CTA owns 128
// pixels and their accumulators) and every CTA loads only HALF of the B tile, which halves the shared-memory
// bytes each SM has to read per MMA -- with one CTA and N = 128 the operand reads alone are 128 B/clk/SM.
//
// Warp roles (384 threads): warp 0 = B producer, warp 3 = A producer (one lane each), warp 1 = MMA issuer
// (one lane, pair leader only; fully unrolled taps, a whole chunk per elect when the weights are resident),
// warp 2 = TMEM allocator, warps 4-11 = epilogue, compiled per flavour:
//   eval        TMEM -> registers -> scale/shift -> 2x2 max via two butterfly shuffle stages -> LeakyReLU -> 16-byte
//               NHWC stores of the pooled tile
//   raw(+stats) bf16 pack -> per-warp 128B-swizzled staging tile (32 pixels x 64 channels) -> one TMA store; with stats
//               also per-channel sum / sum of squares through a per-warp shared-memory transpose, totals in registers
//   eval split  fp32-class: pool and LeakyReLU in fp32, (hi, lo) bf16 pair stores
// Each epilogue warp owns a TMEM lane quadrant (warp % 4) and consecutive pairs of 32-column blocks.
// Accumulators are double buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "ctk_common.h"
#include "ctk_ptx.cuh"

#include <algorithm>

namespace {

using namespace ctk;

constexpr int kTileH = 16;
constexpr int kTileW = 8;
constexpr int kHaloH = kTileH + 2;
constexpr int kHaloW = kTileW + 2;      // dense halo pitch: 10 pixels = 1280 B per row
constexpr int kKC = 64;                 // channels per K chunk = one 128-byte swizzle row
constexpr int kAStages = 4;
constexpr int kAccStages = 2;
constexpr int kEpiWarps = 8;               // two warps per TMEM lane quadrant, interleaved over 32-column blocks
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr uint32_t kABytes = kHaloH * kHaloW * 128;                       // 23040
constexpr int kAStageBytes = ((kABytes + 1023) / 1024) * 1024;            // 23552

// Epilogue flavours (compile time, so the hot loop carries no run-time mode predicates):
enum : int {
  kEpiEvalPool = 0,   // folded BN -> 2x2 max-pool -> LeakyReLU, pooled bf16 store (the inference block)
  kEpiEvalAny = 1,    // folded BN with run-time pool / activation switches (CTK_CONV_NO_POOL / CTK_CONV_NO_ACT)
  kEpiRaw = 2,        // bf16 store of the raw accumulators (dgrad)
  kEpiRawStats = 3,   // raw store + per-channel sum / sum of squares of the fp32 accumulators (train-mode forward)
  kEpiEvalPoolSplit = 4   // fp32-class inference: BN, pool and LeakyReLU in fp32, output stored as bf16 (hi, lo) pairs
};
constexpr int kScratchPitch = 36;                                  // floats; 16-byte aligned rows, conflict-free both ways
constexpr int kScratchBytes = kEpiWarps * 32 * kScratchPitch * 4;  // per-warp transpose tile for the statistics
constexpr int kCtrlBytes = 8192;

template <int kCtaGroup, int kBlockN, int kEpi>
struct Cfg {
  static constexpr int kBRows = kBlockN / kCtaGroup;          // weight rows this CTA loads per (chunk, tap)
  static constexpr int kBStageBytes = kBRows * 128;
  // raw modes with >= 64 channels per epilogue warp: the bf16 tile is staged per warp (32 pixels x 128 B, 128B-swizzled) and
  // written by TMA -- a lane's 64-byte direct stores kept their source registers busy until the store path drained
  // (ncu: the next tcgen05.ld / register clears stalled on them) and touched 32 half-lines per instruction
  static constexpr bool kStage = (kEpi == kEpiRaw || kEpi == kEpiRawStats) && kBlockN >= 128;
  static constexpr int kStageBytes = kStage ? kEpiWarps * 4096 : 0;
  static constexpr int kExtra = kStageBytes + (kEpi == kEpiRawStats ? kScratchBytes : 0);
  static constexpr int kAStg = (kStage && kEpi == kEpiRawStats) ? 3 : kAStages;     // activation halo stages
  // ring of weight stages; when the whole layer's share (chunks x 9 stages) fits, the weights are loaded ONCE and stay
  // resident (b_resident): 64-channel layers otherwise re-stream 74 KB per tile per CTA from L2 and stall on it
  static constexpr int kBStages =
      std::min(18, (227 * 1024 - 1024 - kCtrlBytes - kExtra - kAStg * kAStageBytes) / kBStageBytes);
  static constexpr int kTmemCols = kAccStages * kBlockN;
  static constexpr int kSmemBytes = 1024 + kAStg * kAStageBytes + kBStages * kBStageBytes + kCtrlBytes + kExtra;
};

struct ConvParams {
  int n_img, H, W, cin, cout;
  int tiles_x, tiles_y, tiles_n, spatial_tiles, total_work;
  FastDiv div_n, div_x, div_y;
  int pool, act;
  float slope;
  const float* scale;   // nullptr = identity (raw conv output)
  const float* shift;
  float* stats;         // nullptr, or [gridDim.x][2*cout]: ONE ROW PER CTA of per-channel partial sum / sum of squares of
                        // the raw fp32 accumulators (added up in a fixed order by ctk::reduce_rows_f32: deterministic)
  float* stats_out;     // host side only: [2*cout] totals
  __nv_bfloat16* out;
  __nv_bfloat16* out_lo;   // kEpiEvalPoolSplit: low halves, same pixel / channel addressing as `out`
  int out_cstride, out_coffset;
  int b_resident;       // chunks * 9 <= kBStages: every (chunk, tap) weight stage is loaded once and never recycled
  int cin_phys;         // channels of the activation tensor in memory (= cin, or 2/3 cin for the split layout)
  int a_wrap;           // split layout: K chunk c >= a_wrap re-reads physical chunk c - a_wrap ([hi | lo | hi] from [hi | lo])
};

struct TileCoord {
  int img, y0, x0, n0;
};

// work item -> (pair of spatial tiles, N tile); N tile fastest so that CTAs running together share the halo in L2
template <int kCtaGroup, int kBlockN>
__device__ __forceinline__ TileCoord decode_work(const ConvParams& p, int work, int rank) {
  TileCoord t;
  const uint32_t wq = fdiv(static_cast<uint32_t>(work), p.div_n);
  const int nt = work - static_cast<int>(wq) * p.tiles_n;
  const uint32_t sp = wq * kCtaGroup + rank;          // may be >= spatial_tiles for the padded last pair: img >= n_img
  const uint32_t q1 = fdiv(sp, p.div_x);
  const int tx = static_cast<int>(sp - q1 * p.tiles_x);
  const uint32_t q2 = fdiv(q1, p.div_y);
  const int ty = static_cast<int>(q1 - q2 * p.tiles_y);
  t.img = static_cast<int>(q2);
  t.y0 = ty * kTileH;
  t.x0 = tx * kTileW;
  t.n0 = nt * kBlockN;
  return t;
}

template <int kCtaGroup, int kBlockN, int kEpi>
struct SmemLayout {
  using C = Cfg<kCtaGroup, kBlockN, kEpi>;
  uint64_t a_full[kAStages], a_empty[kAStages];
  uint64_t b_full[C::kBStages], b_empty[C::kBStages];
  uint64_t acc_full[kAccStages], acc_empty[kAccStages];
  uint32_t tmem_base;
  uint32_t pad[3];
  // eval: folded-BN scale / shift of ALL output channels (cout <= 512), staged once per CTA;
  // train (raw + statistics): the same arrays hold the per-CTA partial sum / sum of squares, flushed once at the end
  alignas(16) float ch_a[512];
  alignas(16) float ch_b[512];
};

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct EpiCtx {
  int lane, y, x, Hp, Wp;
  bool valid;
  uint32_t sc_addr, sh_addr;      // shared-space addresses of ch_a / ch_b
  uint32_t scratch_addr;          // this warp's transpose tile (statistics)
  uint32_t stage_addr;            // this warp's 32-pixel x 128-byte output staging tile (raw modes, TMA store)
  int stage_slot;                 // which 64-byte half of the staged rows the current block fills
  __nv_bfloat162 slope2;
};

// One 32-lane x 32-column block of fp32 accumulators (lane = output pixel, v[i] = channel ch0 + i of the tile).
template <int kEpi>
__device__ __forceinline__ void epilogue_block(const ConvParams& p, const EpiCtx& e, const TileCoord& t, int ch0,
                                               const uint32_t (&v)[32], float& acc_s, float& acc_q) {
  const int lane = e.lane;
  uint32_t pk[16];
  if constexpr (kEpi == kEpiRaw || kEpi == kEpiRawStats) {
    if constexpr (kEpi == kEpiRawStats) {
      // batch statistics of the raw conv output: transpose the 32 x 32 block through this warp's scratch tile so that
      // lane c can sum channel c over the warp's 32 pixels (8 STS.128 + 32 LDS + 64 math instead of a 31-shuffle
      // butterfly per moment)
      const uint32_t row = e.scratch_addr + static_cast<uint32_t>(lane * kScratchPitch * 4);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = e.valid ? __uint_as_float(v[4 * j]) : 0.f, b = e.valid ? __uint_as_float(v[4 * j + 1]) : 0.f;
        const float c = e.valid ? __uint_as_float(v[4 * j + 2]) : 0.f, d = e.valid ? __uint_as_float(v[4 * j + 3]) : 0.f;
        sts128(row + j * 16, a, b, c, d);
      }
      __syncwarp();
      float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
      const uint32_t col = e.scratch_addr + static_cast<uint32_t>(lane * 4);
#pragma unroll
      for (int l = 0; l < 32; l += 2) {
        const float x0 = lds32(col + l * kScratchPitch * 4);
        const float x1 = lds32(col + (l + 1) * kScratchPitch * 4);
        s0 += x0; q0 = fmaf(x0, x0, q0);
        s1 += x1; q1 = fmaf(x1, x1, q1);
      }
      __syncwarp();
      // a CTA keeps ONE N tile for all its work items (launch_conv makes the work stride a multiple of tiles_n), so a
      // warp meets the same channels every tile: the totals stay in registers and are combined once, in a fixed order,
      // at the end of the kernel -- no shared or global floating-point atomics anywhere in the statistics
      acc_s += s0 + s1;
      acc_q += q0 + q1;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
    if (e.stage_addr != 0) {
      // row = this lane's pixel (128 B = 64 channels of two consecutive blocks), 16-byte chunks XOR-swizzled like TMA's
      const uint32_t row = e.stage_addr + static_cast<uint32_t>(lane * 128);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t chunk = static_cast<uint32_t>(((e.stage_slot * 4 + i) ^ (lane & 7)) * 16);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + chunk), "r"(pk[4 * i]), "r"(pk[4 * i + 1]),
                     "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3])
                     : "memory");
      }
    } else if (e.valid) {
      __nv_bfloat16* dst = p.out +
          (static_cast<size_t>(t.img) * p.H * p.W + static_cast<size_t>(e.y) * p.W + e.x) * p.out_cstride +
          p.out_coffset + t.n0 + ch0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<uint4*>(dst)[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
    }
    return;
  } else if constexpr (kEpi == kEpiEvalPoolSplit) {
    const uint32_t sc = e.sc_addr + static_cast<uint32_t>((t.n0 + ch0) * 4);
    const uint32_t sh = e.sh_addr + static_cast<uint32_t>((t.n0 + ch0) * 4);
    float f[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 a = lds128(sc + j * 16), b = lds128(sh + j * 16);
      f[4 * j] = fmaf(__uint_as_float(v[4 * j]), a.x, b.x);
      f[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), a.y, b.y);
      f[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), a.z, b.z);
      f[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), a.w, b.w);
    }
    const bool odd_x = (lane & 1) != 0, odd_y = (lane & 8) != 0;
    float q[16], r[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float send = odd_x ? f[i] : f[16 + i];
      const float keep = odd_x ? f[16 + i] : f[i];
      q[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = odd_y ? q[i] : q[8 + i];
      const float keep = odd_y ? q[8 + i] : q[i];
      r[i] = leaky(fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 8)), p.slope);
    }
    if (e.valid) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(r[2 * i]), h1 = __float2bfloat16_rn(r[2 * i + 1]);
        hi[i] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
        lo[i] = pack_bf16x2(r[2 * i] - __bfloat162float(h0), r[2 * i + 1] - __bfloat162float(h1));
      }
      const int ch = t.n0 + ch0 + (odd_x ? 16 : 0) + (odd_y ? 8 : 0);
      const size_t off =
          (static_cast<size_t>(t.img) * e.Hp * e.Wp + static_cast<size_t>(e.y >> 1) * e.Wp + (e.x >> 1)) * p.out_cstride +
          p.out_coffset + ch;
      *reinterpret_cast<uint4*>(p.out + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(p.out_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    return;
  } else {
    const uint32_t sc = e.sc_addr + static_cast<uint32_t>((t.n0 + ch0) * 4);
    const uint32_t sh = e.sh_addr + static_cast<uint32_t>((t.n0 + ch0) * 4);
    const bool act_first = kEpi == kEpiEvalAny && p.act && !p.pool;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 a = lds128(sc + j * 16), b = lds128(sh + j * 16);
      float f0 = fmaf(__uint_as_float(v[4 * j]), a.x, b.x), f1 = fmaf(__uint_as_float(v[4 * j + 1]), a.y, b.y);
      float f2 = fmaf(__uint_as_float(v[4 * j + 2]), a.z, b.z), f3 = fmaf(__uint_as_float(v[4 * j + 3]), a.w, b.w);
      if (act_first) { f0 = leaky(f0, p.slope); f1 = leaky(f1, p.slope); f2 = leaky(f2, p.slope); f3 = leaky(f3, p.slope); }
      pk[2 * j] = pack_bf16x2(f0, f1);
      pk[2 * j + 1] = pack_bf16x2(f2, f3);
    }
    if (kEpi == kEpiEvalPool || p.pool) {
      // 2x2 max over lanes {l, l^1, l^8}: each butterfly stage halves the channels a lane keeps
      const bool odd_x = (lane & 1) != 0;
      uint32_t q[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t send = odd_x ? pk[i] : pk[8 + i];
        const uint32_t keep = odd_x ? pk[8 + i] : pk[i];
        q[i] = max_bf16x2(keep, __shfl_xor_sync(0xffffffffu, send, 1));
      }
      const bool odd_y = (lane & 8) != 0;
      uint4 o;
      uint32_t* ov = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t send = odd_y ? q[i] : q[4 + i];
        const uint32_t keep = odd_y ? q[4 + i] : q[i];
        ov[i] = max_bf16x2(keep, __shfl_xor_sync(0xffffffffu, send, 8));
        if (kEpi == kEpiEvalPool || p.act) ov[i] = leaky_bf16x2(ov[i], e.slope2);   // after the pool: 4x fewer elements
      }
      if (e.valid) {
        const int ch = t.n0 + ch0 + (odd_x ? 16 : 0) + (odd_y ? 8 : 0);
        __nv_bfloat16* dst = p.out +
            (static_cast<size_t>(t.img) * e.Hp * e.Wp + static_cast<size_t>(e.y >> 1) * e.Wp + (e.x >> 1)) * p.out_cstride +
            p.out_coffset + ch;
        *reinterpret_cast<uint4*>(dst) = o;
      }
    } else if (e.valid) {
      __nv_bfloat16* dst = p.out +
          (static_cast<size_t>(t.img) * p.H * p.W + static_cast<size_t>(e.y) * p.W + e.x) * p.out_cstride +
          p.out_coffset + t.n0 + ch0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<uint4*>(dst)[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
    }
  }
}

template <int kCtaGroup, int kBlockN, int kEpi>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                  const __grid_constant__ CUtensorMap tm_out, const ConvParams p) {
  using C = Cfg<kCtaGroup, kBlockN, kEpi>;
  using SL = SmemLayout<kCtaGroup, kBlockN, kEpi>;
  constexpr bool kChunkAcc = kEpi == kEpiEvalPoolSplit;     // fresh accumulator per K chunk, summed in registers
  static_assert(!kChunkAcc || kBlockN <= 128, "chunked accumulation keeps kBlockN / 2 running sums per epilogue thread");
  static_assert(sizeof(SL) <= kCtrlBytes, "barrier block too large");
  static_assert(C::kBStages >= 4, "too few weight stages");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + C::kAStg * kAStageBytes;
  SL* sl = reinterpret_cast<SL*>(b_smem + C::kBStages * C::kBStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int chunks = p.cin / kKC;
  const int rank = kCtaGroup == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int work0 = blockIdx.x / kCtaGroup;
  const int work_stride = gridDim.x / kCtaGroup;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::kAStg; ++i) { mbar_init(&sl->a_full[i], 1); mbar_init(&sl->a_empty[i], 1); }
    for (int i = 0; i < C::kBStages; ++i) { mbar_init(&sl->b_full[i], 1); mbar_init(&sl->b_empty[i], 1); }
    for (int i = 0; i < kAccStages; ++i) { mbar_init(&sl->acc_full[i], 1); mbar_init(&sl->acc_empty[i], kEpiWarps * kCtaGroup); }
    fence_mbar_init();
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
  }
  if (warp == 2) tmem_alloc<kCtaGroup>(&sl->tmem_base, C::kTmemCols);
  for (int i = threadIdx.x; i < 512; i += kThreads) {
    if constexpr (kEpi == kEpiEvalPool || kEpi == kEpiEvalAny || kEpi == kEpiEvalPoolSplit) {
      sl->ch_a[i] = i < p.cout ? __ldg(p.scale + i) : 0.f;
      sl->ch_b[i] = i < p.cout ? __ldg(p.shift + i) : 0.f;
    } else {
      sl->ch_a[i] = 0.f;
      sl->ch_b[i] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kCtaGroup == 2) cluster_sync_all();   // peer barriers / TMEM must exist before anything remote
  tc_fence_after();
  const uint32_t tmem_base = sl->tmem_base;

  // Every role loop below is executed by its WHOLE warp with warp-uniform control flow; only the instruction that
  // must be issued once (TMA, tcgen05.mma, commit) sits under elect_one().  A loop owned by a single divergent lane
  // makes the compiler shuttle every descriptor through ELECT / R2UR.BROADCAST / BRA.U.ANY sequences (~130 SASS
  // instructions per tap, measured 840 clk per tap against 256 clk of MMA work on the 64->128 layer).
  if (warp == 3) {
    // ---------------- A producer: one halo box per (work item, chunk); pair members signal the leader's barrier
    int stage = 0, phase = 0;
    for (int work = work0; work < p.total_work; work += work_stride) {
      const TileCoord t = decode_work<kCtaGroup, kBlockN>(p, work, rank);
      for (int c = 0; c < chunks; ++c) {
        mbar_wait(&sl->a_empty[stage], phase ^ 1);
        uint8_t* dst = a_smem + stage * kAStageBytes;
        const int cc = (p.a_wrap > 0 && c >= p.a_wrap) ? c - p.a_wrap : c;
        if (elect_one()) {
          if constexpr (kCtaGroup == 1) {
            mbar_arrive_expect_tx(&sl->a_full[stage], kABytes);
            tma_load_4d(dst, &tm_a, &sl->a_full[stage], cc * kKC, t.x0 - 1, t.y0 - 1, t.img);
          } else {
            if (rank == 0) mbar_arrive_expect_tx(&sl->a_full[stage], 2 * kABytes);
            tma_load_4d_pair(dst, &tm_a, mapa_shared(smem_u32(&sl->a_full[stage]), 0), cc * kKC, t.x0 - 1, t.y0 - 1,
                             t.img);
          }
        }
        __syncwarp();
        if (++stage == C::kAStg) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 0) {
    // ---------------- B producer: this CTA's share of the weight tile per (work item, chunk, tap)
    int stage = 0, phase = 0;
    for (int work = work0; work < p.total_work; work += work_stride) {
      const TileCoord t = decode_work<kCtaGroup, kBlockN>(p, work, rank);
      for (int c = 0; c < chunks; ++c) {
        for (int tap = 0; tap < 9; ++tap) {
          if (!p.b_resident) mbar_wait(&sl->b_empty[stage], phase ^ 1);
          uint8_t* dst = b_smem + stage * C::kBStageBytes;
          const int row = tap * p.cout + t.n0 + rank * C::kBRows;
          if (elect_one()) {
            if constexpr (kCtaGroup == 1) {
              mbar_arrive_expect_tx(&sl->b_full[stage], C::kBStageBytes);
              tma_load_2d(dst, &tm_b, &sl->b_full[stage], c * kKC, row);
            } else {
              if (rank == 0) mbar_arrive_expect_tx(&sl->b_full[stage], 2 * C::kBStageBytes);
              tma_load_2d_pair(dst, &tm_b, mapa_shared(smem_u32(&sl->b_full[stage]), 0), c * kKC, row);
            }
          }
          __syncwarp();
          if (++stage == C::kBStages) { stage = 0; phase ^= 1; }
        }
      }
      if (p.b_resident) break;      // resident weights (one N tile): loaded with the first work item, reused by all
    }
  } else if (warp == 1 && rank == 0) {
    // ---------------- MMA issuer (pair leader only)
    // The issue loop is on the critical path of the narrow layers: a (256 x 64 x 16) MMA retires in 32 clocks, so the ~80
    // instructions a rolled per-tap iteration used to spend on run-time tap arithmetic, constant-bank reloads and
    // descriptor moves (measured: 300+ clk per 4 MMAs) starved the tensor pipe.  Taps and K steps are fully unrolled
    // (descriptors differ by compile-time constants), and with resident weights a whole chunk -- 36 MMAs and its commits
    // -- is issued under a single elect.
    constexpr uint32_t idesc = umma_idesc_bf16_f32(128 * kCtaGroup, kBlockN);
    constexpr uint32_t sbo = kHaloW * 128;   // one halo row per 8-pixel core-matrix group
    // descriptor templates: everything but the 14-bit start-address field is loop invariant
    const uint64_t adesc0 = umma_smem_desc_sw128(0, sbo, 0);
    const uint64_t bdesc0 = umma_smem_desc_sw128(0, 1024, 0);
    const uint32_t b_smem_addr = smem_u32(b_smem);
    auto mma = [&](uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accum) {
      if constexpr (kCtaGroup == 1) umma_bf16(d_tmem, adesc, bdesc, idesc, accum);
      else umma_bf16_pair(d_tmem, adesc, bdesc, idesc, accum);
    };
    auto commit = [&](uint64_t* bar) {
      if constexpr (kCtaGroup == 1) umma_commit(bar);
      else umma_commit_pair(bar);
    };
    int as = 0, aphase = 0, bs = 0, bphase = 0;
    int it = 0;
    int ia = 0;                              // accumulator stages handed to the epilogue so far
    const bool resident = p.b_resident != 0;
    for (int work = work0; work < p.total_work; work += work_stride, ++it) {
      int acc = 0;
      uint32_t d_tmem = 0;
      for (int c = 0; c < chunks; ++c) {
        if (kChunkAcc || c == 0) {
          // kChunkAcc: every 64-channel chunk (36 MMAs) gets a fresh accumulator stage; the epilogue adds the stages up
          // in fp32 registers.  tcgen05 accumulates with truncation (about -2^-24 relative per MMA step, measured:
          // tools/probe_accum_bias.py), which over the 200-900 steps of a whole fp32-class tile is 1e-5..5e-5.
          acc = ia & 1;
          mbar_wait(&sl->acc_empty[acc], ((ia >> 1) & 1) ^ 1);
          tc_fence_after();
          d_tmem = tmem_base + static_cast<uint32_t>(acc * kBlockN);
          ++ia;
        }
        mbar_wait(&sl->a_full[as], aphase);
        tc_fence_after();
        // The 128B swizzle is a function of the absolute shared-memory address bits (verified on B200: base-offset
        // field 0, any 128-byte-aligned start, any multiple-of-128 group stride), so a shifted window of the
        // TMA-written halo is a valid K-major operand as is.  Stage bases are 1024-byte aligned and every offset below
        // stays inside the stage, so adding (offset >> 4) never carries out of the descriptor's address field.
        const uint64_t adesc_c = adesc0 | static_cast<uint64_t>(smem_u32(a_smem + as * kAStageBytes) >> 4);
        const bool last_chunk = kChunkAcc || c == chunks - 1;
        const bool first_chunk = kChunkAcc || c == 0;
        if (resident && it > 0) {
          // weights of every (chunk, tap) are in place since the first tile: one straight-line burst per chunk
          const uint64_t bdesc_c = bdesc0 | static_cast<uint64_t>((b_smem_addr + c * 9 * C::kBStageBytes) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int s = 0; s < kKC / 16; ++s) {
                const uint32_t accum = (tap | s) != 0 ? 1u : (first_chunk ? 0u : 1u);
                mma(d_tmem, adesc_c + (((tap / 3) * kHaloW + tap % 3) * 128 + s * 32) / 16,
                    bdesc_c + (tap * C::kBStageBytes + s * 32) / 16, accum);
              }
            }
            commit(&sl->a_empty[as]);
            if (last_chunk) commit(&sl->acc_full[acc]);
          }
          __syncwarp();
        } else {
          // streamed weights (or the first tile of a resident layer): one kernel row -- three taps, 12 MMAs -- per elect
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            int st[3];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              if (resident) bs = c * 9 + ky * 3 + kx;
              st[kx] = bs;
              mbar_wait(&sl->b_full[bs], bphase);
              if (!resident && ++bs == C::kBStages) { bs = 0; bphase ^= 1; }
            }
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const uint64_t adesc = adesc_c + ((ky * kHaloW + kx) * 128) / 16;
                const uint64_t bdesc = bdesc0 | static_cast<uint64_t>((b_smem_addr + st[kx] * C::kBStageBytes) >> 4);
#pragma unroll
                for (int s = 0; s < kKC / 16; ++s) {
                  const uint32_t accum = (ky | kx | s) != 0 ? 1u : (first_chunk ? 0u : 1u);
                  mma(d_tmem, adesc + 2 * s, bdesc + 2 * s, accum);      // +32 bytes per UMMA_K step
                }
                if (!resident) commit(&sl->b_empty[st[kx]]);
              }
              if (ky == 2) commit(&sl->a_empty[as]);
              if (ky == 2 && last_chunk) commit(&sl->acc_full[acc]);
            }
            __syncwarp();
          }
        }
        if (++as == C::kAStg) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: thread <-> TMEM lane <-> output pixel of this CTA's tile; the two warps that share a
    // lane quadrant (warp % 4) take alternate PAIRS of 32-column blocks.  TMEM loads are software pipelined: the block
    // after the one being processed is already in flight, and the accumulator stage is released as soon as the last load
    // landed.
    const int ew = warp & 3;
    const int half = (warp - 4) >> 2;
    const int m = ew * 32 + lane;
    const int r = m >> 3, cpx = m & 7;
    const int et = threadIdx.x - 128;
    constexpr int kBlocks = kBlockN / 32;
    constexpr int kPerWarp = kBlocks / 2;
    static_assert(kBlocks % 2 == 0 && kEpiWarps == 8, "two epilogue warps per lane quadrant");
    EpiCtx e;
    e.lane = lane;
    e.Hp = p.H >> 1;
    e.Wp = p.W >> 1;
    e.sc_addr = smem_u32(sl->ch_a);
    e.sh_addr = smem_u32(sl->ch_b);
    e.scratch_addr = smem_u32(reinterpret_cast<uint8_t*>(sl) + kCtrlBytes) + static_cast<uint32_t>(C::kStageBytes) +
                     static_cast<uint32_t>((warp - 4) * 32 * kScratchPitch * 4);
    e.stage_addr = C::kStage ? smem_u32(reinterpret_cast<uint8_t*>(sl) + kCtrlBytes) + static_cast<uint32_t>((warp - 4) * 4096) : 0u;
    e.stage_slot = 0;
    float stat_s[kPerWarp], stat_q[kPerWarp];
#pragma unroll
    for (int i = 0; i < kPerWarp; ++i) { stat_s[i] = 0.f; stat_q[i] = 0.f; }
    // block i of this warp: consecutive PAIRS of 32-column blocks (64 channels = one 128-byte staged row)
    auto block_of = [&](int i) { return kBlocks == 2 ? half : (i >> 1) * 4 + half * 2 + (i & 1); };
    e.slope2 = __float2bfloat162_rn(p.slope);
    int it = 0;
    if constexpr (kChunkAcc) {
      // fp32-class path: one accumulator stage per K chunk, added up here in fp32 (round to nearest) registers
      int ia = 0;
      for (int work = work0; work < p.total_work; work += work_stride) {
        const TileCoord t = decode_work<kCtaGroup, kBlockN>(p, work, rank);
        e.y = t.y0 + r;
        e.x = t.x0 + cpx;
        e.valid = t.img < p.n_img && e.y < p.H && e.x < p.W;
        float sums[kPerWarp][32];
#pragma unroll
        for (int i = 0; i < kPerWarp; ++i)
#pragma unroll
          for (int j = 0; j < 32; ++j) sums[i][j] = 0.f;
        for (int c = 0; c < chunks; ++c, ++ia) {
          const int acc = ia & 1;
          mbar_wait(&sl->acc_full[acc], (ia >> 1) & 1);
          tc_fence_after();
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * kBlockN);
#pragma unroll
          for (int i = 0; i < kPerWarp; ++i) {
            uint32_t v[32];
            tmem_ld_32x32(taddr + block_of(i) * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) sums[i][j] += __uint_as_float(v[j]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kCtaGroup == 1) mbar_arrive(&sl->acc_empty[acc]);
            else mbar_arrive_cluster(mapa_shared(smem_u32(&sl->acc_empty[acc]), 0));
          }
        }
#pragma unroll
        for (int i = 0; i < kPerWarp; ++i) {
          uint32_t v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(sums[i][j]);
          float unused_s = 0.f, unused_q = 0.f;
          epilogue_block<kEpi>(p, e, t, block_of(i) * 32, v, unused_s, unused_q);
        }
      }
    }
    for (int work = work0; !kChunkAcc && work < p.total_work; work += work_stride, ++it) {
      const TileCoord t = decode_work<kCtaGroup, kBlockN>(p, work, rank);
      const int acc = it & 1;
      const int acc_phase = (it >> 1) & 1;
      e.y = t.y0 + r;
      e.x = t.x0 + cpx;
      e.valid = t.img < p.n_img && e.y < p.H && e.x < p.W;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * kBlockN);
      mbar_wait(&sl->acc_full[acc], acc_phase);
      tc_fence_after();
      auto release = [&]() {
        // every TMEM read of this warp for the tile has completed: one arrival per warp on the (leader's) barrier
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kCtaGroup == 1) mbar_arrive(&sl->acc_empty[acc]);
          else mbar_arrive_cluster(mapa_shared(smem_u32(&sl->acc_empty[acc]), 0));
        }
      };
      uint32_t va[32], vb[32];
      tmem_ld_32x32(taddr + block_of(0) * 32, va);
#pragma unroll
      for (int i = 0; i < kPerWarp; i += 2) {
        tmem_ld_wait();
        if (i + 1 < kPerWarp) tmem_ld_32x32(taddr + block_of(i + 1) * 32, vb);
        else release();
        if constexpr (C::kStage) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous store read the tile
          __syncwarp();
          e.stage_slot = 0;
        }
        epilogue_block<kEpi>(p, e, t, block_of(i) * 32, va, stat_s[i], stat_q[i]);
        if (i + 1 < kPerWarp) {
          tmem_ld_wait();
          if (i + 2 < kPerWarp) tmem_ld_32x32(taddr + block_of(i + 2) * 32, va);
          else release();
          e.stage_slot = 1;
          epilogue_block<kEpi>(p, e, t, block_of(i + 1) * 32, vb, stat_s[i + 1 < kPerWarp ? i + 1 : 0],
                               stat_q[i + 1 < kPerWarp ? i + 1 : 0]);
        }
        if constexpr (C::kStage) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            asm volatile(
                "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                    reinterpret_cast<uint64_t>(&tm_out)),
                "r"(e.stage_addr), "r"(p.out_coffset + t.n0 + block_of(i) * 32), "r"(t.x0), "r"(t.y0 + ew * 4), "r"(t.img)
                : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
    }
    if constexpr (C::kStage) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores must land before the CTA exits
      __syncwarp();
    }
    if constexpr (kEpi == kEpiRawStats) {
      // the four lane quadrants (warp % 4) of a channel are added in quadrant order through the (now idle) transpose
      // scratch: quad[ew][moment][kBlockN]; then this CTA's row of the partial-sum matrix is written with plain stores
      const uint32_t quad = smem_u32(reinterpret_cast<uint8_t*>(sl) + kCtrlBytes) + static_cast<uint32_t>(C::kStageBytes);
      static_assert(4 * 2 * kBlockN * 4 <= kScratchBytes, "quadrant totals must fit the scratch region");
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
#pragma unroll
      for (int i = 0; i < kPerWarp; ++i) {
        const uint32_t ch = static_cast<uint32_t>(block_of(i) * 32 + lane);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(quad + ((ew * 2 + 0) * kBlockN + ch) * 4), "f"(stat_s[i]) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(quad + ((ew * 2 + 1) * kBlockN + ch) * 4), "f"(stat_q[i]) : "memory");
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      const int n0 = decode_work<kCtaGroup, kBlockN>(p, work0, rank).n0;
      float* row = p.stats + static_cast<size_t>(blockIdx.x) * 2 * p.cout;
      for (int i = et; i < 2 * p.cout; i += 32 * kEpiWarps) {
        const int mom = i >= p.cout ? 1 : 0;
        const int c = i - mom * p.cout - n0;
        float v = 0.f;
        if (c >= 0 && c < kBlockN) {
          const uint32_t a = quad + static_cast<uint32_t>((mom * kBlockN + c) * 4);
          v = ((lds32(a) + lds32(a + 2 * kBlockN * 4)) + lds32(a + 4 * kBlockN * 4)) + lds32(a + 6 * kBlockN * 4);
        }
        row[i] = v;
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if constexpr (kCtaGroup == 2) cluster_sync_all();   // nobody leaves while the pair still touches its smem / TMEM
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kCtaGroup>(tmem_base, C::kTmemCols);
  }
}

template <int kCtaGroup, int kBlockN, int kEpi>
int launch_conv(const void* x_bf16, const void* w_packed_bf16, ConvParams p, cudaStream_t stream) {
  using C = Cfg<kCtaGroup, kBlockN, kEpi>;
  p.tiles_n = p.cout / kBlockN;
  p.div_n = make_fastdiv(static_cast<uint32_t>(p.tiles_n));
  p.div_x = make_fastdiv(static_cast<uint32_t>(p.tiles_x));
  p.div_y = make_fastdiv(static_cast<uint32_t>(p.tiles_y));
  p.b_resident = (p.tiles_n == 1 && (p.cin / kKC) * 9 <= C::kBStages) ? 1 : 0;
  const long long work = static_cast<long long>((p.spatial_tiles + kCtaGroup - 1) / kCtaGroup) * p.tiles_n;
  if (work >= (1ll << 31)) return CTK_ERR_BAD_ARG;
  p.total_work = static_cast<int>(work);

  CUtensorMap tm_a, tm_b;
  {
    const uint64_t cp = static_cast<uint64_t>(p.cin_phys);
    const uint64_t dims[4] = {cp, static_cast<uint64_t>(p.W), static_cast<uint64_t>(p.H), static_cast<uint64_t>(p.n_img)};
    const uint64_t strides[3] = {cp * 2, static_cast<uint64_t>(p.W) * cp * 2, static_cast<uint64_t>(p.H) * p.W * cp * 2};
    const uint32_t box[4] = {kKC, kHaloW, kHaloH, 1};
    int st = ctk::encode_tmap_bf16_sw128(&tm_a, x_bf16, 4, dims, strides, box);
    if (st != CTK_OK) return st;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(p.cin), static_cast<uint64_t>(9) * p.cout};
    const uint64_t strides[1] = {static_cast<uint64_t>(p.cin) * 2};
    const uint32_t box[2] = {kKC, static_cast<uint32_t>(C::kBRows)};
    int st = ctk::encode_tmap_bf16_sw128(&tm_b, w_packed_bf16, 2, dims, strides, box);
    if (st != CTK_OK) return st;
  }
  CUtensorMap tm_out = tm_a;          // only read by the staged raw epilogues
  if constexpr (C::kStage) {
    const uint64_t cs = static_cast<uint64_t>(p.out_cstride);
    const uint64_t dims[4] = {cs, static_cast<uint64_t>(p.W), static_cast<uint64_t>(p.H), static_cast<uint64_t>(p.n_img)};
    const uint64_t strides[3] = {cs * 2, static_cast<uint64_t>(p.W) * cs * 2, static_cast<uint64_t>(p.H) * p.W * cs * 2};
    const uint32_t box[4] = {64, kTileW, 4, 1};
    int st = ctk::encode_tmap_bf16_sw128(&tm_out, p.out, 4, dims, strides, box);
    if (st != CTK_OK) return st;
  }
  auto kernel = conv3x3_tc_kernel<kCtaGroup, kBlockN, kEpi>;
  CTK_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
  int clusters = std::min(p.total_work, ctk::persistent_sms() / kCtaGroup);
  // statistics: every CTA must keep one N tile (register-resident totals), i.e. the work stride is a multiple of tiles_n
  if (kEpi == kEpiRawStats && clusters > p.tiles_n) clusters -= clusters % p.tiles_n;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * kCtaGroup);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtaGroup;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CTK_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, tm_a, tm_b, tm_out, p));
  int st = ctk::check_launch();
  if (st != CTK_OK || kEpi != kEpiRawStats) return st;
  return ctk::reduce_rows_f32(p.stats, clusters * kCtaGroup, 2 * p.cout, 2 * p.cout, p.stats_out, stream);
}

}  // namespace

// one row of 2 * cout partial sums per CTA of the persistent grid (<= one CTA per SM)
extern "C" size_t ctk_conv3x3_tc_raw_workspace_bytes(int cout) {
  return cout > 0 ? static_cast<size_t>(ctk::num_sms()) * 2 * cout * sizeof(float) : 0;
}

static int conv_dispatch(const void* x_bf16, int n, int H, int W, int cin, const void* w_packed_bf16, int cout,
                         const float* scale, const float* shift, float* stats, float slope, void* out_bf16,
                         int out_cstride, int out_coffset, int flags, void* stream, void* out_lo_bf16 = nullptr,
                         void* workspace = nullptr, size_t workspace_bytes = 0) {
  if (n == 0) return CTK_OK;
  CTK_REQUIRE(x_bf16 && w_packed_bf16 && out_bf16 && (scale == nullptr) == (shift == nullptr));
  CTK_REQUIRE(n > 0 && H > 0 && W > 0 && H % 2 == 0 && W % kTileW == 0 && cin > 0 && cin % kKC == 0 && cout > 0 &&
              cout % 64 == 0 && cout <= 512);
  CTK_REQUIRE(out_coffset >= 0 && out_coffset + cout <= out_cstride && out_cstride % 8 == 0 && out_coffset % 8 == 0);
  CTK_REQUIRE((reinterpret_cast<uintptr_t>(x_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed_bf16) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0);

  ConvParams p = {};
  p.n_img = n; p.H = H; p.W = W; p.cin = cin; p.cout = cout;
  p.tiles_x = W / kTileW;
  p.tiles_y = (H + kTileH - 1) / kTileH;
  const long long spatial = static_cast<long long>(n) * p.tiles_x * p.tiles_y;
  CTK_REQUIRE(spatial < (1ll << 30));
  p.spatial_tiles = static_cast<int>(spatial);
  p.pool = (flags & CTK_CONV_NO_POOL) ? 0 : 1;
  p.act = (flags & CTK_CONV_NO_ACT) ? 0 : 1;
  p.slope = slope;
  p.scale = scale; p.shift = shift;
  p.stats = nullptr; p.stats_out = stats;
  if (stats != nullptr) {
    CTK_REQUIRE_WORKSPACE(workspace, workspace_bytes, ctk_conv3x3_tc_raw_workspace_bytes(cout));
    p.stats = static_cast<float*>(workspace);
  }
  p.out = static_cast<__nv_bfloat16*>(out_bf16);
  p.out_lo = static_cast<__nv_bfloat16*>(out_lo_bf16);
  p.out_cstride = out_cstride; p.out_coffset = out_coffset;
  p.cin_phys = cin; p.a_wrap = 0;
  cudaStream_t s = ctk::as_stream(stream);
  if (out_lo_bf16 != nullptr) {
    // fp32-class: the activation tensor holds [hi | lo] (2/3 of the logical K), K runs over [hi | lo | hi]
    CTK_REQUIRE(cin % 3 == 0 && (cin / 3) % kKC == 0 && scale != nullptr && p.pool && p.act);
    p.cin_phys = cin / 3 * 2;
    p.a_wrap = p.cin_phys / kKC;
    if (cout % 128 == 0) return launch_conv<2, 128, kEpiEvalPoolSplit>(x_bf16, w_packed_bf16, p, s);
    return CTK_ERR_UNSUPPORTED;
  }
  const int epi = scale == nullptr ? (stats != nullptr ? kEpiRawStats : kEpiRaw)
                                   : (p.pool && p.act ? kEpiEvalPool : kEpiEvalAny);
#define CTK_CONV_LAUNCH(G, N)                                                                             \
  switch (epi) {                                                                                          \
    case kEpiEvalPool: return launch_conv<G, N, kEpiEvalPool>(x_bf16, w_packed_bf16, p, s);               \
    case kEpiEvalAny: return launch_conv<G, N, kEpiEvalAny>(x_bf16, w_packed_bf16, p, s);                 \
    case kEpiRaw: return launch_conv<G, N, kEpiRaw>(x_bf16, w_packed_bf16, p, s);                         \
    default: return launch_conv<G, N, kEpiRawStats>(x_bf16, w_packed_bf16, p, s);                         \
  }
  if ((flags & CTK_CONV_SINGLE_CTA) && cout % 128 == 0) { CTK_CONV_LAUNCH(1, 128) }
  if (cout % 256 == 0) { CTK_CONV_LAUNCH(2, 256) }
  if (cout % 128 == 0) { CTK_CONV_LAUNCH(2, 128) }
  CTK_CONV_LAUNCH(2, 64)
#undef CTK_CONV_LAUNCH
}

extern "C" int ctk_conv3x3_tc_eval(const void* x_bf16, int n, int H, int W, int cin, const void* w_packed_bf16,
                                   int cout, const float* scale, const float* shift, float slope, void* out_bf16,
                                   int out_cstride, int out_coffset, int flags, void* stream) {
  CTK_REQUIRE(scale && shift);
  return conv_dispatch(x_bf16, n, H, W, cin, w_packed_bf16, cout, scale, shift, nullptr, slope, out_bf16, out_cstride,
                       out_coffset, flags, stream);
}

extern "C" int ctk_conv3x3_tc_eval_split(const void* x_split_bf16, int n, int H, int W, int cin,
                                         const void* w_split_bf16, int cout, const float* scale, const float* shift,
                                         float slope, void* out_hi_bf16, void* out_lo_bf16, int out_cstride,
                                         int out_coffset, void* stream) {
  CTK_REQUIRE(scale && shift && out_lo_bf16 && (reinterpret_cast<uintptr_t>(out_lo_bf16) & 15) == 0);
  return conv_dispatch(x_split_bf16, n, H, W, 3 * cin, w_split_bf16, cout, scale, shift, nullptr, slope, out_hi_bf16,
                       out_cstride, out_coffset, 0, stream, out_lo_bf16);
}

extern "C" int ctk_conv3x3_tc_raw(const void* x_bf16, int n, int H, int W, int cin, const void* w_packed_bf16,
                                  int cout, void* y_bf16, float* stats, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  return conv_dispatch(x_bf16, n, H, W, cin, w_packed_bf16, cout, nullptr, nullptr, stats, 0.f, y_bf16, cout, 0,
                       CTK_CONV_NO_POOL | CTK_CONV_NO_ACT, stream, nullptr, workspace, workspace_bytes);
}
