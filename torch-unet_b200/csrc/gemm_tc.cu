// Split-K bf16 GEMM on tcgen05 for the first fully connected layer of each regression head
// (nn.Linear at /root/reference/regression_model.py:36 and two_branch_regression.py:42):
//   partial[s][m][n] = sum_{k in split s} A[m][k] * B[n][k]
// A = flattened NHWC activations [batch, K], B = NHWC-column-permuted FC1 weight [512, K]; both K-major, so
// both operands are plain 128B-swizzled TMA tiles.  The batch is tiny (M = 256) and K is huge (262 144 for
// the double-branch model), hence split-K: (M/128) x (N/128) x splits CTAs, each streaming its K range
// through a 6-stage TMA->UMMA pipeline and writing one fp32 partial tile; ctk_head_eval sums the splits.
#include "ctk_common.h"
#include "ctk_ptx.cuh"

namespace {

using namespace ctk;

constexpr int kBM = 128, kBN = 128, kBK = 64;
constexpr int kThreads = 192;                         // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 epilogue
constexpr int kStageBytes = (kBM + kBN) * kBK * 2;    // 32 KiB
// kStages = 6: long K ranges (FC1 forward), one CTA per SM.  kStages = 2: the FC1 backward GEMMs have K = 256 / 512 and
// thousands of output tiles, so a CTA is mostly prologue + epilogue; 66 KB of shared memory lets three of them share an SM
// and overlap each other's TMEM allocation, pipeline fill and 64 KB fp32 tile store.
template <int kStages>
constexpr int smem_bytes() { return 1024 + kStages * kStageBytes + 256; }

template <int kStages>
struct GemmSmem {
  uint64_t full[kStages], empty[kStages], acc_full;
  uint32_t tmem_base;
};

// kBMn: B is given as [K][N] (N contiguous) instead of [N][K]: consumed as an MN-major 128B-swizzled operand -- two TMA
// boxes of 64 N-elements x 64 K-rows per stage (LBO = 8 KiB between them, SBO = 1 KiB between 8-row K atoms), 2 KiB per
// UMMA_K step.  This is how FC1's dX GEMM reads the forward pass's packed weight without a transposed copy.
__host__ __device__ constexpr uint32_t gemm_idesc(bool b_mn) {
  return umma_idesc_bf16_f32(kBM, kBN) | (b_mn ? (1u << 16) : 0u);
}
__device__ __forceinline__ uint64_t gemm_desc_mn_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3fff) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int kStages, bool kBMn>
__global__ void __launch_bounds__(kThreads, kStages <= 2 ? 3 : 1)
gemm_splitk_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                   const __grid_constant__ CUtensorMap tm_out, int M, int kb_base, int kb_extra, int out_is_bf16) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GemmSmem<kStages>* sl = reinterpret_cast<GemmSmem<kStages>*>(smem + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * kBN, split = blockIdx.z;
  // K blocks (64 elements) are dealt out as evenly as they go: the first kb_extra splits take one more than the others, so
  // the split count can follow the SM count (18 x 8 tiles = 144 CTAs for FC1) instead of the divisors of K / 64
  const int k_begin = (split * kb_base + (split < kb_extra ? split : kb_extra)) * kBK;
  const int k_iters = kb_base + (split < kb_extra ? 1 : 0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&sl->full[i], 1); mbar_init(&sl->empty[i], 1); }
    mbar_init(&sl->acc_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_out);
  }
  if (warp == 1) tmem_alloc(&sl->tmem_base, kBN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sl->tmem_base;

  if (warp == 0 && lane == 0) {
    int stage = 0, phase = 0;
    for (int i = 0; i < k_iters; ++i) {
      mbar_wait(&sl->empty[stage], phase ^ 1);
      mbar_arrive_expect_tx(&sl->full[stage], kStageBytes);
      uint8_t* a_dst = smem + stage * kStageBytes;
      tma_load_2d(a_dst, &tm_a, &sl->full[stage], k_begin + i * kBK, m0);
      if constexpr (kBMn) {
        tma_load_2d(a_dst + kBM * kBK * 2, &tm_b, &sl->full[stage], n0, k_begin + i * kBK);
        tma_load_2d(a_dst + kBM * kBK * 2 + 64 * kBK * 2, &tm_b, &sl->full[stage], n0 + 64, k_begin + i * kBK);
      } else {
        tma_load_2d(a_dst + kBM * kBK * 2, &tm_b, &sl->full[stage], k_begin + i * kBK, n0);
      }
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = gemm_idesc(kBMn);
    int stage = 0, phase = 0;
    for (int i = 0; i < k_iters; ++i) {
      mbar_wait(&sl->full[stage], phase);
      tc_fence_after();
      const uint32_t a_base = smem_u32(smem + stage * kStageBytes);
      const uint32_t b_base = a_base + kBM * kBK * 2;
#pragma unroll
      for (int s = 0; s < kBK / 16; ++s) {
        const uint64_t bdesc = kBMn ? gemm_desc_mn_sw128(b_base + s * 2048, 64 * kBK * 2, 1024)
                                    : umma_smem_desc_sw128(b_base + s * 32, 1024, 0);
        umma_bf16(tmem_base, umma_smem_desc_sw128(a_base + s * 32, 1024, 0), bdesc, idesc, (i | s) != 0 ? 1u : 0u);
      }
      umma_commit(&sl->empty[stage]);
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    umma_commit(&sl->acc_full);
  } else if (warp >= 2) {
    // The accumulator barrier fires after the last MMA has completed, i.e. after every pipeline stage has been consumed:
    // the stage buffers are free and become the output staging area.  Each warp owns 16 KiB of it = its 32 rows x 128
    // columns as 128-byte rows in TMA's 128B-swizzled box layout (4 boxes of 32 fp32 columns, or 2 boxes of 64 bf16
    // columns), written with conflict-free 16-byte shared stores and drained by TMA bulk tensor stores: full 128-byte
    // lines to HBM instead of 32 scattered 16-byte pieces per store instruction.
    const int q = warp & 3;                       // TMEM lane quadrant this warp may read
    const int row0 = split * M + m0 + q * 32;     // first output row of this warp (split-major partials)
    mbar_wait(&sl->acc_full, 0);
    tc_fence_after();
    const uint32_t stage_base = smem_u32(smem) + static_cast<uint32_t>(q) * 16384u;
    const uint32_t row_off = static_cast<uint32_t>(lane) * 128u;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
#pragma unroll 1
    for (int cb = 0; cb < kBN / 32; ++cb) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + cb * 32, v);
      tmem_ld_wait();
      if (!out_is_bf16) {
        const uint32_t box = stage_base + static_cast<uint32_t>(cb) * 4096u;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(box + row_off + ((static_cast<uint32_t>(i) ^ sw) << 4)),
                       "r"(v[4 * i]), "r"(v[4 * i + 1]), "r"(v[4 * i + 2]), "r"(v[4 * i + 3])
                       : "memory");
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tm_out)),
                       "r"(box), "r"(n0 + cb * 32), "r"(row0)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {
        const uint32_t box = stage_base + static_cast<uint32_t>(cb >> 1) * 4096u;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(
                           box + row_off + ((static_cast<uint32_t>((cb & 1) * 4 + i) ^ sw) << 4)),
                       "r"(pack_bf16x2(__uint_as_float(v[8 * i]), __uint_as_float(v[8 * i + 1]))),
                       "r"(pack_bf16x2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3]))),
                       "r"(pack_bf16x2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5]))),
                       "r"(pack_bf16x2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7])))
                       : "memory");
        if (cb & 1) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tm_out)),
                         "r"(box), "r"(n0 + (cb >> 1) * 64), "r"(row0)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores must land before the CTA exits
    __syncwarp();
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kBN);
  }
}

}  // namespace

template <int kStages, bool kBMn>
static int gemm_run(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const CUtensorMap& tm_out, dim3 grid, int M,
                    int k_blocks, bool out_is_bf16, void* stream) {
  auto kernel = gemm_splitk_kernel<kStages, kBMn>;
  CTK_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<kStages>()));
  const int splits = static_cast<int>(grid.z);
  kernel<<<grid, kThreads, smem_bytes<kStages>(), ctk::as_stream(stream)>>>(tm_a, tm_b, tm_out, M, k_blocks / splits,
                                                                           k_blocks % splits, out_is_bf16 ? 1 : 0);
  return ctk::check_launch();
}

static int gemm_launch(const void* a_bf16, const void* b_bf16, int M, int N, int K, int splits, float* partial,
                       void* out_bf16, void* stream, bool b_mn = false) {
  CTK_REQUIRE(a_bf16 && b_bf16 && (partial != nullptr) != (out_bf16 != nullptr) && M > 0 && N > 0 && K > 0 && splits > 0);
  CTK_REQUIRE(M % kBM == 0 && N % kBN == 0 && K % kBK == 0 && splits <= K / kBK && splits <= 65535 && N / kBN <= 65535);
  CTK_REQUIRE(out_bf16 == nullptr || splits == 1);
  CTK_REQUIRE((reinterpret_cast<uintptr_t>(a_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(b_bf16) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(partial) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0);
  CUtensorMap tm_a, tm_b;
  const uint32_t box[2] = {kBK, kBM};
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    int st = ctk::encode_tmap_bf16_sw128(&tm_a, a_bf16, 2, dims, strides, box);
    if (st != CTK_OK) return st;
  }
  if (b_mn) {
    const uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(K)};
    const uint64_t strides[1] = {static_cast<uint64_t>(N) * 2};
    const uint32_t box_mn[2] = {64, kBK};
    int st = ctk::encode_tmap_bf16_sw128(&tm_b, b_bf16, 2, dims, strides, box_mn);
    if (st != CTK_OK) return st;
  } else {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    const uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    int st = ctk::encode_tmap_bf16_sw128(&tm_b, b_bf16, 2, dims, strides, box);
    if (st != CTK_OK) return st;
  }
  // output: [splits * M rows][N] fp32 partials in boxes of 32 rows x 32 columns, or [M][N] bf16 in boxes of 32 x 64
  CUtensorMap tm_out;
  {
    const bool bf = out_bf16 != nullptr;
    const uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M) * (bf ? 1 : splits)};
    const uint64_t strides[1] = {static_cast<uint64_t>(N) * (bf ? 2 : 4)};
    const uint32_t box_out[2] = {bf ? 64u : 32u, 32u};
    int st = bf ? ctk::encode_tmap_bf16_sw128(&tm_out, out_bf16, 2, dims, strides, box_out)
                : ctk::encode_tmap_f32_sw128(&tm_out, partial, 2, dims, strides, box_out);
    if (st != CTK_OK) return st;
  }
  dim3 grid(M / kBM, N / kBN, splits);
  const bool short_k = K / kBK / splits <= 8;
  const bool bf = out_bf16 != nullptr;
  if (b_mn) return short_k ? gemm_run<2, true>(tm_a, tm_b, tm_out, grid, M, K / kBK, bf, stream)
                           : gemm_run<6, true>(tm_a, tm_b, tm_out, grid, M, K / kBK, bf, stream);
  return short_k ? gemm_run<2, false>(tm_a, tm_b, tm_out, grid, M, K / kBK, bf, stream)
                 : gemm_run<6, false>(tm_a, tm_b, tm_out, grid, M, K / kBK, bf, stream);
}

extern "C" int ctk_gemm_bf16_splitk(const void* a_bf16, const void* b_bf16, int M, int N, int K, int splits,
                                    float* partial, void* stream) {
  return gemm_launch(a_bf16, b_bf16, M, N, K, splits, partial, nullptr, stream);
}

extern "C" int ctk_gemm_bf16_out_bf16(const void* a_bf16, const void* b_bf16, int M, int N, int K, void* c_bf16,
                                      void* stream) {
  return gemm_launch(a_bf16, b_bf16, M, N, K, 1, nullptr, c_bf16, stream);
}

/* C[M,N] (bf16) = A[M,K] * B[K,N]: B row-major with N contiguous (an MN-major tensor-core operand, no transposed copy). */
extern "C" int ctk_gemm_bf16_bt_out_bf16(const void* a_bf16, const void* b_kn_bf16, int M, int N, int K, void* c_bf16,
                                         void* stream) {
  return gemm_launch(a_bf16, b_kn_bf16, M, N, K, 1, nullptr, c_bf16, stream, true);
}
