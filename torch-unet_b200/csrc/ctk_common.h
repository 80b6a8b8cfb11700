// Host-side helpers shared by the libctk translation units: status plumbing, launch checks and
// TMA tensor-map encoding through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/ctk.h"

namespace ctk {

void set_last_cuda_error(int e);

#define CTK_CUDA_TRY(expr)                                 \
  do {                                                     \
    cudaError_t e__ = (expr);                              \
    if (e__ != cudaSuccess) {                              \
      ::ctk::set_last_cuda_error(static_cast<int>(e__));   \
      return CTK_ERR_CUDA;                                 \
    }                                                      \
  } while (0)

#define CTK_REQUIRE(cond)              \
  do {                                 \
    if (!(cond)) return CTK_ERR_BAD_ARG; \
  } while (0)

inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_cuda_error(static_cast<int>(e));
    return CTK_ERR_CUDA;
  }
  return CTK_OK;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// rank <= 4 bf16 tensor map, 128-byte swizzle.  dims/strides innermost first; strides in bytes for dims 1..rank-1.
int encode_tmap_bf16_sw128(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box);
// same for fp32 elements (box[0] <= 32)
int encode_tmap_f32_sw128(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box);

int num_sms();
// SMs the one-CTA-per-SM persistent tensor-core kernels may fill: num_sms() minus the reserve set through
// ctk_set_persistent_sm_reserve (SMs left free for a concurrently running collective kernel)
int persistent_sms();

// Second stage of the cross-CTA reductions (reduce.cu): out[c] = sum_{r < rows} part[r * stride + c], rows added in a
// fixed order with fp64 accumulation, so the result is independent of CTA scheduling.
int reduce_rows_f32(const float* part, int rows, long long stride, int cols, float* out, cudaStream_t s);
int reduce_rows_f64(const double* part, int rows, long long stride, int cols, double* out, cudaStream_t s);

#define CTK_REQUIRE_WORKSPACE(ptr, have, need)                                         \
  do {                                                                                 \
    if ((ptr) == nullptr || (have) < (need)) return CTK_ERR_WORKSPACE;                 \
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return CTK_ERR_BAD_ARG;          \
  } while (0)

// n / d for 0 <= n < 2^31 by multiply-high (Granlund-Montgomery, 31-bit dividend): the role loops decode a work index
// per tile, and a run-time integer division costs ~45 instructions each in every one of the 11 warps
struct FastDiv {
  uint32_t d, mul, shr;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f = {d, 0u, 0u};
  if (d > 1) {
    uint32_t l = 0;
    while ((1u << l) < d) ++l;
    f.mul = static_cast<uint32_t>(((1ull << (31 + l)) + d - 1) / d);
    f.shr = l - 1;
  }
  return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shr);
}
#endif

}  // namespace ctk
