// Status strings, device check and the TMA tensor-map encoder used by the tcgen05 kernels.
#include "ctk_common.h"

#include <cudaTypedefs.h>
#include <atomic>
#include <mutex>

namespace ctk {

static thread_local int g_last_cuda_error = 0;
void set_last_cuda_error(int e) { g_last_cuda_error = e; }

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
  });
  return fn;
}

static int encode_tmap_sw128(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box) {
  auto fn = get_encode_fn();
  if (!fn) return CTK_ERR_NO_DEVICE;
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(map, dtype, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                  gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_cuda_error(static_cast<int>(r));
    return CTK_ERR_CUDA;
  }
  return CTK_OK;
}

int encode_tmap_bf16_sw128(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box) {
  return encode_tmap_sw128(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box);
}

int encode_tmap_f32_sw128(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box) {
  return encode_tmap_sw128(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

static std::atomic<int> g_sm_reserve{0};
int persistent_sms() {
  const int n = num_sms() - g_sm_reserve.load(std::memory_order_relaxed);
  return n < 8 ? 8 : n;
}

}  // namespace ctk

extern "C" {

int ctk_set_persistent_sm_reserve(int sms) {
  if (sms < 0 || sms > 64) return CTK_ERR_BAD_ARG;
  ctk::g_sm_reserve.store(sms, std::memory_order_relaxed);
  return CTK_OK;
}

int ctk_abi_version(void) { return 2; }

const char* ctk_status_string(int status) {
  switch (status) {
    case CTK_OK: return "ok";
    case CTK_ERR_BAD_ARG: return "bad argument (null pointer, unsupported shape or misaligned buffer)";
    case CTK_ERR_WORKSPACE: return "workspace too small";
    case CTK_ERR_CUDA: return "CUDA call failed (see ctk_last_cuda_error)";
    case CTK_ERR_NO_DEVICE: return "no usable sm_100 CUDA device";
    case CTK_ERR_UNSUPPORTED: return "not implemented in this build";
    default: return "unknown status";
  }
}

int ctk_last_cuda_error(void) { return ctk::g_last_cuda_error; }

int ctk_device_check(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return CTK_ERR_NO_DEVICE;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return CTK_ERR_NO_DEVICE;
  return major == 10 ? CTK_OK : CTK_ERR_NO_DEVICE;
}

}  // extern "C"
