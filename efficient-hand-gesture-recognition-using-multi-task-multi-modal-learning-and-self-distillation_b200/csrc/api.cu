// api.cu — library-level entry points of libehgr_b200.so (version, status text, launch counter).
#include "bnfin.cuh"

#include <cuda.h>

#include <mutex>
#include <unordered_map>

namespace ehgr {
std::atomic<long long> g_launches{0};

void ensure_dyn_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::unordered_map<const void*, int> have;
  std::lock_guard<std::mutex> lock(mu);
  int& cur = have[func];
  if (cur >= bytes) return;
  if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess) cur = bytes;
}
}

namespace ehgr {
FinSlot& fin_slot() {
  static thread_local FinSlot slot;
  return slot;
}
}  // namespace ehgr

namespace ehgr {
namespace tma {
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int make_nhwc_bf16_map(CUtensorMap* out, const void* base, int nt, int h, int w, int c, int box_c, int box_w, int box_h) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return EHGR_E_UNSUPPORTED;
  const cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h),
                              static_cast<cuuint64_t>(nt)};
  const cuuint64_t strides[3] = {static_cast<cuuint64_t>(c) * 2, static_cast<cuuint64_t>(w) * c * 2,
                                 static_cast<cuuint64_t>(h) * w * c * 2};
  const cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? EHGR_OK : EHGR_E_UNSUPPORTED;
}
int make_map_4d(CUtensorMap* out, int elem_bytes, const void* base, const unsigned long long (&dims)[4],
                const unsigned long long (&strides_bytes)[3], const unsigned (&box)[4], int swizzle_bytes) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return EHGR_E_UNSUPPORTED;
  cuuint64_t d[4], st[3];
  cuuint32_t b[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) { d[i] = dims[i]; b[i] = box[i]; }
  for (int i = 0; i < 3; ++i) st[i] = strides_bytes[i];
  const CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                        const_cast<void*>(base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? EHGR_OK : EHGR_E_UNSUPPORTED;
}
int make_map_2d_sw128(CUtensorMap* out, const void* base, unsigned long long cols, unsigned long long rows, unsigned box_rows) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return EHGR_E_UNSUPPORTED;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? EHGR_OK : EHGR_E_UNSUPPORTED;
}
}  // namespace tma
}  // namespace ehgr

extern "C" int ehgr_abi_version(void) { return EHGR_ABI_VERSION; }

extern "C" long long ehgr_launch_count(void) { return ehgr::g_launches.load(std::memory_order_relaxed); }

extern "C" const char* ehgr_status_string(int status) {
  switch (status) {
    case EHGR_OK: return "ok";
    case EHGR_E_NULL: return "null pointer argument";
    case EHGR_E_ALIGN: return "pointer not aligned for dtype/vector width";
    case EHGR_E_DTYPE: return "unsupported dtype or layout";
    case EHGR_E_SHAPE: return "invalid or inconsistent shape";
    case EHGR_E_UNSUPPORTED: return "unsupported configuration";
    default: break;
  }
  if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
  return "unknown ehgr status";
}
