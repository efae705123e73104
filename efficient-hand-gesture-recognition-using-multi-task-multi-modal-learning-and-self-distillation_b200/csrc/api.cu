// api.cu — library-level entry points of libehgr_b200.so (version, status text, launch counter).
#include "common.cuh"

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

namespace ehgr { int g_debug_flags = 0; }
// undocumented bring-up switch (timing experiments only; results are wrong when set)
extern "C" void ehgr_debug_set(int flags) { ehgr::g_debug_flags = flags; }

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
