// pw_tc.cu — tcgen05/TMEM tensor-core engine of the pointwise-conv GEMM family (placeholder: shapes
// are reported unsupported until the kernel lands; pw.cu then uses the fp32 SIMT engine).
#include "rowop.cuh"

namespace ehgr {
bool pw_gemm_tc_supported(const RowOp&, int, long long, int, int, int) { return false; }
int pw_gemm_tc(const RowOp&, const float*, int, void*, const void*, double*, long long, int, int, cudaStream_t) {
  return EHGR_E_UNSUPPORTED;
}
bool pw_wgrad_tc_supported(const RowOp&, const RowOp&, long long, int, int, int) { return false; }
int pw_wgrad_tc(const RowOp&, const RowOp&, float*, long long, int, int, cudaStream_t) { return EHGR_E_UNSUPPORTED; }
}  // namespace ehgr
