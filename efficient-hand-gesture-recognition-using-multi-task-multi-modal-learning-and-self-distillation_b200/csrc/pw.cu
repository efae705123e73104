// pw.cu — C-ABI entry points of the pointwise-conv GEMM family; picks the engine.
#include "rowop.cuh"
#include "bnfin.cuh"

namespace ehgr {
int pw_gemm_simt(const RowOp& a, const float* w, int w_is_kn, void* out, const void* addend, double* stats,
                 long long M, int K, int N, int dtype, cudaStream_t s);
int pw_wgrad_simt(const RowOp& dy, const RowOp& a, float* dw, long long M, int K, int N, int dtype,
                  cudaStream_t s);
// tensor-core engine (pw_tc.cu); *_supported() say whether a shape is taken
bool pw_gemm_tc_supported(const RowOp& a, int w_is_kn, long long M, int K, int N, int dtype);
int pw_gemm_tc(const RowOp& a, const float* w, const void* w16, int w_is_kn, void* out, const void* addend,
               double* stats, long long M, int K, int N, cudaStream_t s);
bool pw_wgrad_tc_supported(const RowOp& dy, const RowOp& a, long long M, int K, int N, int dtype);
int pw_wgrad_tc(const RowOp& dy, const RowOp& a, float* dw, long long M, int K, int N, cudaStream_t s);
}  // namespace ehgr

using namespace ehgr;

extern "C" int ehgr_pw_gemm(const ehgr_rowop* a, const float* w, int w_is_kn, void* out, const void* addend,
                            double* stats, long long M, int K, int N, int dtype, int engine,
                            ehgr_stream_t stream) {
  return ehgr_pw_gemm_w16(a, w, nullptr, w_is_kn, out, addend, stats, M, K, N, dtype, engine, stream);
}

extern "C" int ehgr_pw_gemm_w16(const ehgr_rowop* a, const float* w, const void* w16, int w_is_kn, void* out,
                                const void* addend, double* stats, long long M, int K, int N, int dtype, int engine,
                                ehgr_stream_t stream) {
  return ehgr_pw_gemm_bn(a, w, w16, w_is_kn, out, addend, stats, M, K, N, dtype, engine, nullptr, stream);
}

static int check_bnfin(const ehgr_bnfin* fin, const double* stats) {
  if (!fin) return EHGR_OK;
  if (!fin->scale || !fin->shift || !fin->counter) return EHGR_E_NULL;
  if (fin->training && (!stats || fin->count <= 0)) return EHGR_E_NULL;
  if (!fin->training && (!fin->running_mean || !fin->running_var)) return EHGR_E_NULL;
  return EHGR_OK;
}

extern "C" int ehgr_pw_gemm_bn(const ehgr_rowop* a, const float* w, const void* w16, int w_is_kn, void* out,
                               const void* addend, double* stats, long long M, int K, int N, int dtype, int engine,
                               const ehgr_bnfin* fin, ehgr_stream_t stream) {
  const int es = esize_of(dtype);
  if (int st = check_bnfin(fin, stats)) return st;
  if (es == 0) return EHGR_E_DTYPE;
  if (!w || !out) return EHGR_E_NULL;
  if (int st = validate_rowop(a, es, true)) return st;
  if (M < 0 || K <= 0 || N <= 0 || (K % 8) || (N % 8)) return EHGR_E_SHAPE;
  if (a->mode == EHGR_ROW_CONV3 && (K != 9 * a->cv_cin || M % a->hw || M * a->cv_cin > 0x7fffffffLL)) return EHGR_E_SHAPE;   // 32-bit element offsets
  if (!aligned_to(out, 16) || !aligned_to(w, 16) || (w16 && !aligned_to(w16, 16)) || (addend && !aligned_to(addend, 16)) ||
      (stats && !aligned_to(stats, 8)))
    return EHGR_E_ALIGN;
  if (M == 0) return EHGR_OK;
  cudaStream_t s = as_stream(stream);
  const bool tc_ok = pw_gemm_tc_supported(*a, w_is_kn, M, K, N, dtype);
  if (engine == EHGR_ENGINE_TCGEN05 && !tc_ok) return EHGR_E_UNSUPPORTED;
  fin_slot().fin = fin;            // the tensor-core kernel finalises in its last CTA; any other engine leaves it parked
  const int st = (engine != EHGR_ENGINE_SIMT && tc_ok) ? pw_gemm_tc(*a, w, w16, w_is_kn, out, addend, stats, M, K, N, s)
                                                       : pw_gemm_simt(*a, w, w_is_kn, out, addend, stats, M, K, N, dtype, s);
  return finish_fin(stats, N, s, st);
}

extern "C" int ehgr_pw_wgrad(const ehgr_rowop* dy, const ehgr_rowop* a, float* dw, long long M, int K, int N,
                             int dtype, int engine, ehgr_stream_t stream) {
  const int es = esize_of(dtype);
  if (es == 0) return EHGR_E_DTYPE;
  if (!dw) return EHGR_E_NULL;
  if (int st = validate_rowop(dy, es)) return st;
  if (int st = validate_rowop(a, es, true)) return st;
  if (M < 0 || K <= 0 || N <= 0 || (K % 8) || (N % 8)) return EHGR_E_SHAPE;
  if (a->mode == EHGR_ROW_CONV3 && (K != 9 * a->cv_cin || M % a->hw || M * a->cv_cin > 0x7fffffffLL)) return EHGR_E_SHAPE;   // 32-bit element offsets
  if (!aligned_to(dw, 16)) return EHGR_E_ALIGN;
  if (M == 0) return EHGR_OK;
  cudaStream_t s = as_stream(stream);
  const bool tc_ok = pw_wgrad_tc_supported(*dy, *a, M, K, N, dtype);
  if (engine == EHGR_ENGINE_TCGEN05 && !tc_ok) return EHGR_E_UNSUPPORTED;
  if (engine != EHGR_ENGINE_SIMT && tc_ok) return pw_wgrad_tc(*dy, *a, dw, M, K, N, s);
  return pw_wgrad_simt(*dy, *a, dw, M, K, N, dtype, s);
}
