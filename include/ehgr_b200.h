/*
 * ehgr_b200.h — C ABI of libehgr_b200.so, the B200 (sm_100a) implementation of the
 * TSM/ACTION-MobileNetV2 training hot path.
 *
 * The reference (peter0512lee/Efficient-Hand-Gesture-Recognition-...) is pure PyTorch: it has no
 * FFI of its own, so each entry point below names the reference Python operator it replaces
 * (file:line under /root/reference).  The host-side mirror of those operators (same class names,
 * arguments and error behaviour) lives in the package next to csrc/ and binds these symbols with
 * ctypes — see INTEGRATION.md.
 *
 * Conventions (all entry points):
 *   - plain pointers and ints only; pointers are DEVICE pointers unless stated otherwise;
 *   - return 0 on success, a negative EHGR_E_* code for argument errors, a positive value is a
 *     cudaError_t from the launch; nothing throws, nothing allocates, nothing synchronises;
 *   - work is enqueued on `stream` only (a cudaStream_t passed as void*; NULL = legacy stream);
 *   - dtype: EHGR_F32 / EHGR_BF16 is the STORAGE type of activations; arithmetic is fp32;
 *   - layout: EHGR_NCHW (reference layout) or EHGR_NHWC (channels-last, the layout the fused
 *     block kernels use internally);
 *   - re-entrant and thread-safe: no mutable global state.
 */
#ifndef EHGR_B200_H_
#define EHGR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EHGR_ABI_VERSION 1

enum { EHGR_F32 = 0, EHGR_BF16 = 1 };
enum { EHGR_NCHW = 0, EHGR_NHWC = 1 };

enum {
  EHGR_OK = 0,
  EHGR_E_NULL = -1,      /* null pointer */
  EHGR_E_ALIGN = -2,     /* pointer not aligned to the element size */
  EHGR_E_DTYPE = -3,     /* unsupported dtype / layout enum */
  EHGR_E_SHAPE = -4,     /* non-positive or inconsistent shape */
  EHGR_E_UNSUPPORTED = -5
};

typedef void* ehgr_stream_t; /* cudaStream_t */

int ehgr_abi_version(void);
/* Human-readable text for a return code of this library (static storage). */
const char* ehgr_status_string(int status);
/* Number of kernel launches this library has enqueued in this process (monotonic, atomic). */
long long ehgr_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * K1  temporal shift — replaces TemporalShift.shift (models/temporal_shift.py:27-46, duplicate
 *     models/action.py:135-154) and its autograd backward; equals InplaceShift fwd/bwd
 *     (models/temporal_shift.py:49-76).
 *
 *   x, out : [n_batch*n_segment, c, h, w] (NCHW) or [n_batch*n_segment, h, w, c] (NHWC),
 *            contiguous, `hw` = h*w.  fold = c / fold_div (computed by the caller exactly as the
 *            reference does, models/temporal_shift.py:33).
 *   fwd : out[:, t, :fold]       = x[:, t+1, :fold]        (t <  T-1, else 0)
 *         out[:, t, fold:2fold]  = x[:, t-1, fold:2fold]   (t >= 1,   else 0)
 *         out[:, t, 2fold:]      = x[:, t,   2fold:]
 *   bwd : the adjoint (the mirror shift) applied to grad_out.
 *   Pure copy: bit-exact in every dtype.  x and out must not alias.
 * ------------------------------------------------------------------------------------------- */
int ehgr_temporal_shift_fwd(const void* x, void* out, int n_batch, int n_segment, int c, int hw,
                            int fold, int dtype, int layout, ehgr_stream_t stream);
int ehgr_temporal_shift_bwd(const void* grad_out, void* grad_in, int n_batch, int n_segment, int c,
                            int hw, int fold, int dtype, int layout, ehgr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EHGR_B200_H_ */
