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
 *   - re-entrant and thread-safe: the only mutable process-wide state is the launch counter (atomic) and two
 *     mutex-guarded caches (per-kernel dynamic-shared-memory attribute, driver entry point for tensor maps);
 *     there are no debug switches or other settings that change what a launch computes.
 */
#ifndef EHGR_B200_H_
#define EHGR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EHGR_ABI_VERSION 2

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

/* ---------------------------------------------------------------------------------------------
 * Row operand — how a fused kernel READS an NHWC activation [M, C] (M = NT*H*W rows).
 * Activations are stored raw (conv output before BatchNorm); BatchNorm(+ReLU6), the temporal shift
 * and the BatchNorm-backward combination are applied while loading (csrc/rowop.cuh).  HOST struct,
 * copied into the kernel's parameters; all pointers inside are device pointers.
 * ------------------------------------------------------------------------------------------- */
enum { EHGR_ROW_PLAIN = 0, EHGR_ROW_AFFINE = 1, EHGR_ROW_SHIFT = 2, EHGR_ROW_BNBWD = 3, EHGR_ROW_GATE = 4,
       EHGR_ROW_CONV3 = 5 };

typedef struct ehgr_rowop {
  int32_t mode;        /* EHGR_ROW_* */
  int32_t relu6;       /* activation code: 0 none, 1 ReLU6, 2 plain ReLU.  AFFINE: clamp to [0,hi];  BNBWD: mask in1 where
                          !(0 < in2*scale+shift < hi);  hi = 6 (ReLU6) or +inf (ReLU) */
  const void* in1;     /* [M, C] activation (or, for BNBWD, gradient w.r.t. the post-activation) */
  const void* in2;     /* BNBWD: raw forward output of the layer, [M, C] */
  const float* scale;  /* [C] BatchNorm scale  gamma*invstd       (AFFINE; BNBWD when relu6) */
  const float* shift;  /* [C] BatchNorm shift  beta - mean*scale  (AFFINE; BNBWD when relu6) */
  const float* ca;     /* [C] BNBWD: v = ca*mask*in1 + cb*in2 + cc */
  const float* cb;
  const float* cc;
  int32_t n_segment;   /* SHIFT: frames per clip T */
  int32_t fold;        /* SHIFT: shifted channels per direction */
  int32_t hw;          /* SHIFT: rows per frame (H*W) */
  int32_t shift_dir;   /* SHIFT: +1 forward shift, -1 its adjoint */
  /* GATE (ACTION, models/action.py:83,96,113,115): v = in1 * (3 + g1[m] + g2[m/hw, c] + g3[m/hw, c]) with
   * in2 = g1 (fp32 [M]), scale = g2, shift = g3 (fp32 [frames, C]) and hw = rows per frame. */
  /* CONV3 (pointwise-GEMM family only): the im2col row of a dense 3x3 convolution, pad 1, stride 1 — the operand that
   * turns ehgr_pw_gemm* / ehgr_pw_wgrad into the implicit GEMM of nn.Conv2d(cin, cout, 3, padding=1)
   * (models/models_MTMM.py:131-149).  Row m = (frame, ho, wo) of the OUTPUT grid [cv_h, cv_w] (hw = cv_h*cv_w), the
   * GEMM's K = 9*cv_cin and column k = tap*cv_cin + c with tap = 3*(dy+1) + (dx+1):
   *     v[m, k] = act(in1[frame, (ho+dy) >> cv_up, (wo+dx) >> cv_up, c] * scale[c] + shift[c]),  0 outside the grid
   * in1 is the stored NHWC tensor [frames, cv_h >> cv_up, cv_w >> cv_up, cv_cin]; cv_up = 1 folds the nearest-neighbour
   * nn.Upsample(scale_factor=2) that precedes the convolution into the gather (the upsampled tensor never exists);
   * scale == NULL: no affine / activation (plain tensor).  `relu6` is the activation code as for AFFINE. */
  int32_t cv_h, cv_w;  /* CONV3: output grid */
  int32_t cv_cin;      /* CONV3: channels of in1 (multiple of 8) */
  int32_t cv_up;       /* CONV3: 0 | 1 */
} ehgr_rowop;

enum { EHGR_ENGINE_AUTO = 0, EHGR_ENGINE_SIMT = 1, EHGR_ENGINE_TCGEN05 = 2 };

/* ---------------------------------------------------------------------------------------------
 * BatchNorm bookkeeping folded into the producing kernel (csrc/bnfin.cuh).  The *_bn entry points below take one
 * of these HOST structs (device pointers inside): the CTA that finishes LAST turns the batch statistics the launch
 * has just accumulated into what the consumers need, so no separate ehgr_bn_finalize / ehgr_bn_bwd_finalize launch
 * follows.  `counter` is a device uint32 that is ZERO before the launch (the kernel leaves it zero again).
 *   ehgr_bnfin (forward, nn.BatchNorm2d semantics as ehgr_bn_finalize): stats[2C] -> scale, shift, mean, invstd;
 *     training: running statistics updated with `momentum`; eval (training = 0): scale/shift from the running ones.
 *   ehgr_bnbwd (backward, as ehgr_bn_bwd_finalize): sums[2C] -> ca, cb, cc, d(gamma), d(beta).
 * A NULL struct pointer means "no finalisation" (the plain entry points).
 * ------------------------------------------------------------------------------------------- */
typedef struct ehgr_bnfin {
  const float* gamma;          /* [C] or NULL (= 1) */
  const float* beta;           /* [C] or NULL (= 0) */
  float* running_mean;         /* [C]; may be NULL in training mode (no running-statistics update) */
  float* running_var;
  float* scale;                /* out [C]: gamma * invstd */
  float* shift;                /* out [C]: beta - mean * scale */
  float* mean;                 /* out [C] or NULL */
  float* invstd;               /* out [C] or NULL */
  unsigned int* counter;
  long long count;             /* elements per channel (M) */
  float momentum, eps;
  int32_t training;
} ehgr_bnfin;

typedef struct ehgr_bnbwd {
  const float* gamma;          /* [C] or NULL (= 1) */
  const float* mean;           /* [C] saved by the forward finalisation */
  const float* invstd;
  float* ca;                   /* out [C]: coefficients of the BNBWD row operand */
  float* cb;
  float* cc;
  float* dgamma;               /* out [C] or NULL (stored, not accumulated) */
  float* dbeta;
  unsigned int* counter;
  long long count;
  int32_t training;
} ehgr_bnbwd;

/* ---------------------------------------------------------------------------------------------
 * K8/K11  pointwise (1x1) convolution as a GEMM over NHWC rows — replaces the nn.Conv2d 1x1 layers of
 *   InvertedResidual / conv_1x1_bn (archs/mobilenet_v2.py:15-20,44-45,50-52,58-59) and, with
 *   w_is_kn=1, their input-gradient (autograd's conv dgrad).
 *     out[M,N] = rowop(a)[M,K] * B[K,N] (+ addend[M,N])
 *     B[k][n] = w[n*K + k]  (w_is_kn = 0: w is the conv weight [N_out, K_in], forward)
 *             = w[k*N + n]  (w_is_kn = 1: w is the conv weight [K, N] read as is: dgrad)
 *   w is fp32 (master weights); out/addend have `dtype`.  stats (nullable) = double[2N]:
 *   stats[n] += sum_m out[m,n], stats[N+n] += sum_m out[m,n]^2 (BatchNorm batch statistics, K14),
 *   accumulated from the fp32 accumulators.  K % 8 == 0 and N % 8 == 0 required.
 *   engine: AUTO picks the tcgen05/TMEM tensor-core kernel for bf16 when the shape is supported and
 *   the fp32 SIMT kernel otherwise; SIMT / TCGEN05 force one (TCGEN05 -> EHGR_E_UNSUPPORTED if not).
 * ------------------------------------------------------------------------------------------- */
int ehgr_pw_gemm(const ehgr_rowop* a, const float* w, int w_is_kn, void* out, const void* addend,
                 double* stats, long long M, int K, int N, int dtype, int engine, ehgr_stream_t stream);
/* Same contract with the weights ALSO supplied as a bf16 mirror `w16` (same [N,K] / [K,N] layout as `w`,
 * 16-byte aligned; the caller refreshes it after every optimiser step).  The tensor-core engine then
 * stages the B operand with 16-byte asynchronous copies instead of converting fp32 -> bf16 in every CTA.
 * The SIMT engine (fp32 storage) ignores w16.  w16 == NULL is exactly ehgr_pw_gemm. */
int ehgr_pw_gemm_w16(const ehgr_rowop* a, const float* w, const void* w16, int w_is_kn, void* out,
                     const void* addend, double* stats, long long M, int K, int N, int dtype, int engine,
                     ehgr_stream_t stream);

/* ehgr_pw_gemm_w16 followed, inside the same launch, by the BatchNorm finalisation of `stats` (fin may be NULL). */
int ehgr_pw_gemm_bn(const ehgr_rowop* a, const float* w, const void* w16, int w_is_kn, void* out,
                    const void* addend, double* stats, long long M, int K, int N, int dtype, int engine,
                    const ehgr_bnfin* fin, ehgr_stream_t stream);

/* weight gradient of the same layer: dw[N,K] += sum_m rowop(dy)[m,n] * rowop(a)[m,k]   (fp32, atomics;
 * the caller zeroes dw).  dy is normally a BNBWD operand, a the layer's forward operand. */
int ehgr_pw_wgrad(const ehgr_rowop* dy, const ehgr_rowop* a, float* dw, long long M, int K, int N,
                  int dtype, int engine, ehgr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * N2  dense 3x3 convolution (depth decoder, models/models_MTMM.py:129-155) on the pointwise-GEMM kernels.
 *   forward : ehgr_pw_gemm_bn(a = CONV3 operand, w = wf, w16 = wf16, w_is_kn = 0, M = frames*H*W, K = 9*cin, N = cout)
 *   dgrad   : the same call on d(raw) with a CONV3 operand of cout channels and the flipped weights wd
 *             (K = 9*cout, N = cin); with an upsampled input the result is the gradient w.r.t. the UPSAMPLED tensor and
 *             ehgr_upsample2_bwd folds it back;
 *   wgrad   : ehgr_pw_wgrad(dy, a = the forward CONV3 operand, dwp[cout, 9*cin]) then ehgr_conv3_unpack_grad.
 * ehgr_conv3_pack lays the nn.Conv2d weight w[cout, cin, 3, 3] (fp32) out for those GEMMs, in `dtype`:
 *   wf[n, tap*cin + c] = w[n, c, tap]            (forward:  B[k][n] = wf[n*K + k])
 *   wd[c, tap*cout + n] = w[n, c, 8 - tap]       (dgrad: the spatially flipped, transposed filter)
 * Either output may be NULL.  ehgr_conv3_unpack_grad: dw[n, c, tap] += dwp[n, tap*cin + c].
 * ehgr_upsample2_bwd: g[f, h, w, c] = sum of the 2x2 block of g_up[f, 2h.., 2w.., c] (adjoint of nearest x2), NHWC.
 * ------------------------------------------------------------------------------------------- */
int ehgr_conv3_pack(const float* w, void* wf, void* wd, int cout, int cin, int dtype, ehgr_stream_t stream);
int ehgr_conv3_unpack_grad(const float* dwp, float* dw, int cout, int cin, ehgr_stream_t stream);
/* y[f, 2h+a, 2w+b, c] = x[f, h, w, c]: nn.Upsample(scale_factor=2, mode='nearest') materialised (NHWC) — used when the
 * following convolution gathers its operand with TMA boxes, which cannot fold the x2 index map. */
int ehgr_upsample2_fwd(const void* x, void* y, long long frames, int h, int w, int c, int dtype, ehgr_stream_t stream);
int ehgr_upsample2_bwd(const void* g_up, void* g, long long frames, int h, int w, int c, int dtype, ehgr_stream_t stream);

/* Depth head: Conv2d(C, 1, 1, bias) + Sigmoid (models/models_MTMM.py:151-154) as one pass over the rows.
 *   fwd: out[m] = sigmoid(sum_c rowop(a)[m, c] * w[c] + bias[0])                       (out fp32 [M])
 *   bwd: dz = dout[m] * out[m] * (1 - out[m]);  g_a[m, c] = dz * w[c] (dtype);  dw[c] += sum_m dz * rowop(a)[m, c];
 *        dbias[0] += sum_m dz.  C % 8 == 0, C <= 256. */
int ehgr_depth_head_fwd(const ehgr_rowop* a, const float* w, const float* bias, float* out, long long m, int c, int dtype,
                        ehgr_stream_t stream);
int ehgr_depth_head_bwd(const ehgr_rowop* a, const float* w, const float* out, const float* dout, void* g_a, float* dw,
                        float* dbias, long long m, int c, int dtype, ehgr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K7  depthwise 3x3 convolution, pad 1, stride 1|2 (archs/mobilenet_v2.py:40,54), NHWC.
 *   fwd   : out[nt,ho,wo,c] = sum_{kh,kw} rowop(a)[nt, ho*s+kh-1, wo*s+kw-1, c] * w[c,kh,kw]
 *           (taps outside the image contribute 0 — the padding is applied AFTER the row operand).
 *           stats as in ehgr_pw_gemm (double[2C], nullable).
 *   dgrad : da[nt,h,w,c] = sum over taps of rowop(dy)[...] * w   (gradient w.r.t. rowop(a))
 *   wgrad : dw[c,kh,kw] += sum rowop(dy)[q] * rowop(a)[p(q,kh,kw)]   (fp32 atomics, caller zeroes)
 *   w: fp32 [C,1,3,3] contiguous.  C % 8 == 0.  Ho = (H-1)/s + 1.
 * ------------------------------------------------------------------------------------------- */
int ehgr_dw_fwd(const ehgr_rowop* a, const float* w, void* out, double* stats, int nt, int h, int wd, int c,
                int stride, int dtype, ehgr_stream_t stream);
/* ehgr_dw_fwd + in-launch BatchNorm finalisation of `stats` (fin may be NULL) */
int ehgr_dw_fwd_bn(const ehgr_rowop* a, const float* w, void* out, double* stats, int nt, int h, int wd, int c,
                   int stride, int dtype, const ehgr_bnfin* fin, ehgr_stream_t stream);
int ehgr_dw_dgrad(const ehgr_rowop* dy, const float* w, void* da, int nt, int h, int wd, int c, int stride,
                  int dtype, ehgr_stream_t stream);
int ehgr_dw_wgrad(const ehgr_rowop* dy, const ehgr_rowop* a, float* dw, int nt, int h, int wd, int c,
                  int stride, int dtype, ehgr_stream_t stream);
/* fused backward: da (as ehgr_dw_dgrad) and dw (as ehgr_dw_wgrad) from ONE shared-memory staging of
 * rowop(dy) and rowop(a) per tile — the training path uses this; dgrad/wgrad above are the separable form. */
int ehgr_dw_bwd(const ehgr_rowop* dy, const ehgr_rowop* a, const float* w, void* da, float* dw, int nt, int h,
                int wd, int c, int stride, int dtype, ehgr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K9  stem: dense 3x3 stride-2 pad-1 convolution 3 -> cout from the NCHW network input
 *   (archs/mobilenet_v2.py:7-12,90) to an NHWC raw output + batch statistics; and its weight
 *   gradient (the network input needs no gradient).  x: [nt,3,h,w] of x_dtype; w: fp32 [cout,3,3,3].
 * ------------------------------------------------------------------------------------------- */
int ehgr_stem_fwd(const void* x, const float* w, void* out, double* stats, int nt, int h, int wd, int cout,
                  int x_dtype, int dtype, ehgr_stream_t stream);
int ehgr_stem_fwd_bn(const void* x, const float* w, void* out, double* stats, int nt, int h, int wd, int cout,
                     int x_dtype, int dtype, const ehgr_bnfin* fin, ehgr_stream_t stream);
int ehgr_stem_wgrad(const ehgr_rowop* dy, const void* x, float* dw, int nt, int h, int wd, int cout,
                    int x_dtype, int dtype, ehgr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K14  BatchNorm2d bookkeeping (nn.BatchNorm2d semantics: biased variance normalises, unbiased
 *   variance feeds running_var, momentum update; eval mode uses the running statistics).
 *   finalize     : stats (double[2C]: sum, sum of squares over `count` elements per channel) ->
 *                  scale = gamma*invstd, shift = beta - mean*scale, mean, invstd (all fp32 [C]);
 *                  training != 0 also updates running_mean/var in place.  training == 0 ignores
 *                  stats/count and uses running_mean/var.
 *   bwd_reduce   : sums[c] += sum_m mask*g[m,c];  sums[C+c] += sum_m mask*g[m,c]*raw[m,c]
 *                  mask = relu6 ? (0 < raw*scale+shift < hi) : 1, hi = 6 (relu6 = 1) or +inf (relu6 = 2: plain ReLU)
 *                  (double[2C], caller zeroes)
 *   bwd_finalize : sums -> (ca, cb, cc) of the BNBWD row operand and dgamma, dbeta.
 * ------------------------------------------------------------------------------------------- */
int ehgr_bn_finalize(const double* stats, long long count, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, float momentum, float eps, int training,
                     float* scale, float* shift, float* mean, float* invstd, int c, ehgr_stream_t stream);
int ehgr_bn_bwd_reduce(const void* g, const void* raw, const float* scale, const float* shift, int relu6,
                       double* sums, long long m, int c, int dtype, ehgr_stream_t stream);
int ehgr_bn_bwd_finalize(const double* sums, long long count, const float* gamma, const float* mean,
                         const float* invstd, int training, float* ca, float* cb, float* cc, float* dgamma,
                         float* dbeta, int c, ehgr_stream_t stream);
/* bwd_reduce + bwd_finalize in ONE launch: the last CTA of the reduction turns the sums into the coefficients */
int ehgr_bn_bwd_reduce_fin(const void* g, const void* raw, const float* scale, const float* shift, int relu6,
                           double* sums, long long m, int c, int dtype, const ehgr_bnbwd* fin,
                           ehgr_stream_t stream);

/* out[M,C] = rowop(a) (+ addend): materialises a lazy activation — BatchNorm(+ReLU6)(+residual add,
 * InvertedResidual.forward archs/mobilenet_v2.py:62-66) — or, with a SHIFT operand of shift_dir=-1,
 * the input gradient of a shifted block plus the residual gradient. */
int ehgr_row_apply(const ehgr_rowop* a, const void* addend, void* out, long long m, int c, int dtype,
                   ehgr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * N4  device-side end of the input pipeline: uint8 frames -> normalised float planes.
 *   dst[p][i] = (float(src[p][i]) / div - mean[c]) / std[c],  c = p % channels, evaluated in fp32 in
 *   exactly the order ToTorchFormatTensor(div=True) + GroupNormalize use on the CPU
 *   (models/spatial_transforms.py:66-80,489-503), so an fp32 dst is bit-identical to the reference's
 *   CPU tensors.  src: [n_planes, plane] uint8 (NCHW frames: n_planes = frames*channels, plane = H*W),
 *   dst: same shape, fp32 or bf16.  mean/std: device float[channels], NULL = 0 / 1 (depth maps: div only).
 * ------------------------------------------------------------------------------------------- */
int ehgr_normalize_u8(const void* src, void* dst, long long n_planes, int channels, long long plane,
                      const float* mean, const float* stdv, float div, int dst_dtype, ehgr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * N4  TemporalPool.temporal_pool (models/temporal_shift.py:89-98, duplicate models/action.py:165-176): max over the
 *   frames {2t'-1, 2t', 2t'+1} of a clip, stride 2 (max_pool3d kernel (3,1,1), stride (2,1,1), padding (1,0,0)).
 *   x: [n, t_in, frame_elems], out: [n, (t_in-1)/2+1, frame_elems]; frame_elems = c*h*w of either layout (16-byte
 *   vectors when it is a multiple of the vector width, single elements otherwise).  bwd recomputes the first maximum of every window from x (strict '>' scan, as max_pool3d):
 *   dx[n, t_in, frame_elems] is written in full, no atomics.
 * ------------------------------------------------------------------------------------------- */
int ehgr_temporal_pool_fwd(const void* x, void* out, long long n, int t_in, long long frame_elems, int dtype,
                           ehgr_stream_t stream);
int ehgr_temporal_pool_bwd(const void* x, const void* g, void* dx, long long n, int t_in, long long frame_elems,
                           int dtype, ehgr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * N1  optimiser step: torch.optim.SGD (momentum, weight decay, dampening 0, no Nesterov) over the policy
 *   groups of get_optim_policies (train_mtmm.py:576-585; lr_mult / decay_mult per group) as one kernel
 *   over flat fp32 buffers:  d = g + wd*decay_mult[k]*p;  buf = momentum*buf + d;  p -= lr*lr_mult[k]*buf.
 *   code: one byte per element = its group k (255 = padding, untouched); n % 4 == 0; buf starts at zero
 *   (reproduces torch's first step).  lr is read from DEVICE memory (float[1]): a learning-rate schedule
 *   (utils.py:39-46) only rewrites that scalar and never invalidates a captured CUDA graph.
 *   ema (optional, NULL = off): the EMA copy of the parameters, laid out like p; updated in the same pass with the
 *   freshly stepped parameter, ema = decay*ema + (1-decay)*p — EMAWrapper.update (train_mtmm.py:110-128), which the
 *   reference calls right after optimizer.step() (train_mtmm.py:242-245).  The expression is evaluated as the reference
 *   does (two fp32 products and one fp32 sum, scalars rounded to fp32): bit-identical.
 *   p16 (optional, NULL = off): bf16 mirror of p (same element offsets), rewritten in the same pass: the tensor-core
 *   GEMMs stage their weight operand from it with asynchronous copies (ehgr_pw_gemm_w16) — no per-layer cast kernels.
 * ehgr_ema_update: the same update for the other state_dict entries — floating-point buffers (is_int64 = 0: BatchNorm
 *   running statistics) and the int64 num_batches_tracked counters (is_int64 = 1: evaluated in fp32 and truncated,
 *   as python_float * int64_tensor followed by copy_() does in the reference).
 * ------------------------------------------------------------------------------------------- */
int ehgr_sgd_step(float* p, const float* g, float* buf, const void* code, const float* lr_mult,
                  const float* decay_mult, int n_groups, const float* lr_dev, float momentum, float weight_decay,
                  long long n, float* ema, double ema_decay, void* p16, ehgr_stream_t stream);
int ehgr_ema_update(void* ema, const void* x, long long n, double decay, int is_int64, ehgr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K10  classifier head: x.mean(3).mean(2) (archs/mobilenet_v2.py:112), new_fc and the segment
 *   consensus (models/models.py:341-356, models/basic_ops.py:9-37).
 *   pool_fwd : pooled[nt,c] (fp32) = mean over hw rows of rowop(a)
 *   pool_bwd : da[nt*hw, c] = dpooled[nt,c] / hw
 *   fc_consensus_fwd : meanfeat[n,f] = mean_t feat[n*T+t, f];  logits[n,k] = bias[k] + sum_f W[k,f]*meanfeat[n,f]
 *   fc_consensus_bwd : dfeat[n*T+t,f] = (1/T) sum_k dlogits[n,k] W[k,f];
 *                      dW[k,f] += sum_n dlogits[n,k] meanfeat[n,f];  db[k] += sum_n dlogits[n,k]
 *                      (fp32 atomics; the caller zeroes dW, db)
 * ------------------------------------------------------------------------------------------- */
int ehgr_pool_fwd(const ehgr_rowop* a, float* pooled, int nt, int hw, int c, int dtype, ehgr_stream_t stream);
int ehgr_pool_bwd(const float* dpooled, void* da, int nt, int hw, int c, int dtype, ehgr_stream_t stream);
int ehgr_fc_consensus_fwd(const float* feat, const float* w, const float* bias, float* meanfeat, float* logits,
                          int n, int n_segment, int f, int k, ehgr_stream_t stream);
int ehgr_fc_consensus_bwd(const float* dlogits, const float* meanfeat, const float* w, float* dfeat, float* dw,
                          float* dbias, int n, int n_segment, int f, int k, ehgr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K12  MTMM loss, forward and backward in one launch (train_mtmm.py:223-231):
 *   loss = CE(logits, labels) + depth_weight * MSE(pred, bilinear_{gh x gw -> ph x pw}(depth_gt))
 *   with gh = 4*ph, gw = 4*pw (align_corners=False => mean of the 2x2 centre pixels of each 4x4 cell).
 *   logits fp32 [n,k]; labels int64 [n]; pred `dtype` [frames,ph,pw]; depth_gt fp32 [frames,gh,gw].
 *   out: loss_out[0] = total, loss_out[1] = CE, loss_out[2] = MSE (fp32, caller zeroes);
 *        dlogits fp32 [n,k]; dpred fp32 [frames,ph,pw]   (gradients of `total`).
 *   n == 0 (logits / labels / dlogits may then be NULL): the depth term alone — the MTMM+SD step (train_mtmm_sd.py:240-293)
 *   adds it, weighted (1 - alpha) * 0.01, to the terms of the self-distillation kernel below.
 * K13  self-distillation loss (train_sd.py:178-193,227-265), forward+backward in one launch:
 *   logits: HOST array of 4 device pointers fp32 [n,k] (final, mid1..3); feats: HOST array of 4 device
 *   pointers fp32 [rows,f] (final, mid1..3).
 *   terms_out fp32[11] = total, 4 CE, 3 KD (x T^2), 3 feature sums (caller zeroes);
 *   dlogits: HOST array of 4 device pointers [n,k]; dfeats: HOST array of 3 device pointers [rows,f]
 *   (mid1..3; the final feature is detached).
 * ------------------------------------------------------------------------------------------- */
int ehgr_mtmm_loss(const float* logits, const long long* labels, const void* pred, const float* depth_gt,
                   float depth_weight, float* loss_out, float* dlogits, float* dpred, int n, int k,
                   int frames, int ph, int pw, int dtype, ehgr_stream_t stream);
int ehgr_sd_loss(const float* const* logits, const float* const* feats, const long long* labels, float alpha,
                 float beta, float temperature, float* terms_out, float* const* dlogits, float* const* dfeats,
                 int n, int k, long long rows, int f, ehgr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K2-K6  ACTION module (models/action.py:8-116): temporal FIR + spatio-temporal / channel / motion
 *   excitation around a 1x1 convolution.  All small tensors are fp32; x / xs / gy / dxs / dx have `dtype`
 *   and are NHWC rows [n*t*h*w, c].  The struct is a HOST struct of device pointers (copied into the
 *   kernel parameters); the caller owns every buffer and zeroes the accumulators marked (+=).
 *   forward :  ehgr_action_xs -> ehgr_bn_finalize(qstats -> bn3_scale/shift) -> ehgr_action_gates ->
 *              ehgr_pw_gemm with a GATE row operand (in1 = xs, in2 = g1, scale = g2, shift = g3)
 *   backward:  gy = dgrad of that GEMM;  ehgr_action_bwd_reduce -> ehgr_action_bwd_small ->
 *              ehgr_bn_bwd_finalize(bn3_sums -> bn3_ca/cb/cc, dgamma, dbeta) -> ehgr_action_bwd_dxs ->
 *              ehgr_action_fir_bwd (adds the residual gradient `addend`, nullable)
 * ------------------------------------------------------------------------------------------- */
typedef struct ehgr_action {
  int32_t n, t, h, w, c, cr;      /* clips, segments, spatial size, channels, reduced channels c/16 */
  /* parameters (fp32, the reference's module names) */
  const float* shift_w;     /* action_shift.weight      [c,1,3]    */
  const float* p1_w;        /* action_p1_conv1.weight   [1,1,3,3,3] */
  const float* p2_squeeze;  /* action_p2_squeeze.weight [cr,c]     */
  const float* p2_conv1;    /* action_p2_conv1.weight   [cr,cr,3]  */
  const float* p2_expand;   /* action_p2_expand.weight  [c,cr]     */
  const float* p3_squeeze;  /* action_p3_squeeze.weight [cr,c]     */
  const float* p3_conv1;    /* action_p3_conv1.weight   [cr,1,3,3] */
  const float* p3_expand;   /* action_p3_expand.weight  [c,cr]     */
  const float* bn3_scale;   /* action_p3_bn1 as scale/shift [cr] (ehgr_bn_finalize) */
  const float* bn3_shift;
  /* forward products, saved for backward */
  float* mrow;              /* [M]      mean over channels of xs */
  float* pool;              /* [n*t, c] spatial SUM of xs */
  float* q;                 /* [M, cr]  p3_squeeze(xs), before BatchNorm */
  double* qstats;           /* [2*cr]   (+=) sum q, sum q^2 */
  float* g1;                /* [M]      spatio-temporal gate */
  float* g2;                /* [n*t, c] channel gate */
  float* g3;                /* [n*t, c] motion gate */
  float* s;                 /* [n*t, cr] p2_squeeze output */
  float* u;                 /* [n*t, cr] p2_conv1 output (pre-ReLU) */
  float* pi;                /* [n*t, cr] pooled motion feature */
  /* backward workspaces */
  float* dg1;               /* [M]      sum_c gy*xs, overwritten with d(a1) */
  float* dgc;               /* [n*t, c] sum_hw gy*xs */
  float* dm;                /* [M]      gradient w.r.t. mrow */
  float* dpool;             /* [n*t, c] gradient w.r.t. the spatial mean */
  float* dd;                /* [n*t, cr] gradient w.r.t. the motion difference map (spatially constant) */
  double* bn3_sums;         /* [2*cr]   (+=) sum dx3, sum dx3*q */
  const float* bn3_ca;      /* [cr] BatchNorm-backward coefficients of action_p3_bn1 */
  const float* bn3_cb;
  const float* bn3_cc;
  /* parameter gradients (+=) */
  float* d_shift_w;
  float* d_p1_w;
  float* d_p2_squeeze;
  float* d_p2_conv1;
  float* d_p2_expand;
  float* d_p3_squeeze;
  float* d_p3_conv1;
  float* d_p3_expand;
} ehgr_action;

int ehgr_action_xs(const ehgr_action* a, const void* x, void* xs, int dtype, ehgr_stream_t stream);
int ehgr_action_gates(const ehgr_action* a, ehgr_stream_t stream);
int ehgr_action_bwd_reduce(const ehgr_action* a, const void* gy, const void* xs, int dtype, ehgr_stream_t stream);
int ehgr_action_bwd_small(const ehgr_action* a, ehgr_stream_t stream);
int ehgr_action_bwd_dxs(const ehgr_action* a, const void* gy, const void* xs, void* dxs, int dtype,
                        ehgr_stream_t stream);
int ehgr_action_fir_bwd(const ehgr_action* a, const void* dxs, const void* x, const void* addend, void* dx, int dtype,
                        ehgr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * N3  torchvision ResNet bottleneck path (csrc/resnet.cu) — the reference builds torchvision.models.resnet50 for
 *     base_model='resnet50' (models/models.py:108-117) and wraps conv1 of every bottleneck with a temporal module
 *     (models/temporal_shift.py:101-146).  The 1x1 / 3x3 convolutions of a bottleneck run on ehgr_pw_gemm* /
 *     ehgr_pw_wgrad (PLAIN / AFFINE / SHIFT / CONV3 operands); these entry points are the rest.  All activations NHWC.
 *
 *   ehgr_stem7_fwd   : replaces ResNet.conv1 = Conv2d(3, 64, 7, stride=2, padding=3, bias=False).  x [frames,3,h,w] NCHW
 *                      (x_dtype), w [64,3,7,7] fp32, out [frames, ho, wo, 64] raw (before BatchNorm), ho = (h-1)/2+1;
 *                      stats (optional, += ) [2*64] doubles: per-channel sum and sum of squares of the stored values.
 *   ehgr_stem7_wgrad : dw [64,3,7,7] += rowop(dy)^T * patches(x); dy is usually a BNBWD operand over (g, raw).
 *   ehgr_stem7_im2col / ehgr_stem7_pack / ehgr_stem7_unpack_grad : the same convolution as a tensor-core GEMM (bf16 path):
 *                      a [frames*ho*wo, kp] = the 147-column patch matrix (column t = (ci*7+ky)*7+kx, zero padded to kp, a
 *                      multiple of 8), wp [cout, kp] = the filter in that column order (fp32 or bf16), and
 *                      dw [cout,3,7,7] += dwp [cout, kp][:, :147]; the GEMM itself is ehgr_pw_gemm_bn / ehgr_pw_wgrad.
 *   ehgr_maxpool3_fwd: replaces ReLU + MaxPool2d(3, stride=2, padding=1) after bn1: y [frames,ho,wo,c] = window max of
 *                      rowop(a) (an AFFINE operand = lazy BatchNorm+ReLU of the stem output); idx (uint8, same shape
 *                      as y) = winning tap 3*ky+kx, first maximum in scan order as at::max_pool2d.
 *   ehgr_maxpool3_bwd: gx [frames,h,w,c] = gradient w.r.t. the pooled tensor's input (gather over <= 4 windows, no atomics).
 *   ehgr_subsample2_fwd: y = x[:, ::2, ::2, :] (+ statistics of y when stats != NULL).  conv(k, stride 2, padding (k-1)/2)
 *                      == subsample2(conv(k, stride 1)), so layer{2,3,4}.0.conv2 and the downsample 1x1 (torchvision
 *                      Bottleneck / ResNet._make_layer) reuse the stride-1 GEMMs.
 *   ehgr_subsample2_bwd: the adjoint: gx [frames,h,w,c] = g at even (h, w), zero elsewhere.
 *   ehgr_bn_add_relu : out = relu(raw*scale + shift + addend)  — `out = relu(bn3(conv3(.)) + identity)` (Bottleneck.forward).
 *   ehgr_relu_bwd    : gz = g where out > 0 else 0, n elements (a multiple of the 16-byte vector).
 * ------------------------------------------------------------------------------------------- */
int ehgr_stem7_fwd(const void* x, const float* w, void* out, double* stats, long long frames, int h, int w_in, int cout,
                   int x_dtype, int out_dtype, ehgr_stream_t stream);
int ehgr_stem7_wgrad(const ehgr_rowop* dy, const void* x, float* dw, long long frames, int h, int w_in, int cout,
                     int x_dtype, int dtype, ehgr_stream_t stream);
int ehgr_stem7_im2col(const void* x, void* a, long long frames, int h, int w_in, int kp, int x_dtype, int dtype,
                      ehgr_stream_t stream);
int ehgr_stem7_pack(const float* w, void* wp, int cout, int kp, int dtype, ehgr_stream_t stream);
int ehgr_stem7_unpack_grad(const float* dwp, float* dw, int cout, int kp, ehgr_stream_t stream);
int ehgr_maxpool3_fwd(const ehgr_rowop* a, void* y, void* idx, long long frames, int h, int w, int c, int dtype,
                      ehgr_stream_t stream);
int ehgr_maxpool3_bwd(const void* g, const void* idx, void* gx, long long frames, int h, int w, int c, int dtype,
                      ehgr_stream_t stream);
int ehgr_subsample2_fwd(const void* x, void* y, double* stats, long long frames, int h, int w, int c, int dtype,
                        ehgr_stream_t stream);
int ehgr_subsample2_bwd(const void* g, void* gx, long long frames, int h, int w, int c, int dtype, ehgr_stream_t stream);
int ehgr_bn_add_relu(const void* raw, const float* scale, const float* shift, const void* addend, void* out, long long m,
                     int c, int dtype, ehgr_stream_t stream);
int ehgr_relu_bwd(const void* g, const void* out, void* gz, long long n, int dtype, ehgr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EHGR_B200_H_ */
