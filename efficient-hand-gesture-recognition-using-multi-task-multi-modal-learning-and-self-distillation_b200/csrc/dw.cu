// dw.cu — K7 depthwise 3x3 convolution (pad 1, stride 1|2) on NHWC rows: forward (+ batch statistics),
// input gradient and weight gradient.  Reference ops replaced: the grouped nn.Conv2d of
// archs/mobilenet_v2.py:40,54 with its BatchNorm/ReLU6 neighbours folded into the row operands
// (the input is read through an AFFINE operand = BN+ReLU6 of the previous layer; the gradient is read
// through a BNBWD operand), and autograd's grouped-conv dgrad / wgrad.
//
// Memory-bound.  Thread block = (C/VEC channel vectors) x (P pixels): a thread owns ONE 16-byte channel
// vector for its whole life (row-operand coefficients and, in wgrad, the 9xVEC accumulators live in
// registers) and walks over pixels with a grid-stride loop (persistent blocks), so the per-block
// reductions (statistics, weight gradients) are flushed once per block.  Every pixel is processed in
// two phases: FETCH issues all tap loads (out-of-image taps are clamped to a valid pixel and masked
// later, so no load sits behind a branch), FINISH does the arithmetic — one memory round trip per
// pixel instead of nine.  Channels are innermost: every warp access is a run of full 16-byte vectors;
// the 3x3 re-use is served by L1/L2.
#include "rowop.cuh"
#include "bnfin.cuh"

namespace ehgr {

struct DwGeom {
  int nt, h, w, c, stride, ho, wo;
  uint32_t n_out;  // nt*ho*wo
  uint32_t n_in;   // nt*h*w
};

template <typename T, bool kTwo>
__global__ void __launch_bounds__(256)
dw_fwd_kernel(RowOp a, const float* __restrict__ wgt, T* __restrict__ out, double* __restrict__ stats, DwGeom g) {
  constexpr int V = VecOf<T>::N;
  using Loader = RowLoader<T, V, kTwo>;
  extern __shared__ float smem[];
  float* ws = smem;                 // [9][C]
  float* s_sum = smem + 9 * g.c;    // [2C]
  const int nthreads = blockDim.x * blockDim.y;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < 9 * g.c; i += nthreads) {
    const int tap = i / g.c, c = i - tap * g.c;
    ws[i] = wgt[c * 9 + tap];
  }
  if (stats) for (int i = tid; i < 2 * g.c; i += nthreads) s_sum[i] = 0.f;
  __syncthreads();

  const int c0 = threadIdx.x * V;
  Loader ld;
  ld.init(a, c0, g.c);
  float tsum[V], tsq[V];
#pragma unroll
  for (int i = 0; i < V; ++i) tsum[i] = tsq[i] = 0.f;
  const uint32_t stride_q = gridDim.x * blockDim.y;
  for (uint32_t q = blockIdx.x * blockDim.y + threadIdx.y; q < g.n_out; q += stride_q) {
    const uint32_t r = q / g.wo;
    const int wo = static_cast<int>(q - r * g.wo);
    const uint32_t nt = r / g.ho;
    const int ho = static_cast<int>(r - nt * g.ho);
    const int hc = ho * g.stride, wc = wo * g.stride;      // centre tap: always inside the image
    const long long base = static_cast<long long>(nt) * g.h;
    typename Loader::Raw raw[9];
    bool ok[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int hi = hc + t / 3 - 1, wi = wc + t % 3 - 1;
      ok[t] = hi >= 0 && hi < g.h && wi >= 0 && wi < g.w;
      raw[t] = ld.fetch(a, (base + (ok[t] ? hi : hc)) * g.w + (ok[t] ? wi : wc));
    }
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float v[V], wv[V];
      ld.finish(a, raw[t], v);
      load_vec<float, V>(ws + t * g.c + c0, wv);
      if (ok[t]) {
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = fmaf(v[i], wv[i], acc[i]);
      }
    }
    store_vec<T, V>(out + static_cast<long long>(q) * g.c + c0, acc);
#pragma unroll
    for (int i = 0; i < V; ++i) { tsum[i] += acc[i]; tsq[i] = fmaf(acc[i], acc[i], tsq[i]); }
  }
  if (stats) {
#pragma unroll
    for (int i = 0; i < V; ++i) { atomicAdd(&s_sum[c0 + i], tsum[i]); atomicAdd(&s_sum[g.c + c0 + i], tsq[i]); }
    __syncthreads();
    for (int i = tid; i < 2 * g.c; i += nthreads) atomicAdd(&stats[i], static_cast<double>(s_sum[i]));
  }
}

// da[p] = sum over taps (kh,kw) with (hi+1-kh) % s == 0, (wi+1-kw) % s == 0 of dy[(hi+1-kh)/s, (wi+1-kw)/s] * w[kh][kw]
// stride 1: nine candidate taps; stride 2: at most 2 x 2 (kh in {kh0, kh0+2}, kw likewise).
template <typename T, bool kTwo, int STRIDE>
__global__ void __launch_bounds__(256)
dw_dgrad_kernel(RowOp dy, const float* __restrict__ wgt, T* __restrict__ da, DwGeom g) {
  constexpr int V = VecOf<T>::N;
  constexpr int NS = STRIDE == 1 ? 3 : 2;   // candidate taps per axis
  using Loader = RowLoader<T, V, kTwo>;
  extern __shared__ float smem[];
  float* ws = smem;  // [9][C]
  const int nthreads = blockDim.x * blockDim.y;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < 9 * g.c; i += nthreads) {
    const int tap = i / g.c, c = i - tap * g.c;
    ws[i] = wgt[c * 9 + tap];
  }
  __syncthreads();
  const int c0 = threadIdx.x * V;
  Loader ld;
  ld.init(dy, c0, g.c);
  const uint32_t stride_p = gridDim.x * blockDim.y;
  for (uint32_t p = blockIdx.x * blockDim.y + threadIdx.y; p < g.n_in; p += stride_p) {
    const uint32_t r = p / g.w;
    const int wi = static_cast<int>(p - r * g.w);
    const uint32_t nt = r / g.h;
    const int hi = static_cast<int>(r - nt * g.h);
    const long long base = static_cast<long long>(nt) * g.ho;
    const int kh0 = STRIDE == 1 ? 0 : ((hi + 1) & 1), kw0 = STRIDE == 1 ? 0 : ((wi + 1) & 1);
    typename Loader::Raw raw[NS * NS];
    bool ok[NS * NS];
    int tap[NS * NS];
#pragma unroll
    for (int t = 0; t < NS * NS; ++t) {
      const int kh = kh0 + (t / NS) * STRIDE, kw = kw0 + (t % NS) * STRIDE;
      const int th = hi + 1 - kh, tw = wi + 1 - kw;
      const int ho = STRIDE == 1 ? th : th >> 1, wo = STRIDE == 1 ? tw : tw >> 1;
      ok[t] = kh < 3 && kw < 3 && th >= 0 && tw >= 0 && ho < g.ho && wo < g.wo;
      tap[t] = ok[t] ? kh * 3 + kw : 0;
      // clamp to a valid output pixel when the tap is outside
      const int hs = min(max(ho, 0), g.ho - 1), wsafe = min(max(wo, 0), g.wo - 1);
      raw[t] = ld.fetch(dy, (base + hs) * g.wo + wsafe);
    }
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
    for (int t = 0; t < NS * NS; ++t) {
      float v[V], wv[V];
      ld.finish(dy, raw[t], v);
      load_vec<float, V>(ws + tap[t] * g.c + c0, wv);
      if (ok[t]) {
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = fmaf(v[i], wv[i], acc[i]);
      }
    }
    store_vec<T, V>(da + static_cast<long long>(p) * g.c + c0, acc);
  }
}

// dw[c][tap] += sum_q dy[q][c] * a[p(q,tap)][c]
template <typename T, bool kTwoDy>
__global__ void __launch_bounds__(256)
dw_wgrad_kernel(RowOp dy, RowOp a, float* __restrict__ dwgt, DwGeom g) {
  constexpr int V = VecOf<T>::N;
  using LoaderDy = RowLoader<T, V, kTwoDy>;
  using LoaderA = RowLoader<T, V, false>;
  extern __shared__ float smem[];
  float* s_acc = smem;  // [9][C]
  const int nthreads = blockDim.x * blockDim.y;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < 9 * g.c; i += nthreads) s_acc[i] = 0.f;
  __syncthreads();
  const int c0 = threadIdx.x * V;
  LoaderDy ld_dy;
  LoaderA ld_a;
  ld_dy.init(dy, c0, g.c);
  ld_a.init(a, c0, g.c);
  float acc[9][V];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[t][i] = 0.f;
  const uint32_t stride_q = gridDim.x * blockDim.y;
  for (uint32_t q = blockIdx.x * blockDim.y + threadIdx.y; q < g.n_out; q += stride_q) {
    const uint32_t r = q / g.wo;
    const int wo = static_cast<int>(q - r * g.wo);
    const uint32_t nt = r / g.ho;
    const int ho = static_cast<int>(r - nt * g.ho);
    const int hc = ho * g.stride, wc = wo * g.stride;
    const long long base = static_cast<long long>(nt) * g.h;
    const typename LoaderDy::Raw raw_d = ld_dy.fetch(dy, q);
    typename LoaderA::Raw raw[9];
    bool ok[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int hi = hc + t / 3 - 1, wi = wc + t % 3 - 1;
      ok[t] = hi >= 0 && hi < g.h && wi >= 0 && wi < g.w;
      raw[t] = ld_a.fetch(a, (base + (ok[t] ? hi : hc)) * g.w + (ok[t] ? wi : wc));
    }
    float d[V];
    ld_dy.finish(dy, raw_d, d);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float v[V];
      ld_a.finish(a, raw[t], v);
      if (ok[t]) {
#pragma unroll
        for (int i = 0; i < V; ++i) acc[t][i] = fmaf(d[i], v[i], acc[t][i]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < V; ++i) atomicAdd(&s_acc[t * g.c + c0 + i], acc[t][i]);
  __syncthreads();
  for (int i = tid; i < 9 * g.c; i += nthreads) {
    const int tap = i / g.c, c = i - tap * g.c;
    atomicAdd(&dwgt[c * 9 + tap], s_acc[i]);
  }
}

static int dw_geom(DwGeom& g, int nt, int h, int w, int c, int stride, int dtype) {
  const int es = esize_of(dtype);
  if (es == 0) return EHGR_E_DTYPE;
  const int V = 16 / es;
  if (nt < 0 || h <= 0 || w <= 0 || c <= 0 || (c % V) || (stride != 1 && stride != 2)) return EHGR_E_SHAPE;
  if (c / V > 256) return EHGR_E_UNSUPPORTED;
  g.nt = nt; g.h = h; g.w = w; g.c = c; g.stride = stride;
  g.ho = (h - 1) / stride + 1;
  g.wo = (w - 1) / stride + 1;
  const long long n_out = static_cast<long long>(nt) * g.ho * g.wo, n_in = static_cast<long long>(nt) * h * w;
  if (n_in >= 0x7fffffffLL) return EHGR_E_SHAPE;
  g.n_out = static_cast<uint32_t>(n_out);
  g.n_in = static_cast<uint32_t>(n_in);
  return EHGR_OK;
}

static dim3 dw_block(int c, int V) {
  const int cv = c / V;
  int p = 256 / cv;
  if (p < 1) p = 1;
  return dim3(cv, p);
}

// persistent grid: enough blocks to fill the machine, never more than the work needs
static unsigned dw_grid(uint32_t pixels, const dim3& block, int per_sm) {
  const long long need = cdiv(pixels, block.y);
  return static_cast<unsigned>(std::max(1LL, std::min(need, static_cast<long long>(kNumSMs) * per_sm)));
}

template <typename T>
static int dw_fwd_launch(const RowOp& a, const float* w, void* out, double* stats, const DwGeom& g, cudaStream_t s) {
  const dim3 block = dw_block(g.c, VecOf<T>::N);
  const unsigned grid = dw_grid(g.n_out, block, 8);
  const size_t smem = static_cast<size_t>(11) * g.c * sizeof(float);
  if (a.mode == EHGR_ROW_BNBWD)
    dw_fwd_kernel<T, true><<<grid, block, smem, s>>>(a, w, static_cast<T*>(out), stats, g);
  else
    dw_fwd_kernel<T, false><<<grid, block, smem, s>>>(a, w, static_cast<T*>(out), stats, g);
  return launch_status();
}

template <typename T>
static int dw_dgrad_launch(const RowOp& dy, const float* w, void* da, const DwGeom& g, cudaStream_t s) {
  const dim3 block = dw_block(g.c, VecOf<T>::N);
  const unsigned grid = dw_grid(g.n_in, block, 8);
  const size_t smem = static_cast<size_t>(9) * g.c * sizeof(float);
  const bool two = dy.mode == EHGR_ROW_BNBWD;
  T* o = static_cast<T*>(da);
  if (g.stride == 1) {
    if (two) dw_dgrad_kernel<T, true, 1><<<grid, block, smem, s>>>(dy, w, o, g);
    else dw_dgrad_kernel<T, false, 1><<<grid, block, smem, s>>>(dy, w, o, g);
  } else {
    if (two) dw_dgrad_kernel<T, true, 2><<<grid, block, smem, s>>>(dy, w, o, g);
    else dw_dgrad_kernel<T, false, 2><<<grid, block, smem, s>>>(dy, w, o, g);
  }
  return launch_status();
}

template <typename T>
static int dw_wgrad_launch(const RowOp& dy, const RowOp& a, float* dw, const DwGeom& g, cudaStream_t s) {
  const dim3 block = dw_block(g.c, VecOf<T>::N);
  const unsigned grid = dw_grid(g.n_out, block, 4);
  const size_t smem = static_cast<size_t>(9) * g.c * sizeof(float);
  if (dy.mode == EHGR_ROW_BNBWD) dw_wgrad_kernel<T, true><<<grid, block, smem, s>>>(dy, a, dw, g);
  else dw_wgrad_kernel<T, false><<<grid, block, smem, s>>>(dy, a, dw, g);
  return launch_status();
}

}  // namespace ehgr

namespace ehgr {
int dw_fwd_tiled(const RowOp& a, const float* w, void* out, double* stats, int nt, int h, int wd, int c, int stride,
                 int dtype, cudaStream_t s);
}
using namespace ehgr;

extern "C" int ehgr_dw_fwd(const ehgr_rowop* a, const float* w, void* out, double* stats, int nt, int h, int wd,
                           int c, int stride, int dtype, ehgr_stream_t stream) {
  return ehgr_dw_fwd_bn(a, w, out, stats, nt, h, wd, c, stride, dtype, nullptr, stream);
}

extern "C" int ehgr_dw_fwd_bn(const ehgr_rowop* a, const float* w, void* out, double* stats, int nt, int h, int wd,
                              int c, int stride, int dtype, const ehgr_bnfin* fin, ehgr_stream_t stream) {
  if (fin) {
    if (!fin->scale || !fin->shift || !fin->counter) return EHGR_E_NULL;
    if (fin->training && (!stats || fin->count <= 0)) return EHGR_E_NULL;
    if (!fin->training && (!fin->running_mean || !fin->running_var)) return EHGR_E_NULL;
  }
  DwGeom g;
  if (int st = dw_geom(g, nt, h, wd, c, stride, dtype)) return st;
  if (!w || !out) return EHGR_E_NULL;
  if (int st = validate_rowop_nogate(a, esize_of(dtype))) return st;
  if (!aligned_to(out, 16)) return EHGR_E_ALIGN;
  if (g.n_out == 0) return EHGR_OK;
  cudaStream_t s = as_stream(stream);
  fin_slot().fin = fin;            // the bf16 TMA kernel finalises in its last CTA; the fp32 kernels leave it parked
  const int st = dw_fwd_tiled(*a, w, out, stats, nt, h, wd, c, stride, dtype, s);   // shared-memory-tiled kernel (dw_tiled.cu)
  return finish_fin(stats, c, s, st);
}

extern "C" int ehgr_dw_dgrad(const ehgr_rowop* dy, const float* w, void* da, int nt, int h, int wd, int c,
                             int stride, int dtype, ehgr_stream_t stream) {
  DwGeom g;
  if (int st = dw_geom(g, nt, h, wd, c, stride, dtype)) return st;
  if (!w || !da) return EHGR_E_NULL;
  if (int st = validate_rowop_nogate(dy, esize_of(dtype))) return st;
  if (!aligned_to(da, 16)) return EHGR_E_ALIGN;
  if (g.n_in == 0) return EHGR_OK;
  cudaStream_t s = as_stream(stream);
  return dtype == EHGR_F32 ? dw_dgrad_launch<float>(*dy, w, da, g, s)
                           : dw_dgrad_launch<__nv_bfloat16>(*dy, w, da, g, s);
}

extern "C" int ehgr_dw_wgrad(const ehgr_rowop* dy, const ehgr_rowop* a, float* dw, int nt, int h, int wd, int c,
                             int stride, int dtype, ehgr_stream_t stream) {
  DwGeom g;
  if (int st = dw_geom(g, nt, h, wd, c, stride, dtype)) return st;
  if (!dw) return EHGR_E_NULL;
  if (int st = validate_rowop_nogate(dy, esize_of(dtype))) return st;
  if (int st = validate_rowop_nogate(a, esize_of(dtype))) return st;
  if (a->mode == EHGR_ROW_BNBWD) return EHGR_E_UNSUPPORTED;  // the forward operand is never a BN-backward operand
  if (g.n_out == 0) return EHGR_OK;
  cudaStream_t s = as_stream(stream);
  return dtype == EHGR_F32 ? dw_wgrad_launch<float>(*dy, *a, dw, g, s)
                           : dw_wgrad_launch<__nv_bfloat16>(*dy, *a, dw, g, s);
}
