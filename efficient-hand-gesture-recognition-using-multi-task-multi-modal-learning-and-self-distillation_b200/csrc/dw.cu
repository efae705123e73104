// dw.cu — K7 depthwise 3x3 convolution (pad 1, stride 1|2) on NHWC rows: forward (+ batch statistics),
// input gradient and weight gradient.  Reference ops replaced: the grouped nn.Conv2d of
// archs/mobilenet_v2.py:40,54 with its BatchNorm/ReLU6 neighbours folded into the row operands
// (the input is read through an AFFINE operand = BN+ReLU6 of the previous layer; the gradient is read
// through a BNBWD operand), and autograd's grouped-conv dgrad / wgrad.
//
// Memory-bound.  Thread block = (C/VEC channel vectors) x (P pixels): a thread owns ONE 16-byte channel
// vector for its whole life (row-operand coefficients and, in wgrad, the 9xVEC accumulators live in
// registers) and walks over pixels with a grid-stride loop (persistent blocks: grid = k x 148), so
// the per-block reductions (statistics, weight gradients) are flushed once per block.  Channels are
// innermost: every warp access is a run of full 16-byte vectors; the 3x3 re-use is served by L1.
#include "rowop.cuh"

namespace ehgr {

struct DwGeom {
  int nt, h, w, c, stride, ho, wo;
  uint32_t n_out;  // nt*ho*wo
  uint32_t n_in;   // nt*h*w
};

template <typename T>
__global__ void __launch_bounds__(256)
dw_fwd_kernel(RowOp a, const float* __restrict__ wgt, T* __restrict__ out, double* __restrict__ stats, DwGeom g) {
  constexpr int V = VecOf<T>::N;
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
  RowLoader<T, V> ld;
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
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hi = ho * g.stride + kh - 1;
      if (hi < 0 || hi >= g.h) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wi = wo * g.stride + kw - 1;
        if (wi < 0 || wi >= g.w) continue;
        float v[V], wv[V];
        ld.load(a, (static_cast<long long>(nt) * g.h + hi) * g.w + wi, v);
        load_vec<float, V>(ws + (kh * 3 + kw) * g.c + c0, wv);
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

// da[p] = sum_{kh,kw : (hi+1-kh) % s == 0} dy[(hi+1-kh)/s, (wi+1-kw)/s] * w[kh][kw]
template <typename T>
__global__ void __launch_bounds__(256)
dw_dgrad_kernel(RowOp dy, const float* __restrict__ wgt, T* __restrict__ da, DwGeom g) {
  constexpr int V = VecOf<T>::N;
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
  RowLoader<T, V> ld;
  ld.init(dy, c0, g.c);
  const uint32_t stride_p = gridDim.x * blockDim.y;
  for (uint32_t p = blockIdx.x * blockDim.y + threadIdx.y; p < g.n_in; p += stride_p) {
    const uint32_t r = p / g.w;
    const int wi = static_cast<int>(p - r * g.w);
    const uint32_t nt = r / g.h;
    const int hi = static_cast<int>(r - nt * g.h);
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int th = hi + 1 - kh;
      if (th < 0 || (g.stride == 2 && (th & 1))) continue;
      const int ho = g.stride == 2 ? th >> 1 : th;
      if (ho >= g.ho) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int tw = wi + 1 - kw;
        if (tw < 0 || (g.stride == 2 && (tw & 1))) continue;
        const int wo = g.stride == 2 ? tw >> 1 : tw;
        if (wo >= g.wo) continue;
        float v[V], wv[V];
        ld.load(dy, (static_cast<long long>(nt) * g.ho + ho) * g.wo + wo, v);
        load_vec<float, V>(ws + (kh * 3 + kw) * g.c + c0, wv);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = fmaf(v[i], wv[i], acc[i]);
      }
    }
    store_vec<T, V>(da + static_cast<long long>(p) * g.c + c0, acc);
  }
}

// dw[c][tap] += sum_q dy[q][c] * a[p(q,tap)][c]
template <typename T>
__global__ void __launch_bounds__(256)
dw_wgrad_kernel(RowOp dy, RowOp a, float* __restrict__ dwgt, DwGeom g) {
  constexpr int V = VecOf<T>::N;
  extern __shared__ float smem[];
  float* s_acc = smem;  // [9][C]
  const int nthreads = blockDim.x * blockDim.y;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < 9 * g.c; i += nthreads) s_acc[i] = 0.f;
  __syncthreads();
  const int c0 = threadIdx.x * V;
  RowLoader<T, V> ld_dy, ld_a;
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
    float d[V];
    ld_dy.load(dy, q, d);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hi = ho * g.stride + kh - 1;
      if (hi < 0 || hi >= g.h) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wi = wo * g.stride + kw - 1;
        if (wi < 0 || wi >= g.w) continue;
        float v[V];
        ld_a.load(a, (static_cast<long long>(nt) * g.h + hi) * g.w + wi, v);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[kh * 3 + kw][i] = fmaf(d[i], v[i], acc[kh * 3 + kw][i]);
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

// persistent grid: enough blocks to fill the machine (`waves` x 148 x resident blocks), never more than
// the work needs
static unsigned dw_grid(uint32_t pixels, const dim3& block, int per_sm) {
  const long long need = cdiv(pixels, block.y);
  return static_cast<unsigned>(std::max(1LL, std::min(need, static_cast<long long>(kNumSMs) * per_sm)));
}

}  // namespace ehgr

using namespace ehgr;

extern "C" int ehgr_dw_fwd(const ehgr_rowop* a, const float* w, void* out, double* stats, int nt, int h, int wd,
                           int c, int stride, int dtype, ehgr_stream_t stream) {
  DwGeom g;
  if (int st = dw_geom(g, nt, h, wd, c, stride, dtype)) return st;
  if (!w || !out) return EHGR_E_NULL;
  if (int st = validate_rowop(a, esize_of(dtype))) return st;
  if (!aligned_to(out, 16)) return EHGR_E_ALIGN;
  if (g.n_out == 0) return EHGR_OK;
  const int V = 16 / esize_of(dtype);
  const dim3 block = dw_block(c, V);
  const unsigned grid = dw_grid(g.n_out, block, 16);
  const size_t smem = static_cast<size_t>(11) * c * sizeof(float);
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32)
    dw_fwd_kernel<float><<<grid, block, smem, s>>>(*a, w, static_cast<float*>(out), stats, g);
  else
    dw_fwd_kernel<__nv_bfloat16><<<grid, block, smem, s>>>(*a, w, static_cast<__nv_bfloat16*>(out), stats, g);
  return launch_status();
}

extern "C" int ehgr_dw_dgrad(const ehgr_rowop* dy, const float* w, void* da, int nt, int h, int wd, int c,
                             int stride, int dtype, ehgr_stream_t stream) {
  DwGeom g;
  if (int st = dw_geom(g, nt, h, wd, c, stride, dtype)) return st;
  if (!w || !da) return EHGR_E_NULL;
  if (int st = validate_rowop(dy, esize_of(dtype))) return st;
  if (!aligned_to(da, 16)) return EHGR_E_ALIGN;
  if (g.n_in == 0) return EHGR_OK;
  const int V = 16 / esize_of(dtype);
  const dim3 block = dw_block(c, V);
  const unsigned grid = dw_grid(g.n_in, block, 16);
  const size_t smem = static_cast<size_t>(9) * c * sizeof(float);
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32)
    dw_dgrad_kernel<float><<<grid, block, smem, s>>>(*dy, w, static_cast<float*>(da), g);
  else
    dw_dgrad_kernel<__nv_bfloat16><<<grid, block, smem, s>>>(*dy, w, static_cast<__nv_bfloat16*>(da), g);
  return launch_status();
}

extern "C" int ehgr_dw_wgrad(const ehgr_rowop* dy, const ehgr_rowop* a, float* dw, int nt, int h, int wd, int c,
                             int stride, int dtype, ehgr_stream_t stream) {
  DwGeom g;
  if (int st = dw_geom(g, nt, h, wd, c, stride, dtype)) return st;
  if (!dw) return EHGR_E_NULL;
  if (int st = validate_rowop(dy, esize_of(dtype))) return st;
  if (int st = validate_rowop(a, esize_of(dtype))) return st;
  if (g.n_out == 0) return EHGR_OK;
  const int V = 16 / esize_of(dtype);
  const dim3 block = dw_block(c, V);
  const unsigned grid = dw_grid(g.n_out, block, 4);
  const size_t smem = static_cast<size_t>(9) * c * sizeof(float);
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32)
    dw_wgrad_kernel<float><<<grid, block, smem, s>>>(*dy, *a, dw, g);
  else
    dw_wgrad_kernel<__nv_bfloat16><<<grid, block, smem, s>>>(*dy, *a, dw, g);
  return launch_status();
}
