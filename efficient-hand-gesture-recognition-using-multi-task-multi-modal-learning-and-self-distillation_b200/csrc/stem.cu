// stem.cu — K9: the network's first layer, dense 3x3 stride-2 pad-1 conv 3 -> Cout, reading the NCHW
// input the caller supplies (models/models.py:336: input.view(-1, 3, H, W)) and writing the NHWC raw
// activation + batch statistics the fused chain works on; and its weight gradient.
// Reference: conv_bn(3, 32, 2), archs/mobilenet_v2.py:7-12,90.
#include "rowop.cuh"

namespace ehgr {

template <typename X>
__device__ __forceinline__ float ld_x(const X* p);
template <>
__device__ __forceinline__ float ld_x<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_x<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

struct StemGeom {
  int nt, h, w, cout, ho, wo;
  long long n_out;
};

// block = (Cout/V, P); thread = V output channels of one output pixel
template <typename X, typename T, int V>
__global__ void __launch_bounds__(256)
stem_fwd_kernel(const X* __restrict__ x, const float* __restrict__ wgt, T* __restrict__ out,
                double* __restrict__ stats, StemGeom g) {
  extern __shared__ float smem[];
  float* ws = smem;                     // [27][Cout]
  float* s_sum = smem + 27 * g.cout;    // [2*Cout]
  const int nthreads = blockDim.x * blockDim.y;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < 27 * g.cout; i += nthreads) {
    const int tap = i / g.cout, co = i - tap * g.cout;  // tap = ci*9 + kh*3 + kw
    ws[i] = wgt[co * 27 + tap];
  }
  if (stats) for (int i = tid; i < 2 * g.cout; i += nthreads) s_sum[i] = 0.f;
  __syncthreads();
  const int c0 = threadIdx.x * V;
  const long long q = static_cast<long long>(blockIdx.x) * blockDim.y + threadIdx.y;
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  const bool live = q < g.n_out;
  if (live) {
    const int wo = static_cast<int>(q % g.wo);
    const long long r = q / g.wo;
    const int ho = static_cast<int>(r % g.ho);
    const long long nt = r / g.ho;
    const size_t plane = static_cast<size_t>(g.h) * g.w;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const X* xp = x + (static_cast<size_t>(nt) * 3 + ci) * plane;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int hi = ho * 2 + kh - 1;
        if (hi < 0 || hi >= g.h) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int wi = wo * 2 + kw - 1;
          if (wi < 0 || wi >= g.w) continue;
          const float xv = ld_x<X>(xp + static_cast<size_t>(hi) * g.w + wi);
          float wv[V];
          load_vec<float, V>(ws + (ci * 9 + kh * 3 + kw) * g.cout + c0, wv);
#pragma unroll
          for (int i = 0; i < V; ++i) acc[i] = fmaf(xv, wv[i], acc[i]);
        }
      }
    }
    store_vec<T, V>(out + q * g.cout + c0, acc);
  }
  if (stats) {
    if (live) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        atomicAdd(&s_sum[c0 + i], acc[i]);
        atomicAdd(&s_sum[g.cout + c0 + i], acc[i] * acc[i]);
      }
    }
    __syncthreads();
    for (int i = tid; i < 2 * g.cout; i += nthreads) atomicAdd(&stats[i], static_cast<double>(s_sum[i]));
  }
}

// dw[co][ci][kh][kw] += sum_q dy[q][co] * x[nt][ci][2ho+kh-1][2wo+kw-1];  block = (Cout/4, P)
template <typename X, typename T>
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(RowOp dy, const X* __restrict__ x, float* __restrict__ dwgt, StemGeom g, int iters) {
  extern __shared__ float s_acc[];  // [27][Cout]
  const int nthreads = blockDim.x * blockDim.y;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < 27 * g.cout; i += nthreads) s_acc[i] = 0.f;
  __syncthreads();
  const int c0 = threadIdx.x * 4;
  float acc[27][4];
#pragma unroll
  for (int t = 0; t < 27; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[t][i] = 0.f;
  const long long base = static_cast<long long>(blockIdx.x) * (static_cast<long long>(blockDim.y) * iters) + threadIdx.y;
  const size_t plane = static_cast<size_t>(g.h) * g.w;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const long long q = base + static_cast<long long>(it) * blockDim.y;
    if (q >= g.n_out) break;
    const int wo = static_cast<int>(q % g.wo);
    const long long r = q / g.wo;
    const int ho = static_cast<int>(r % g.ho);
    const long long nt = r / g.ho;
    float d[4];
    load_row<T, 4>(dy, q, c0, g.cout, d);
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const X* xp = x + (static_cast<size_t>(nt) * 3 + ci) * plane;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int hi = ho * 2 + kh - 1;
        if (hi < 0 || hi >= g.h) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int wi = wo * 2 + kw - 1;
          if (wi < 0 || wi >= g.w) continue;
          const float xv = ld_x<X>(xp + static_cast<size_t>(hi) * g.w + wi);
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[ci * 9 + kh * 3 + kw][i] = fmaf(d[i], xv, acc[ci * 9 + kh * 3 + kw][i]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 27; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) atomicAdd(&s_acc[t * g.cout + c0 + i], acc[t][i]);
  __syncthreads();
  for (int i = tid; i < 27 * g.cout; i += nthreads) {
    const int tap = i / g.cout, co = i - tap * g.cout;
    atomicAdd(&dwgt[co * 27 + tap], s_acc[i]);
  }
}

static int stem_geom(StemGeom& g, int nt, int h, int w, int cout) {
  if (nt < 0 || h <= 0 || w <= 0 || cout <= 0 || (cout % 8) || cout > 256) return EHGR_E_SHAPE;
  g.nt = nt; g.h = h; g.w = w; g.cout = cout;
  g.ho = (h - 1) / 2 + 1;
  g.wo = (w - 1) / 2 + 1;
  g.n_out = static_cast<long long>(nt) * g.ho * g.wo;
  return EHGR_OK;
}

template <typename X>
static int stem_fwd_launch(const void* x, const float* w, void* out, double* stats, const StemGeom& g, int dtype,
                           cudaStream_t s) {
  const size_t smem = static_cast<size_t>(29) * g.cout * sizeof(float);
  if (dtype == EHGR_F32) {
    const dim3 block(g.cout / 4, std::max(1, 256 / (g.cout / 4)));
    const long long blocks = cdiv(g.n_out, block.y);
    stem_fwd_kernel<X, float, 4><<<static_cast<unsigned>(blocks), block, smem, s>>>(
        static_cast<const X*>(x), w, static_cast<float*>(out), stats, g);
  } else {
    const dim3 block(g.cout / 8, std::max(1, 256 / (g.cout / 8)));
    const long long blocks = cdiv(g.n_out, block.y);
    stem_fwd_kernel<X, __nv_bfloat16, 8><<<static_cast<unsigned>(blocks), block, smem, s>>>(
        static_cast<const X*>(x), w, static_cast<__nv_bfloat16*>(out), stats, g);
  }
  return launch_status();
}

template <typename X>
static int stem_wgrad_launch(const RowOp& dy, const void* x, float* dw, const StemGeom& g, int dtype,
                             cudaStream_t s) {
  const dim3 block(g.cout / 4, std::max(1, 256 / (g.cout / 4)));
  long long iters = cdiv(g.n_out, 4LL * kNumSMs * block.y);
  iters = std::max(1LL, std::min(iters, 4096LL));
  const long long blocks = cdiv(g.n_out, static_cast<long long>(block.y) * iters);
  const size_t smem = static_cast<size_t>(27) * g.cout * sizeof(float);
  if (dtype == EHGR_F32)
    stem_wgrad_kernel<X, float><<<static_cast<unsigned>(blocks), block, smem, s>>>(dy, static_cast<const X*>(x), dw, g,
                                                                                  static_cast<int>(iters));
  else
    stem_wgrad_kernel<X, __nv_bfloat16><<<static_cast<unsigned>(blocks), block, smem, s>>>(
        dy, static_cast<const X*>(x), dw, g, static_cast<int>(iters));
  return launch_status();
}

}  // namespace ehgr

using namespace ehgr;

extern "C" int ehgr_stem_fwd(const void* x, const float* w, void* out, double* stats, int nt, int h, int wd,
                             int cout, int x_dtype, int dtype, ehgr_stream_t stream) {
  StemGeom g;
  if (esize_of(x_dtype) == 0 || esize_of(dtype) == 0) return EHGR_E_DTYPE;
  if (!x || !w || !out) return EHGR_E_NULL;
  if (int st = stem_geom(g, nt, h, wd, cout)) return st;
  if (!aligned_to(out, 16) || !aligned_to(x, esize_of(x_dtype))) return EHGR_E_ALIGN;
  if (g.n_out == 0) return EHGR_OK;
  cudaStream_t s = as_stream(stream);
  return x_dtype == EHGR_F32 ? stem_fwd_launch<float>(x, w, out, stats, g, dtype, s)
                             : stem_fwd_launch<__nv_bfloat16>(x, w, out, stats, g, dtype, s);
}

extern "C" int ehgr_stem_wgrad(const ehgr_rowop* dy, const void* x, float* dw, int nt, int h, int wd, int cout,
                               int x_dtype, int dtype, ehgr_stream_t stream) {
  StemGeom g;
  if (esize_of(x_dtype) == 0 || esize_of(dtype) == 0) return EHGR_E_DTYPE;
  if (!x || !dw) return EHGR_E_NULL;
  if (int st = validate_rowop(dy, esize_of(dtype))) return st;
  if (int st = stem_geom(g, nt, h, wd, cout)) return st;
  if (g.n_out == 0) return EHGR_OK;
  cudaStream_t s = as_stream(stream);
  return x_dtype == EHGR_F32 ? stem_wgrad_launch<float>(*dy, x, dw, g, dtype, s)
                             : stem_wgrad_launch<__nv_bfloat16>(*dy, x, dw, g, dtype, s);
}
