// tpool.cu — TemporalPool.temporal_pool (models/temporal_shift.py:89-98; duplicate models/action.py:165-176):
//   x.view(n, T, c, h, w).transpose(1, 2) -> max_pool3d(kernel (3,1,1), stride (2,1,1), padding (1,0,0)) -> back,
// i.e. out[n, t', :] = max over the frames {2t'-1, 2t', 2t'+1} that exist, T' = (T - 1) / 2 + 1.
// The pooling never mixes positions inside a frame, so NCHW and NHWC are the same problem: [n, T, S] with
// S = c*h*w contiguous elements per frame, one 16-byte vector per thread step.  Backward routes the gradient to the
// FIRST maximum of a window (strict '>' scan, as max_pool3d does), recomputed from x: every input element looks at
// the (at most two) windows it belongs to — no atomics, no index tensor.
#include "rowop.cuh"

namespace ehgr {

// 16-byte vectors when a frame is a whole number of them, single elements otherwise (V = 1)
template <typename T, int V>
__device__ __forceinline__ void ldv(const T* __restrict__ p, float (&v)[V]) {
  if constexpr (V == 1) v[0] = static_cast<float>(*p); else load_vec<T, V>(p, v);
}
template <typename T, int V>
__device__ __forceinline__ void stv(T* __restrict__ p, const float (&v)[V]) {
  if constexpr (V == 1) *p = static_cast<T>(v[0]); else store_vec<T, V>(p, v);
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
tpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, long long n, int t_in, int t_out, long long vecs_per_frame) {
  const long long total = n * t_out * vecs_per_frame;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long v = i % vecs_per_frame, r = i / vecs_per_frame;
    const int to = static_cast<int>(r % t_out);
    const long long b = r / t_out;
    float m[V];
#pragma unroll
    for (int k = 0; k < V; ++k) m[k] = -INFINITY;
#pragma unroll
    for (int d = -1; d <= 1; ++d) {
      const int t = 2 * to + d;
      if (t < 0 || t >= t_in) continue;
      float a[V];
      ldv<T, V>(x + ((b * t_in + t) * vecs_per_frame + v) * V, a);
#pragma unroll
      for (int k = 0; k < V; ++k) m[k] = (a[k] > m[k] || a[k] != a[k]) ? a[k] : m[k];
    }
    stv<T, V>(out + i * V, m);
  }
}

// index (frame) of the first maximum of window `to`
template <typename T, int V>
__device__ __forceinline__ void window_argmax(const T* __restrict__ x, long long b, int to, int t_in, long long vpf, long long v,
                                              int (&arg)[V]) {
  float m[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { m[k] = -INFINITY; arg[k] = -1; }
#pragma unroll
  for (int d = -1; d <= 1; ++d) {
    const int t = 2 * to + d;
    if (t < 0 || t >= t_in) continue;
    float a[V];
    ldv<T, V>(x + ((b * t_in + t) * vpf + v) * V, a);
#pragma unroll
    for (int k = 0; k < V; ++k)
      if (a[k] > m[k] || a[k] != a[k] || arg[k] < 0) { m[k] = a[k]; arg[k] = t; }
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
tpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ g, T* __restrict__ dx, long long n, int t_in, int t_out,
                 long long vecs_per_frame) {
  const long long total = n * t_in * vecs_per_frame;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long v = i % vecs_per_frame, r = i / vecs_per_frame;
    const int t = static_cast<int>(r % t_in);
    const long long b = r / t_in;
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    // windows containing frame t: t' with |2t' - t| <= 1
    const int lo = t / 2, hi = (t + 1) / 2;          // even t: one window (t/2); odd t: (t-1)/2 and (t+1)/2
    for (int to = lo; to <= hi; ++to) {
      if (to >= t_out) continue;
      int arg[V];
      window_argmax<T, V>(x, b, to, t_in, vecs_per_frame, v, arg);
      float gv[V];
      ldv<T, V>(g + ((b * t_out + to) * vecs_per_frame + v) * V, gv);
#pragma unroll
      for (int k = 0; k < V; ++k)
        if (arg[k] == t) acc[k] += gv[k];
    }
    stv<T, V>(dx + i * V, acc);
  }
}

}  // namespace ehgr

using namespace ehgr;

static int tpool_check(const void* x, const void* out, long long n, int t_in, long long frame_elems, int dtype, int* V) {
  const int es = esize_of(dtype);
  if (es == 0) return EHGR_E_DTYPE;
  if (!x || !out) return EHGR_E_NULL;
  if (n < 0 || t_in <= 0 || frame_elems <= 0) return EHGR_E_SHAPE;
  *V = (frame_elems % (16 / es) == 0 && aligned_to(x, 16) && aligned_to(out, 16)) ? 16 / es : 1;
  if (!aligned_to(x, es) || !aligned_to(out, es)) return EHGR_E_ALIGN;
  return EHGR_OK;
}

template <typename T, int V>
static void tpool_launch(bool bwd, const void* x, const void* g, void* out, long long n, int t_in, int t_out, long long vpf,
                         cudaStream_t s) {
  const long long total = n * (bwd ? t_in : t_out) * vpf;
  const unsigned blocks = static_cast<unsigned>(std::min(cdiv(total, 256), 16LL * kNumSMs));
  if (bwd)
    tpool_bwd_kernel<T, V><<<blocks, 256, 0, s>>>(static_cast<const T*>(x), static_cast<const T*>(g), static_cast<T*>(out), n, t_in,
                                                  t_out, vpf);
  else
    tpool_fwd_kernel<T, V><<<blocks, 256, 0, s>>>(static_cast<const T*>(x), static_cast<T*>(out), n, t_in, t_out, vpf);
}

static int tpool_go(bool bwd, const void* x, const void* g, void* out, long long n, int t_in, long long frame_elems, int dtype,
                    ehgr_stream_t stream) {
  int V = 0;
  if (int st = tpool_check(x, out, n, t_in, frame_elems, dtype, &V)) return st;
  if (bwd) {
    if (!g) return EHGR_E_NULL;
    if (V > 1 && !aligned_to(g, 16)) V = 1;
  }
  const int t_out = (t_in - 1) / 2 + 1;
  const long long vpf = frame_elems / V;
  if (n * vpf == 0) return EHGR_OK;
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32) {
    if (V == 4) tpool_launch<float, 4>(bwd, x, g, out, n, t_in, t_out, vpf, s); else tpool_launch<float, 1>(bwd, x, g, out, n, t_in, t_out, vpf, s);
  } else {
    if (V == 8) tpool_launch<__nv_bfloat16, 8>(bwd, x, g, out, n, t_in, t_out, vpf, s);
    else tpool_launch<__nv_bfloat16, 1>(bwd, x, g, out, n, t_in, t_out, vpf, s);
  }
  return launch_status();
}

extern "C" int ehgr_temporal_pool_fwd(const void* x, void* out, long long n, int t_in, long long frame_elems, int dtype,
                                      ehgr_stream_t stream) {
  return tpool_go(false, x, nullptr, out, n, t_in, frame_elems, dtype, stream);
}

extern "C" int ehgr_temporal_pool_bwd(const void* x, const void* g, void* dx, long long n, int t_in, long long frame_elems,
                                      int dtype, ehgr_stream_t stream) {
  return tpool_go(true, x, g, dx, n, t_in, frame_elems, dtype, stream);
}
