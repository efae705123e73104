// head.cu — K10 classifier head: global average pool (x.mean(3).mean(2), archs/mobilenet_v2.py:112),
// new_fc and the segment consensus (models/models.py:341-356, models/basic_ops.py:9-37).  Because
// the consensus is a mean over segments and new_fc is linear, logits = W * mean_t(feat) + b: the mean
// over T is taken first (one pass over [N*T, F]) and the tiny GEMV runs once per clip.
#include "rowop.cuh"

namespace ehgr {

// one block per frame; thread -> channel vectors, loop over the hw rows of the frame
template <typename T>
__global__ void __launch_bounds__(256)
pool_fwd_kernel(RowOp a, float* __restrict__ pooled, int hw, int C) {
  constexpr int V = VecOf<T>::N;
  const long long nt = blockIdx.x;
  const float inv = 1.f / static_cast<float>(hw);
  for (int cv = threadIdx.x; cv < C / V; cv += blockDim.x) {
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    for (int p = 0; p < hw; ++p) {
      float v[V];
      load_row<T, V>(a, nt * hw + p, cv * V, C, v);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += v[i];
    }
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] *= inv;
    store_vec<float, V>(pooled + nt * C + cv * V, acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const float* __restrict__ dpooled, T* __restrict__ da, int hw, int C) {
  constexpr int V = VecOf<T>::N;
  const long long nt = blockIdx.x;
  const float inv = 1.f / static_cast<float>(hw);
  for (int cv = threadIdx.x; cv < C / V; cv += blockDim.x) {
    float g[V];
    load_vec<float, V>(dpooled + nt * C + cv * V, g);
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] *= inv;
    for (int p = 0; p < hw; ++p) store_vec<T, V>(da + (nt * hw + p) * C + cv * V, g);
  }
}

// one block per clip
__global__ void __launch_bounds__(256)
fc_consensus_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ w, const float* __restrict__ bias,
                        float* __restrict__ meanfeat, float* __restrict__ logits, int T, int F, int K) {
  extern __shared__ float mf[];  // [F]
  const long long n = blockIdx.x;
  const float inv = 1.f / static_cast<float>(T);
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += feat[(n * T + t) * F + f];
    s *= inv;
    mf[f] = s;
    meanfeat[n * F + f] = s;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int k = warp; k < K; k += nwarps) {
    float s = 0.f;
    for (int f = lane; f < F; f += 32) s = fmaf(w[static_cast<size_t>(k) * F + f], mf[f], s);
    s = warp_sum(s);
    if (lane == 0) logits[n * K + k] = s + (bias ? bias[k] : 0.f);
  }
}

// dfeat: one block per clip, thread per feature
__global__ void __launch_bounds__(256)
fc_consensus_dfeat_kernel(const float* __restrict__ dlogits, const float* __restrict__ w, float* __restrict__ dfeat,
                          int T, int F, int K) {
  extern __shared__ float dl[];  // [K]
  const long long n = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) dl[k] = dlogits[n * K + k];
  __syncthreads();
  const float inv = 1.f / static_cast<float>(T);
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s = fmaf(dl[k], w[static_cast<size_t>(k) * F + f], s);
    s *= inv;
    for (int t = 0; t < T; ++t) dfeat[(n * T + t) * F + f] = s;
  }
}

// dW[k][f] += sum_n dl[n][k] * mf[n][f]; grid = (ceil(F/256), K); db from blockIdx.x == 0
__global__ void __launch_bounds__(256)
fc_consensus_dw_kernel(const float* __restrict__ dlogits, const float* __restrict__ meanfeat, float* __restrict__ dw,
                       float* __restrict__ dbias, int N, int F, int K) {
  const int k = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < F) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(dlogits[static_cast<size_t>(n) * K + k], meanfeat[static_cast<size_t>(n) * F + f], s);
    atomicAdd(&dw[static_cast<size_t>(k) * F + f], s);
  }
  if (dbias && blockIdx.x == 0 && threadIdx.x == 0) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += dlogits[static_cast<size_t>(n) * K + k];
    atomicAdd(&dbias[k], s);
  }
}

}  // namespace ehgr

using namespace ehgr;

extern "C" int ehgr_pool_fwd(const ehgr_rowop* a, float* pooled, int nt, int hw, int c, int dtype,
                             ehgr_stream_t stream) {
  const int es = esize_of(dtype);
  if (es == 0) return EHGR_E_DTYPE;
  if (!pooled) return EHGR_E_NULL;
  if (int st = validate_rowop(a, es)) return st;
  if (nt < 0 || hw <= 0 || c <= 0 || (c % (16 / es))) return EHGR_E_SHAPE;
  if (nt == 0) return EHGR_OK;
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32) pool_fwd_kernel<float><<<nt, 256, 0, s>>>(*a, pooled, hw, c);
  else pool_fwd_kernel<__nv_bfloat16><<<nt, 256, 0, s>>>(*a, pooled, hw, c);
  return launch_status();
}

extern "C" int ehgr_pool_bwd(const float* dpooled, void* da, int nt, int hw, int c, int dtype,
                             ehgr_stream_t stream) {
  const int es = esize_of(dtype);
  if (es == 0) return EHGR_E_DTYPE;
  if (!dpooled || !da) return EHGR_E_NULL;
  if (nt < 0 || hw <= 0 || c <= 0 || (c % (16 / es))) return EHGR_E_SHAPE;
  if (nt == 0) return EHGR_OK;
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32) pool_bwd_kernel<float><<<nt, 256, 0, s>>>(dpooled, static_cast<float*>(da), hw, c);
  else pool_bwd_kernel<__nv_bfloat16><<<nt, 256, 0, s>>>(dpooled, static_cast<__nv_bfloat16*>(da), hw, c);
  return launch_status();
}

extern "C" int ehgr_fc_consensus_fwd(const float* feat, const float* w, const float* bias, float* meanfeat,
                                     float* logits, int n, int n_segment, int f, int k, ehgr_stream_t stream) {
  if (!feat || !w || !meanfeat || !logits) return EHGR_E_NULL;
  if (n < 0 || n_segment <= 0 || f <= 0 || k <= 0 || f > 12000) return EHGR_E_SHAPE;
  if (n == 0) return EHGR_OK;
  fc_consensus_fwd_kernel<<<n, 256, static_cast<size_t>(f) * sizeof(float), as_stream(stream)>>>(
      feat, w, bias, meanfeat, logits, n_segment, f, k);
  return launch_status();
}

extern "C" int ehgr_fc_consensus_bwd(const float* dlogits, const float* meanfeat, const float* w, float* dfeat,
                                     float* dw, float* dbias, int n, int n_segment, int f, int k,
                                     ehgr_stream_t stream) {
  if (!dlogits || !meanfeat || !w) return EHGR_E_NULL;
  if (n < 0 || n_segment <= 0 || f <= 0 || k <= 0 || k > 12000) return EHGR_E_SHAPE;
  if (n == 0) return EHGR_OK;
  cudaStream_t s = as_stream(stream);
  int st = EHGR_OK;
  if (dfeat) {
    fc_consensus_dfeat_kernel<<<n, 256, static_cast<size_t>(k) * sizeof(float), s>>>(dlogits, w, dfeat, n_segment, f, k);
    st = launch_status();
    if (st) return st;
  }
  if (dw) {
    dim3 grid(static_cast<unsigned>(cdiv(f, 256)), k);
    fc_consensus_dw_kernel<<<grid, 256, 0, s>>>(dlogits, meanfeat, dw, dbias, n, f, k);
    st = launch_status();
  }
  return st;
}
