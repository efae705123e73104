// pw_simt.cu — K8 pointwise-conv GEMM and its weight gradient on the fp32 CUDA-core path.
//
// This is the exact-fp32 engine (parity target 1e-5 against the reference) and the
// fallback for shapes the tcgen05 kernel (pw_tc.cu) does not take.  Same row-operand prologue
// (rowop.cuh) and the same epilogue contract as the tensor-core kernel: optional addend, optional
// per-channel batch statistics.
//   reference ops replaced: nn.Conv2d 1x1 forward / dgrad / wgrad of archs/mobilenet_v2.py:44-59.
#include "rowop.cuh"

namespace ehgr {

constexpr int GM = 64, GN = 64, GK = 16, GPAD = 4;

template <typename T>
__global__ void __launch_bounds__(256)
pw_gemm_simt_kernel(RowOp a, const float* __restrict__ w, int w_is_kn, T* __restrict__ out,
                    const T* __restrict__ addend, double* __restrict__ stats, long long M, int K, int N) {
  __shared__ __align__(16) float As[GK][GM + GPAD];
  __shared__ __align__(16) float Bs[GK][GN + GPAD];
  __shared__ float s_sum[GN], s_sq[GN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = static_cast<long long>(blockIdx.x) * GM;
  const int n0 = blockIdx.y * GN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  if (stats && tid < GN) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }

  const int a_row = tid >> 2, a_kq = tid & 3;
  for (int k0 = 0; k0 < K; k0 += GK) {
    {  // A tile (row operand)
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const long long m = m0 + a_row;
      const int k = k0 + a_kq * 4;
      if (m < M && k < K) load_row<T, 4>(a, m, k, K, v);
#pragma unroll
      for (int i = 0; i < 4; ++i) As[a_kq * 4 + i][a_row] = v[i];
    }
    if (!w_is_kn) {  // B[k][n] = w[n*K + k]
      const int n = n0 + (tid >> 2), k = k0 + (tid & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < N && k < K) v = *reinterpret_cast<const float4*>(w + static_cast<size_t>(n) * K + k);
      const int kk = (tid & 3) * 4, nn = tid >> 2;
      Bs[kk + 0][nn] = v.x; Bs[kk + 1][nn] = v.y; Bs[kk + 2][nn] = v.z; Bs[kk + 3][nn] = v.w;
    } else {         // B[k][n] = w[k*N + n]
      const int k = k0 + (tid >> 4), n = n0 + (tid & 15) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K && n < N) v = *reinterpret_cast<const float4*>(w + static_cast<size_t>(k) * N + n);
      *reinterpret_cast<float4*>(&Bs[tid >> 4][(tid & 15) * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int n = n0 + tx * 4;
  float csum[4] = {0.f, 0.f, 0.f, 0.f}, csq[4] = {0.f, 0.f, 0.f, 0.f};
  if (n < N) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long m = m0 + ty * 4 + i;
      if (m >= M) continue;
      float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
#pragma unroll
      for (int j = 0; j < 4; ++j) { csum[j] += v[j]; csq[j] = fmaf(v[j], v[j], csq[j]); }
      if (addend) {
        float ad[4];
        load_vec<T, 4>(addend + m * N + n, ad);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += ad[j];
      }
      store_vec<T, 4>(out + m * N + n, v);
    }
  }
  if (stats) {
    if (n < N) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&s_sum[tx * 4 + j], csum[j]);
        atomicAdd(&s_sq[tx * 4 + j], csq[j]);
      }
    }
    __syncthreads();
    if (tid < GN && n0 + tid < N) atomicAdd(&stats[n0 + tid], static_cast<double>(s_sum[tid]));
    else if (tid >= GN && tid < 2 * GN && n0 + tid - GN < N)
      atomicAdd(&stats[N + n0 + tid - GN], static_cast<double>(s_sq[tid - GN]));
  }
}

// dw[n][k] += sum_m dy[m][n] * a[m][k]
template <typename T>
__global__ void __launch_bounds__(256)
pw_wgrad_simt_kernel(RowOp dy, RowOp a, float* __restrict__ dw, long long M, int K, int N, int k_tiles,
                     long long rows_per_split) {
  __shared__ __align__(16) float Ds[GK][GN + GPAD];
  __shared__ __align__(16) float As[GK][GN + GPAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = (blockIdx.x / k_tiles) * GN, k0 = (blockIdx.x % k_tiles) * GN;
  const long long m_begin = static_cast<long long>(blockIdx.y) * rows_per_split;
  const long long m_end = min(M, m_begin + rows_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = tid >> 4, lq = (tid & 15) * 4;
  for (long long mb = m_begin; mb < m_end; mb += GK) {
    const long long m = mb + lr;
    float dv[4] = {0.f, 0.f, 0.f, 0.f}, av[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < m_end) {
      if (n0 + lq < N) load_row<T, 4>(dy, m, n0 + lq, N, dv);
      if (k0 + lq < K) load_row<T, 4>(a, m, k0 + lq, K, av);
    }
    *reinterpret_cast<float4*>(&Ds[lr][lq]) = make_float4(dv[0], dv[1], dv[2], dv[3]);
    *reinterpret_cast<float4*>(&As[lr][lq]) = make_float4(av[0], av[1], av[2], av[3]);
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < GK; ++mm) {
      const float4 d4 = *reinterpret_cast<const float4*>(&Ds[mm][ty * 4]);
      const float4 a4 = *reinterpret_cast<const float4*>(&As[mm][tx * 4]);
      const float dd[4] = {d4.x, d4.y, d4.z, d4.w}, aa[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dd[i], aa[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(dw + static_cast<size_t>(n) * K + k, acc[i][j]);
    }
  }
}

template <typename T>
static int pw_gemm_simt_launch(const RowOp& a, const float* w, int w_is_kn, void* out, const void* addend,
                               double* stats, long long M, int K, int N, cudaStream_t s) {
  const long long mt = cdiv(M, GM);
  if (mt > 0x7fffffffLL) return EHGR_E_SHAPE;
  dim3 grid(static_cast<unsigned>(mt), static_cast<unsigned>(cdiv(N, GN)));
  pw_gemm_simt_kernel<T><<<grid, 256, 0, s>>>(a, w, w_is_kn, static_cast<T*>(out), static_cast<const T*>(addend),
                                              stats, M, K, N);
  return launch_status();
}

int pw_gemm_simt(const RowOp& a, const float* w, int w_is_kn, void* out, const void* addend, double* stats,
                 long long M, int K, int N, int dtype, cudaStream_t s) {
  return dtype == EHGR_F32 ? pw_gemm_simt_launch<float>(a, w, w_is_kn, out, addend, stats, M, K, N, s)
                           : pw_gemm_simt_launch<__nv_bfloat16>(a, w, w_is_kn, out, addend, stats, M, K, N, s);
}

int pw_wgrad_simt(const RowOp& dy, const RowOp& a, float* dw, long long M, int K, int N, int dtype,
                  cudaStream_t s) {
  const int n_tiles = static_cast<int>(cdiv(N, GN)), k_tiles = static_cast<int>(cdiv(K, GN));
  const long long tiles = static_cast<long long>(n_tiles) * k_tiles;
  long long splits = cdiv(4LL * kNumSMs, tiles);
  splits = std::max(1LL, std::min(splits, cdiv(M, 256)));
  long long rows = cdiv(cdiv(M, splits), GK) * GK;
  splits = cdiv(M, rows);
  dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(splits));
  if (dtype == EHGR_F32)
    pw_wgrad_simt_kernel<float><<<grid, 256, 0, s>>>(dy, a, dw, M, K, N, k_tiles, rows);
  else
    pw_wgrad_simt_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(dy, a, dw, M, K, N, k_tiles, rows);
  return launch_status();
}

}  // namespace ehgr
