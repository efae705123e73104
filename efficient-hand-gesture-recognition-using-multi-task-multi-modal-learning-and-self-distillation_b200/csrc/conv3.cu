// conv3.cu — N2: the pieces of the MTMM depth decoder (models/models_MTMM.py:129-155) around the implicit-GEMM
// convolution.  The dense 3x3 convolutions themselves run on the pointwise-GEMM kernels through the CONV3 row
// operand (rowop.cuh, pw_tc.cu, pw_tc_wgrad.cu: im2col by TMA boxes for a finished activation, or a cp.async gather
// with the producer's BatchNorm+ReLU and the nearest x2 upsample folded into the load); this file holds
//   * the weight layouts those GEMMs read (conv3_pack) and the inverse for the weight gradient (conv3_unpack_grad),
//   * nn.Upsample(scale_factor=2, mode='nearest') materialised (upsample2_fwd: the TMA gather cannot fold the x2 index
//     map) and its adjoint (upsample2_bwd: 2x2 block sums),
//   * the depth head Conv2d(32, 1, 1, bias) + Sigmoid, forward and backward, as one streaming pass each.
#include "rowop.cuh"

namespace ehgr {

template <typename T>
__device__ __forceinline__ void put(T* p, float v);
template <>
__device__ __forceinline__ void put<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void put<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// thread = one weight element w[n][c][tap]; the two scattered stores are into L2-resident arrays of a few MB
template <typename T>
__global__ void __launch_bounds__(256)
conv3_pack_kernel(const float* __restrict__ w, T* __restrict__ wf, T* __restrict__ wd, int cout, int cin) {
  const long long total = 9LL * cout * cin;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
    const int tap = static_cast<int>(i % 9);
    const long long nc = i / 9;
    const int c = static_cast<int>(nc % cin), n = static_cast<int>(nc / cin);
    const float v = w[i];
    if (wf) put<T>(wf + (static_cast<long long>(n) * 9 + tap) * cin + c, v);
    if (wd) put<T>(wd + (static_cast<long long>(c) * 9 + (8 - tap)) * cout + n, v);
  }
}

// dw[n][c][tap] += dwp[n][tap*cin + c]
__global__ void __launch_bounds__(256)
conv3_unpack_grad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int cout, int cin) {
  const long long total = 9LL * cout * cin;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
    const int tap = static_cast<int>(i % 9);
    const long long nc = i / 9;
    const int c = static_cast<int>(nc % cin), n = static_cast<int>(nc / cin);
    dw[i] += dwp[(static_cast<long long>(n) * 9 + tap) * cin + c];
  }
}

// g[f,h,w,c] = sum over the 2x2 block of g_up; thread = one 16-byte channel vector of one output pixel
template <typename T>
__global__ void __launch_bounds__(256)
upsample2_bwd_kernel(const T* __restrict__ gu, T* __restrict__ g, long long frames, int h, int w, int c) {
  constexpr int V = VecOf<T>::N;
  const int cv = c / V;
  const long long total = frames * h * w * cv;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
    const int v = static_cast<int>(i % cv);
    long long pix = i / cv;
    const int x = static_cast<int>(pix % w);
    pix /= w;
    const int y = static_cast<int>(pix % h);
    const long long f = pix / h;
    const T* src = gu + ((f * 2 * h + 2 * y) * 2 * w + 2 * x) * c + v * V;
    float a[V], b[V], acc[V];
    load_vec<T, V>(src, a);
    load_vec<T, V>(src + c, b);
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = a[k] + b[k];
    load_vec<T, V>(src + 2LL * w * c, a);
    load_vec<T, V>(src + 2LL * w * c + c, b);
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] += a[k] + b[k];
    store_vec<T, V>(g + ((f * h + y) * w + x) * c + v * V, acc);
  }
}

// out[f, 2h+a, 2w+b, c] = in[f, h, w, c]; thread = one 16-byte channel vector of one OUTPUT pixel (coalesced stores)
template <typename T>
__global__ void __launch_bounds__(256)
upsample2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long frames, int h, int w, int c) {
  constexpr int V = VecOf<T>::N;
  const int cv = c / V;
  const long long total = frames * 4 * h * w * cv;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
    const int v = static_cast<int>(i % cv);
    long long pix = i / cv;
    const int xo = static_cast<int>(pix % (2 * w));
    pix /= 2 * w;
    const int yo = static_cast<int>(pix % (2 * h));
    const long long f = pix / (2 * h);
    const uint4 val = *reinterpret_cast<const uint4*>(x + ((f * h + (yo >> 1)) * w + (xo >> 1)) * c + v * V);
    *reinterpret_cast<uint4*>(y + i * V) = val;
  }
}

// Depth head.  A group of C/V lanes owns a row (C = 32, bf16: 4 lanes; fp32: 8 lanes): each lane loads one 16-byte
// vector through the row operand (the last decoder BatchNorm+ReLU is applied here), partial dot products are
// combined with shuffles inside the group.
template <typename T, bool kBwd>
__global__ void __launch_bounds__(256)
depth_head_kernel(RowOp a, const float* __restrict__ wv, const float* __restrict__ bias, float* __restrict__ out,
                  const float* __restrict__ dout, T* __restrict__ g_a, float* __restrict__ dw, float* __restrict__ dbias,
                  long long M, int C) {
  constexpr int V = VecOf<T>::N;
  const int lanes = C / V;                         // power of two <= 32 (host-checked)
  const int lane = threadIdx.x & 31;
  const int sub = lane & (lanes - 1), grp = lane / lanes, groups = 32 / lanes;
  float wreg[V];
  load_vec<float, V>(wv + sub * V, wreg);
  const float b0 = bias ? bias[0] : 0.f;
  float dw_acc[V], db_acc = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) dw_acc[k] = 0.f;
  const long long warp_global = (blockIdx.x * 256LL + threadIdx.x) >> 5;
  const long long n_warps = (gridDim.x * 256LL) >> 5;
  for (long long m0 = warp_global * groups; m0 < M; m0 += n_warps * groups) {
    const long long m = m0 + grp;
    const bool live = m < M;
    float v[V];
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = 0.f;
    if (live) load_row<T, V, false>(a, m, sub * V, C, v);
    if (!kBwd) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < V; ++k) s = fmaf(v[k], wreg[k], s);
      for (int o = lanes >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (live && sub == 0) out[m] = 1.f / (1.f + __expf(-(s + b0)));
    } else {
      float dz = 0.f;
      if (live) {
        const float y = out[m];
        dz = dout[m] * y * (1.f - y);
      }
      float gv[V];
#pragma unroll
      for (int k = 0; k < V; ++k) {
        gv[k] = dz * wreg[k];
        dw_acc[k] = fmaf(dz, v[k], dw_acc[k]);
      }
      if (live) store_vec<T, V>(g_a + m * C + sub * V, gv);
      if (sub == 0) db_acc += dz;
    }
  }
  if (kBwd) {
    // lanes with equal `sub` hold partial sums of the same channels: reduce over the groups, then one atomic per channel
    for (int o = 16; o >= lanes; o >>= 1) {
#pragma unroll
      for (int k = 0; k < V; ++k) dw_acc[k] += __shfl_xor_sync(0xffffffffu, dw_acc[k], o);
      db_acc += __shfl_xor_sync(0xffffffffu, db_acc, o);
    }
    __shared__ float s_dw[256], s_db;
    if (threadIdx.x < 256) s_dw[threadIdx.x] = 0.f;
    if (threadIdx.x == 0) s_db = 0.f;
    __syncthreads();
    if (grp == 0) {
#pragma unroll
      for (int k = 0; k < V; ++k) atomicAdd(&s_dw[sub * V + k], dw_acc[k]);
      if (sub == 0) atomicAdd(&s_db, db_acc);
    }
    __syncthreads();
    if (threadIdx.x < C) atomicAdd(&dw[threadIdx.x], s_dw[threadIdx.x]);
    if (threadIdx.x == 0 && dbias) atomicAdd(dbias, s_db);
  }
}

}  // namespace ehgr

using namespace ehgr;

extern "C" int ehgr_conv3_pack(const float* w, void* wf, void* wd, int cout, int cin, int dtype, ehgr_stream_t stream) {
  if (esize_of(dtype) == 0) return EHGR_E_DTYPE;
  if (!w) return EHGR_E_NULL;
  if (cout <= 0 || cin <= 0) return EHGR_E_SHAPE;
  if (!wf && !wd) return EHGR_OK;
  const long long total = 9LL * cout * cin;
  const unsigned grid = static_cast<unsigned>(std::min<long long>(cdiv(total, 256), 8LL * kNumSMs));
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32)
    conv3_pack_kernel<float><<<grid, 256, 0, s>>>(w, static_cast<float*>(wf), static_cast<float*>(wd), cout, cin);
  else
    conv3_pack_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(w, static_cast<__nv_bfloat16*>(wf), static_cast<__nv_bfloat16*>(wd),
                                                          cout, cin);
  return launch_status();
}

extern "C" int ehgr_conv3_unpack_grad(const float* dwp, float* dw, int cout, int cin, ehgr_stream_t stream) {
  if (!dwp || !dw) return EHGR_E_NULL;
  if (cout <= 0 || cin <= 0) return EHGR_E_SHAPE;
  const long long total = 9LL * cout * cin;
  const unsigned grid = static_cast<unsigned>(std::min<long long>(cdiv(total, 256), 8LL * kNumSMs));
  conv3_unpack_grad_kernel<<<grid, 256, 0, as_stream(stream)>>>(dwp, dw, cout, cin);
  return launch_status();
}

extern "C" int ehgr_upsample2_bwd(const void* g_up, void* g, long long frames, int h, int w, int c, int dtype,
                                  ehgr_stream_t stream) {
  const int es = esize_of(dtype);
  if (es == 0) return EHGR_E_DTYPE;
  if (!g_up || !g) return EHGR_E_NULL;
  if (frames < 0 || h <= 0 || w <= 0 || c <= 0 || (c % (16 / es))) return EHGR_E_SHAPE;
  if (!aligned_to(g_up, 16) || !aligned_to(g, 16)) return EHGR_E_ALIGN;
  if (frames == 0) return EHGR_OK;
  const long long total = frames * h * w * (c / (16 / es));
  const unsigned grid = static_cast<unsigned>(std::min<long long>(cdiv(total, 256), 16LL * kNumSMs));
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32)
    upsample2_bwd_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(g_up), static_cast<float*>(g), frames, h, w, c);
  else
    upsample2_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(g_up),
                                                             static_cast<__nv_bfloat16*>(g), frames, h, w, c);
  return launch_status();
}

extern "C" int ehgr_upsample2_fwd(const void* x, void* y, long long frames, int h, int w, int c, int dtype,
                                  ehgr_stream_t stream) {
  const int es = esize_of(dtype);
  if (es == 0) return EHGR_E_DTYPE;
  if (!x || !y) return EHGR_E_NULL;
  if (frames < 0 || h <= 0 || w <= 0 || c <= 0 || (c % (16 / es))) return EHGR_E_SHAPE;
  if (!aligned_to(x, 16) || !aligned_to(y, 16)) return EHGR_E_ALIGN;
  if (frames == 0) return EHGR_OK;
  const long long total = frames * 4 * h * w * (c / (16 / es));
  const unsigned grid = static_cast<unsigned>(std::min<long long>(cdiv(total, 256), 16LL * kNumSMs));
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32)
    upsample2_fwd_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(x), static_cast<float*>(y), frames, h, w, c);
  else
    upsample2_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y),
                                                             frames, h, w, c);
  return launch_status();
}

static int depth_head_check(const ehgr_rowop* a, const float* w, long long m, int c, int dtype) {
  const int es = esize_of(dtype);
  if (es == 0) return EHGR_E_DTYPE;
  if (!w) return EHGR_E_NULL;
  if (int st = validate_rowop_nogate(a, es)) return st;
  const int v = 16 / es;
  if (m < 0 || c <= 0 || (c % v) || c > 256) return EHGR_E_SHAPE;
  const int lanes = c / v;
  if (lanes > 32 || (lanes & (lanes - 1))) return EHGR_E_SHAPE;     // a row is owned by a power-of-two lane group
  return EHGR_OK;
}

extern "C" int ehgr_depth_head_fwd(const ehgr_rowop* a, const float* w, const float* bias, float* out, long long m, int c,
                                   int dtype, ehgr_stream_t stream) {
  if (int st = depth_head_check(a, w, m, c, dtype)) return st;
  if (!out) return EHGR_E_NULL;
  if (m == 0) return EHGR_OK;
  const int groups = 32 / (c / (16 / esize_of(dtype)));
  const unsigned grid = static_cast<unsigned>(std::min<long long>(cdiv(m, 8LL * groups), 8LL * kNumSMs));
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32)
    depth_head_kernel<float, false><<<grid, 256, 0, s>>>(*a, w, bias, out, nullptr, nullptr, nullptr, nullptr, m, c);
  else
    depth_head_kernel<__nv_bfloat16, false><<<grid, 256, 0, s>>>(*a, w, bias, out, nullptr, nullptr, nullptr, nullptr, m, c);
  return launch_status();
}

extern "C" int ehgr_depth_head_bwd(const ehgr_rowop* a, const float* w, const float* out, const float* dout, void* g_a,
                                   float* dw, float* dbias, long long m, int c, int dtype, ehgr_stream_t stream) {
  if (int st = depth_head_check(a, w, m, c, dtype)) return st;
  if (!out || !dout || !g_a || !dw) return EHGR_E_NULL;
  if (!aligned_to(g_a, 16)) return EHGR_E_ALIGN;
  if (m == 0) return EHGR_OK;
  const int groups = 32 / (c / (16 / esize_of(dtype)));
  const unsigned grid = static_cast<unsigned>(std::min<long long>(cdiv(m, 8LL * groups), 4LL * kNumSMs));
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32)
    depth_head_kernel<float, true><<<grid, 256, 0, s>>>(*a, w, nullptr, const_cast<float*>(out), dout, static_cast<float*>(g_a),
                                                       dw, dbias, m, c);
  else
    depth_head_kernel<__nv_bfloat16, true><<<grid, 256, 0, s>>>(*a, w, nullptr, const_cast<float*>(out), dout,
                                                               static_cast<__nv_bfloat16*>(g_a), dw, dbias, m, c);
  return launch_status();
}
