// resnet.cu — N3: the pieces of the torchvision ResNet bottleneck path that the MobileNetV2 kernels do not cover.
//
// The reference builds `torchvision.models.resnet50` for base_model='resnet50' (models/models.py:108-117) and wraps
// `conv1` of every bottleneck with a temporal module (models/temporal_shift.py:101-146).  Inside a bottleneck the 1x1
// convolutions run on the pointwise GEMM kernels (PLAIN / AFFINE / SHIFT row operands) and the 3x3 convolution on the
// same kernels through the CONV3 operand (csrc/pw_tc.cu, rowop.cuh).  What is left and lives here:
//   * stem7_fwd / stem7_wgrad : Conv2d(3, 64, 7, stride 2, padding 3) from the NCHW input, BatchNorm statistics in the
//                               forward epilogue (torchvision ResNet.conv1);
//   * maxpool3_fwd / _bwd     : MaxPool2d(3, stride 2, padding 1) over the lazy BatchNorm+ReLU of the stem output; the
//                               forward keeps the arg-max tap (one byte per element) so the backward is a gather without atomics;
//   * subsample2_fwd / _bwd   : x[:, ::2, ::2] and its adjoint (zero insertion).  A stride-2 convolution with
//                               kernel k and padding (k-1)/2 equals the stride-1 convolution sampled at even pixels
//                               (3x3 of layer{2,3,4}.0.conv2), and a stride-2 1x1 convolution equals the 1x1 convolution
//                               of the subsampled input (the downsample branch);
//   * bn_add_relu / relu_bwd  : out = relu(BN(raw) + identity) at the end of a bottleneck and the mask of its backward.
// All of them are streaming, HBM-bound passes over NHWC tensors (16-byte channel vectors).
#include "rowop.cuh"

namespace ehgr {

template <typename TX>
__device__ __forceinline__ float ld_in(const TX* p);
template <>
__device__ __forceinline__ float ld_in<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_in<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

constexpr int kStem7C = 64;       // torchvision ResNet: inplanes = 64
constexpr int kStem7Taps = 147;   // 3 * 7 * 7

// thread = 8 output channels of one output pixel; the filter sits in shared memory as [tap][cout]
template <typename TX, typename T>
__global__ void __launch_bounds__(256)
stem7_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ w, T* __restrict__ out, double* __restrict__ stats,
                 long long frames, int H, int W, int Ho, int Wo) {
  __shared__ float s_w[kStem7Taps * kStem7C];
  __shared__ float s_stat[2 * kStem7C];
  const int tid = threadIdx.x;
  for (int i = tid; i < kStem7Taps * kStem7C; i += 256) {
    const int co = i / kStem7Taps, t = i - co * kStem7Taps;      // w is [cout][147]
    s_w[t * kStem7C + co] = w[i];
  }
  if (tid < 2 * kStem7C) s_stat[tid] = 0.f;
  __syncthreads();
  const int cg = tid & 7, pl = tid >> 3;                         // 8 channel groups x 32 pixels
  const long long total = frames * Ho * Wo;
  float ssum[8], ssq[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) ssum[k] = ssq[k] = 0.f;
  for (long long pix = blockIdx.x * 32LL + pl; pix < total; pix += 32LL * gridDim.x) {
    const int ow = static_cast<int>(pix % Wo);
    const long long r = pix / Wo;
    const int oh = static_cast<int>(r % Ho);
    const long long f = r / Ho;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int ci = 0; ci < 3; ++ci) {
      const TX* xp = x + (f * 3 + ci) * static_cast<long long>(H) * W;
      for (int ky = 0; ky < 7; ++ky) {
        const int ih = 2 * oh - 3 + ky;
        if (static_cast<unsigned>(ih) >= static_cast<unsigned>(H)) continue;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const int iw = 2 * ow - 3 + kx;
          if (static_cast<unsigned>(iw) >= static_cast<unsigned>(W)) continue;
          const float xv = ld_in<TX>(xp + static_cast<long long>(ih) * W + iw);
          const float* wr = s_w + ((ci * 7 + ky) * 7 + kx) * kStem7C + cg * 8;
          const float4 w0 = *reinterpret_cast<const float4*>(wr);
          const float4 w1 = *reinterpret_cast<const float4*>(wr + 4);
          acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]);
          acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
          acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
          acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc[k] = round_to<T>(acc[k]);                              // statistics of the stored values
      ssum[k] += acc[k];
      ssq[k] = fmaf(acc[k], acc[k], ssq[k]);
    }
    store_vec<T, 8>(out + pix * kStem7C + cg * 8, acc);
  }
  if (stats) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(&s_stat[cg * 8 + k], ssum[k]);
      atomicAdd(&s_stat[kStem7C + cg * 8 + k], ssq[k]);
    }
    __syncthreads();
    if (tid < 2 * kStem7C) atomicAdd(&stats[tid], static_cast<double>(s_stat[tid]));
  }
}

// dw[co][tap] += sum over output pixels of rowop(dy)[pix][co] * x[patch(pix)][tap].  A CTA stages 32 output pixels (their
// d(raw) rows and their 147-value input patches) in shared memory; thread (co, tg) owns taps tg, tg+4, ... in registers.
template <typename TX, typename T>
__global__ void __launch_bounds__(256)
stem7_wgrad_kernel(RowOp dy, const TX* __restrict__ x, float* __restrict__ dw, long long frames, int H, int W, int Ho, int Wo) {
  constexpr int P = 32, V = VecOf<T>::N, CV = kStem7C / V, XS = kStem7Taps + 1;
  __shared__ float s_d[P * kStem7C];
  __shared__ float s_x[P * XS];
  const int tid = threadIdx.x;
  const int co = tid & 63, tg = tid >> 6;
  const long long total = frames * Ho * Wo;
  float acc[37];
#pragma unroll
  for (int j = 0; j < 37; ++j) acc[j] = 0.f;
  for (long long base = blockIdx.x * static_cast<long long>(P); base < total; base += static_cast<long long>(P) * gridDim.x) {
    for (int i = tid; i < P * CV; i += 256) {
      const int p = i / CV, cv = i - p * CV;
      float v[V];
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = 0.f;
      if (base + p < total) load_row<T, V, false>(dy, base + p, cv * V, kStem7C, v);
#pragma unroll
      for (int k = 0; k < V; ++k) s_d[p * kStem7C + cv * V + k] = v[k];
    }
    for (int i = tid; i < P * XS; i += 256) {
      const int p = i / XS, t = i - p * XS;
      float xv = 0.f;
      const long long pix = base + p;
      if (t < kStem7Taps && pix < total) {
        const int ow = static_cast<int>(pix % Wo);
        const long long r = pix / Wo;
        const int oh = static_cast<int>(r % Ho);
        const long long f = r / Ho;
        const int ci = t / 49, rem = t - ci * 49, ky = rem / 7, kx = rem - ky * 7;
        const int ih = 2 * oh - 3 + ky, iw = 2 * ow - 3 + kx;
        if (static_cast<unsigned>(ih) < static_cast<unsigned>(H) && static_cast<unsigned>(iw) < static_cast<unsigned>(W))
          xv = ld_in<TX>(x + ((f * 3 + ci) * static_cast<long long>(H) + ih) * W + iw);
      }
      s_x[i] = xv;
    }
    __syncthreads();
    for (int p = 0; p < P; ++p) {
      const float d = s_d[p * kStem7C + co];
      const float* xr = s_x + p * XS + tg;
#pragma unroll
      for (int j = 0; j < 37; ++j) acc[j] = fmaf(d, xr[4 * j], acc[j]);   // tap 147 is the zero pad column
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 37; ++j) {
    const int t = tg + 4 * j;
    if (t < kStem7Taps) atomicAdd(&dw[co * kStem7Taps + t], acc[j]);
  }
}

// im2col of the 7x7/2 stem for the tensor-core GEMM: a[pix][t] = x[f, ci, 2oh-3+ky, 2ow-3+kx] with t = (ci*7 + ky)*7 + kx,
// zero outside the image and for the pad columns t >= 147.  thread = 8 consecutive columns of one output pixel.
template <typename TX, typename T>
__global__ void __launch_bounds__(256)
stem7_im2col_kernel(const TX* __restrict__ x, T* __restrict__ a, long long frames, int H, int W, int Ho, int Wo, int KP) {
  const int kvn = KP / 8;
  const long long total = frames * Ho * Wo * kvn;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
    const int kv = static_cast<int>(i % kvn);
    long long pix = i / kvn;
    const int ow = static_cast<int>(pix % Wo);
    pix /= Wo;
    const int oh = static_cast<int>(pix % Ho);
    const long long f = pix / Ho;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int t = kv * 8 + k;
      float xv = 0.f;
      if (t < kStem7Taps) {
        const int ci = t / 49, rem = t - ci * 49, ky = rem / 7, kx = rem - ky * 7;
        const int ih = 2 * oh - 3 + ky, iw = 2 * ow - 3 + kx;
        if (static_cast<unsigned>(ih) < static_cast<unsigned>(H) && static_cast<unsigned>(iw) < static_cast<unsigned>(W))
          xv = ld_in<TX>(x + ((f * 3 + ci) * static_cast<long long>(H) + ih) * W + iw);
      }
      v[k] = xv;
    }
    store_vec<T, 8>(a + i * 8, v);
  }
}

// wp[co][t] = w[co][t] for t < 147, 0 for the pad columns (fp32 or bf16)
template <typename T>
__global__ void __launch_bounds__(256)
stem7_pack_kernel(const float* __restrict__ w, T* __restrict__ wp, int cout, int KP) {
  const int total = cout * KP;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += 256 * gridDim.x) {
    const int co = i / KP, t = i - co * KP;
    const float v = t < kStem7Taps ? w[co * kStem7Taps + t] : 0.f;
    if constexpr (sizeof(T) == 4) wp[i] = v; else wp[i] = __float2bfloat16_rn(v);
  }
}

// dw[co][t] += dwp[co][t], t < 147
__global__ void __launch_bounds__(256)
stem7_unpack_grad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int cout, int KP) {
  const int total = cout * kStem7Taps;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += 256 * gridDim.x) {
    const int co = i / kStem7Taps, t = i - co * kStem7Taps;
    dw[i] += dwp[co * KP + t];
  }
}

// y[f,oh,ow,c] = max over the 3x3 window (stride 2, padding 1; padding never wins) of rowop(a); idx = winning tap 3*ky+kx
// (the first maximum in scan order, as at::max_pool2d).  thread = one 16-byte channel vector of one output pixel.
template <typename T>
__global__ void __launch_bounds__(256)
maxpool3_fwd_kernel(RowOp a, T* __restrict__ y, uint8_t* __restrict__ idx, long long frames, int H, int W, int Ho, int Wo, int C) {
  constexpr int V = VecOf<T>::N;
  const int cvn = C / V;
  const long long total = frames * Ho * Wo * cvn;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
    const int cv = static_cast<int>(i % cvn);
    long long pix = i / cvn;
    const int ow = static_cast<int>(pix % Wo);
    pix /= Wo;
    const int oh = static_cast<int>(pix % Ho);
    const long long f = pix / Ho;
    float best[V];
    int bi[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { best[k] = -INFINITY; bi[k] = 0; }
    for (int ky = 0; ky < 3; ++ky) {
      const int ih = 2 * oh - 1 + ky;
      if (static_cast<unsigned>(ih) >= static_cast<unsigned>(H)) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int iw = 2 * ow - 1 + kx;
        if (static_cast<unsigned>(iw) >= static_cast<unsigned>(W)) continue;
        float v[V];
        load_row<T, V, false>(a, (f * H + ih) * W + iw, cv * V, C, v);
#pragma unroll
        for (int k = 0; k < V; ++k)
          if (v[k] > best[k] || v[k] != v[k]) { best[k] = v[k]; bi[k] = ky * 3 + kx; }
      }
    }
    store_vec<T, V>(y + i * V, best);
    uint8_t* ip = idx + i * V;
    if constexpr (V == 8) {
      const uint32_t lo = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      const uint32_t hi = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      *reinterpret_cast<uint2*>(ip) = make_uint2(lo, hi);
    } else {
      *reinterpret_cast<uint32_t*>(ip) = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
    }
  }
}

// gx[f,ih,iw,c] = sum of g[f,oh,ow,c] over the (at most four) windows that contain (ih,iw) and whose arg-max is that pixel
template <typename T>
__global__ void __launch_bounds__(256)
maxpool3_bwd_kernel(const T* __restrict__ g, const uint8_t* __restrict__ idx, T* __restrict__ gx, long long frames, int H, int W,
                    int Ho, int Wo, int C) {
  constexpr int V = VecOf<T>::N;
  const int cvn = C / V;
  const long long total = frames * H * W * cvn;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
    const int cv = static_cast<int>(i % cvn);
    long long pix = i / cvn;
    const int iw = static_cast<int>(pix % W);
    pix /= W;
    const int ih = static_cast<int>(pix % H);
    const long long f = pix / H;
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    const int oh_hi = min((ih + 1) >> 1, Ho - 1), ow_hi = min((iw + 1) >> 1, Wo - 1);
    for (int oh = ih >> 1; oh <= oh_hi; ++oh) {
      const int ky = ih - 2 * oh + 1;
      for (int ow = iw >> 1; ow <= ow_hi; ++ow) {
        const int tap = ky * 3 + (iw - 2 * ow + 1);
        const long long o = ((f * Ho + oh) * Wo + ow) * C + cv * V;
        float gv[V];
        load_vec<T, V>(g + o, gv);
        alignas(8) uint8_t b[V];
        if constexpr (V == 8) *reinterpret_cast<uint2*>(b) = *reinterpret_cast<const uint2*>(idx + o);
        else *reinterpret_cast<uint32_t*>(b) = *reinterpret_cast<const uint32_t*>(idx + o);
#pragma unroll
        for (int k = 0; k < V; ++k)
          if (b[k] == tap) acc[k] += gv[k];
      }
    }
    store_vec<T, V>(gx + i * V, acc);
  }
}

// y[f,yo,xo,:] = x[f,2yo,2xo,:] (+ per-channel sum / sum of squares of y).  block = (channel vectors, rows): a thread keeps
// one channel vector and strides over output pixels, so the statistics stay in registers until the end.
template <typename T>
__global__ void __launch_bounds__(256)
subsample2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, double* __restrict__ stats, long long frames, int H, int W,
                      int Ho, int Wo, int C) {
  constexpr int V = VecOf<T>::N;
  extern __shared__ float s_sub[];                   // [2C] when stats
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthreads = blockDim.x * blockDim.y;
  if (stats) {
    for (int i = tid; i < 2 * C; i += nthreads) s_sub[i] = 0.f;
    __syncthreads();
  }
  const int cvn = C / V;
  const long long total = frames * Ho * Wo;
  const long long row_stride = static_cast<long long>(gridDim.x) * blockDim.y;
  for (int cv = threadIdx.x; cv < cvn; cv += blockDim.x) {
    float ssum[V], ssq[V];
#pragma unroll
    for (int k = 0; k < V; ++k) ssum[k] = ssq[k] = 0.f;
    for (long long m = static_cast<long long>(blockIdx.x) * blockDim.y + threadIdx.y; m < total; m += row_stride) {
      const int xo = static_cast<int>(m % Wo);
      const long long r = m / Wo;
      const int yo = static_cast<int>(r % Ho);
      const long long f = r / Ho;
      float v[V];
      load_vec<T, V>(x + ((f * H + 2 * yo) * W + 2 * xo) * C + cv * V, v);
      store_vec<T, V>(y + m * C + cv * V, v);
#pragma unroll
      for (int k = 0; k < V; ++k) { ssum[k] += v[k]; ssq[k] = fmaf(v[k], v[k], ssq[k]); }
    }
    if (stats) {
#pragma unroll
      for (int k = 0; k < V; ++k) { atomicAdd(&s_sub[cv * V + k], ssum[k]); atomicAdd(&s_sub[C + cv * V + k], ssq[k]); }
    }
  }
  if (stats) {
    __syncthreads();
    for (int i = tid; i < 2 * C; i += nthreads) atomicAdd(&stats[i], static_cast<double>(s_sub[i]));
  }
}

// gx[f,h,w,:] = g[f,h/2,w/2,:] when h and w are even, else 0; thread = one 16-byte vector of the full-resolution tensor
template <typename T>
__global__ void __launch_bounds__(256)
subsample2_bwd_kernel(const T* __restrict__ g, T* __restrict__ gx, long long frames, int H, int W, int Ho, int Wo, int C) {
  constexpr int V = VecOf<T>::N;
  const int cvn = C / V;
  const long long total = frames * H * W * cvn;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
    const int cv = static_cast<int>(i % cvn);
    long long pix = i / cvn;
    const int w = static_cast<int>(pix % W);
    pix /= W;
    const int h = static_cast<int>(pix % H);
    const long long f = pix / H;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (!((h | w) & 1)) val = *reinterpret_cast<const uint4*>(g + ((f * Ho + (h >> 1)) * Wo + (w >> 1)) * C + cv * V);
    *reinterpret_cast<uint4*>(gx + i * V) = val;
  }
}

// out = relu(raw*scale + shift + addend)
template <typename T>
__global__ void __launch_bounds__(256)
bn_add_relu_kernel(const T* __restrict__ raw, const float* __restrict__ scale, const float* __restrict__ shift,
                   const T* __restrict__ addend, T* __restrict__ out, long long M, int C) {
  constexpr int V = VecOf<T>::N;
  const int cvn = C / V;
  const long long total = M * cvn;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
    const int c0 = static_cast<int>(i % cvn) * V;
    float v[V], a[V], s[V], b[V];
    load_vec<T, V>(raw + i * V, v);
    load_vec<T, V>(addend + i * V, a);
    load_vec<float, V>(scale + c0, s);
    load_vec<float, V>(shift + c0, b);
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = fmaxf(fmaf(v[k], s[k], b[k]) + a[k], 0.f);
    store_vec<T, V>(out + i * V, v);
  }
}

// gz = g where out > 0, else 0
template <typename T>
__global__ void __launch_bounds__(256)
relu_bwd_kernel(const T* __restrict__ g, const T* __restrict__ out, T* __restrict__ gz, long long nvec) {
  constexpr int V = VecOf<T>::N;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < nvec; i += 256LL * gridDim.x) {
    float gv[V], ov[V];
    load_vec<T, V>(g + i * V, gv);
    load_vec<T, V>(out + i * V, ov);
#pragma unroll
    for (int k = 0; k < V; ++k) gv[k] = ov[k] > 0.f ? gv[k] : 0.f;
    store_vec<T, V>(gz + i * V, gv);
  }
}

static unsigned stream_grid(long long threads) {
  return static_cast<unsigned>(std::max<long long>(1, std::min<long long>(cdiv(threads, 256), 16LL * kNumSMs)));
}

static int half_up(int v) { return (v - 1) / 2 + 1; }   // output extent of a stride-2 window with padding (k-1)/2

}  // namespace ehgr

using namespace ehgr;

extern "C" int ehgr_stem7_fwd(const void* x, const float* w, void* out, double* stats, long long frames, int h, int w_in,
                              int cout, int x_dtype, int out_dtype, ehgr_stream_t stream) {
  if (esize_of(x_dtype) == 0 || esize_of(out_dtype) == 0) return EHGR_E_DTYPE;
  if (!x || !w || !out) return EHGR_E_NULL;
  if (frames < 0 || h <= 0 || w_in <= 0) return EHGR_E_SHAPE;
  if (cout != kStem7C) return EHGR_E_UNSUPPORTED;
  if (!aligned_to(out, 16) || !aligned_to(x, esize_of(x_dtype)) || (stats && !aligned_to(stats, 8))) return EHGR_E_ALIGN;
  if (frames == 0) return EHGR_OK;
  const int ho = half_up(h), wo = half_up(w_in);
  const long long total = frames * ho * wo;
  const unsigned grid = static_cast<unsigned>(std::max<long long>(1, std::min<long long>(cdiv(total, 32), 8LL * kNumSMs)));
  cudaStream_t s = as_stream(stream);
#define EHGR_S7(TX, T) stem7_fwd_kernel<TX, T><<<grid, 256, 0, s>>>(static_cast<const TX*>(x), w, static_cast<T*>(out), stats, frames, h, w_in, ho, wo)
  if (x_dtype == EHGR_F32) { if (out_dtype == EHGR_F32) EHGR_S7(float, float); else EHGR_S7(float, __nv_bfloat16); }
  else { if (out_dtype == EHGR_F32) EHGR_S7(__nv_bfloat16, float); else EHGR_S7(__nv_bfloat16, __nv_bfloat16); }
#undef EHGR_S7
  return launch_status();
}

extern "C" int ehgr_stem7_wgrad(const ehgr_rowop* dy, const void* x, float* dw, long long frames, int h, int w_in, int cout,
                                int x_dtype, int dtype, ehgr_stream_t stream) {
  const int es = esize_of(dtype);
  if (es == 0 || esize_of(x_dtype) == 0) return EHGR_E_DTYPE;
  if (!x || !dw) return EHGR_E_NULL;
  if (int st = validate_rowop_nogate(dy, es)) return st;
  if (dy->mode == EHGR_ROW_SHIFT) return EHGR_E_UNSUPPORTED;
  if (frames < 0 || h <= 0 || w_in <= 0) return EHGR_E_SHAPE;
  if (cout != kStem7C) return EHGR_E_UNSUPPORTED;
  if (frames == 0) return EHGR_OK;
  const int ho = half_up(h), wo = half_up(w_in);
  const long long total = frames * ho * wo;
  const unsigned grid = static_cast<unsigned>(std::max<long long>(1, std::min<long long>(cdiv(total, 32), 2LL * kNumSMs)));
  cudaStream_t s = as_stream(stream);
#define EHGR_S7W(TX, T) stem7_wgrad_kernel<TX, T><<<grid, 256, 0, s>>>(*dy, static_cast<const TX*>(x), dw, frames, h, w_in, ho, wo)
  if (x_dtype == EHGR_F32) { if (dtype == EHGR_F32) EHGR_S7W(float, float); else EHGR_S7W(float, __nv_bfloat16); }
  else { if (dtype == EHGR_F32) EHGR_S7W(__nv_bfloat16, float); else EHGR_S7W(__nv_bfloat16, __nv_bfloat16); }
#undef EHGR_S7W
  return launch_status();
}

extern "C" int ehgr_stem7_im2col(const void* x, void* a, long long frames, int h, int w_in, int kp, int x_dtype, int dtype,
                                 ehgr_stream_t stream) {
  if (esize_of(x_dtype) == 0 || esize_of(dtype) == 0) return EHGR_E_DTYPE;
  if (!x || !a) return EHGR_E_NULL;
  if (frames < 0 || h <= 0 || w_in <= 0 || kp < kStem7Taps || (kp % 8)) return EHGR_E_SHAPE;
  if (!aligned_to(a, 16) || !aligned_to(x, esize_of(x_dtype))) return EHGR_E_ALIGN;
  if (frames == 0) return EHGR_OK;
  const int ho = half_up(h), wo = half_up(w_in);
  const unsigned grid = stream_grid(frames * ho * wo * (kp / 8));
  cudaStream_t s = as_stream(stream);
#define EHGR_S7I(TX, T) stem7_im2col_kernel<TX, T><<<grid, 256, 0, s>>>(static_cast<const TX*>(x), static_cast<T*>(a), frames, h, w_in, ho, wo, kp)
  if (x_dtype == EHGR_F32) { if (dtype == EHGR_F32) EHGR_S7I(float, float); else EHGR_S7I(float, __nv_bfloat16); }
  else { if (dtype == EHGR_F32) EHGR_S7I(__nv_bfloat16, float); else EHGR_S7I(__nv_bfloat16, __nv_bfloat16); }
#undef EHGR_S7I
  return launch_status();
}

extern "C" int ehgr_stem7_pack(const float* w, void* wp, int cout, int kp, int dtype, ehgr_stream_t stream) {
  if (esize_of(dtype) == 0) return EHGR_E_DTYPE;
  if (!w || !wp) return EHGR_E_NULL;
  if (cout <= 0 || kp < kStem7Taps || (kp % 8)) return EHGR_E_SHAPE;
  const unsigned grid = stream_grid(static_cast<long long>(cout) * kp);
  if (dtype == EHGR_F32) stem7_pack_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(w, static_cast<float*>(wp), cout, kp);
  else stem7_pack_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(w, static_cast<__nv_bfloat16*>(wp), cout, kp);
  return launch_status();
}

extern "C" int ehgr_stem7_unpack_grad(const float* dwp, float* dw, int cout, int kp, ehgr_stream_t stream) {
  if (!dwp || !dw) return EHGR_E_NULL;
  if (cout <= 0 || kp < kStem7Taps || (kp % 8)) return EHGR_E_SHAPE;
  stem7_unpack_grad_kernel<<<stream_grid(static_cast<long long>(cout) * kStem7Taps), 256, 0, as_stream(stream)>>>(dwp, dw, cout, kp);
  return launch_status();
}

static int pool_like_check(const void* a, const void* b, long long frames, int h, int w, int c, int dtype) {
  const int es = esize_of(dtype);
  if (es == 0) return EHGR_E_DTYPE;
  if (!a || !b) return EHGR_E_NULL;
  if (frames < 0 || h <= 0 || w <= 0 || c <= 0 || (c % (16 / es))) return EHGR_E_SHAPE;
  if (!aligned_to(a, 16) || !aligned_to(b, 16)) return EHGR_E_ALIGN;
  return EHGR_OK;
}

extern "C" int ehgr_maxpool3_fwd(const ehgr_rowop* a, void* y, void* idx, long long frames, int h, int w, int c, int dtype,
                                 ehgr_stream_t stream) {
  if (int st = pool_like_check(y, idx, frames, h, w, c, dtype)) return st;
  if (int st = validate_rowop_nogate(a, esize_of(dtype))) return st;
  if (a->mode == EHGR_ROW_SHIFT) return EHGR_E_UNSUPPORTED;
  if (frames == 0) return EHGR_OK;
  const int ho = half_up(h), wo = half_up(w);
  const unsigned grid = stream_grid(frames * ho * wo * (c / (16 / esize_of(dtype))));
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32)
    maxpool3_fwd_kernel<float><<<grid, 256, 0, s>>>(*a, static_cast<float*>(y), static_cast<uint8_t*>(idx), frames, h, w, ho, wo, c);
  else
    maxpool3_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(*a, static_cast<__nv_bfloat16*>(y), static_cast<uint8_t*>(idx), frames,
                                                            h, w, ho, wo, c);
  return launch_status();
}

extern "C" int ehgr_maxpool3_bwd(const void* g, const void* idx, void* gx, long long frames, int h, int w, int c, int dtype,
                                 ehgr_stream_t stream) {
  if (int st = pool_like_check(g, gx, frames, h, w, c, dtype)) return st;
  if (!idx) return EHGR_E_NULL;
  if (!aligned_to(idx, 16)) return EHGR_E_ALIGN;
  if (frames == 0) return EHGR_OK;
  const int ho = half_up(h), wo = half_up(w);
  const unsigned grid = stream_grid(frames * h * w * (c / (16 / esize_of(dtype))));
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32)
    maxpool3_bwd_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(g), static_cast<const uint8_t*>(idx),
                                                    static_cast<float*>(gx), frames, h, w, ho, wo, c);
  else
    maxpool3_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(g), static_cast<const uint8_t*>(idx),
                                                            static_cast<__nv_bfloat16*>(gx), frames, h, w, ho, wo, c);
  return launch_status();
}

extern "C" int ehgr_subsample2_fwd(const void* x, void* y, double* stats, long long frames, int h, int w, int c, int dtype,
                                   ehgr_stream_t stream) {
  if (int st = pool_like_check(x, y, frames, h, w, c, dtype)) return st;
  if (stats && !aligned_to(stats, 8)) return EHGR_E_ALIGN;
  if (frames == 0) return EHGR_OK;
  const int ho = half_up(h), wo = half_up(w);
  const int cv = c / (16 / esize_of(dtype));
  int bx = cv;
  while (bx > 256) bx = (bx + 1) / 2;
  const dim3 block(bx, std::max(1, 256 / bx));
  const long long rows = frames * ho * wo;
  const long long blocks = std::max(1LL, std::min(cdiv(rows, 4LL * block.y), 2LL * kNumSMs));
  const size_t smem = stats ? static_cast<size_t>(2) * c * sizeof(float) : 0;
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32)
    subsample2_fwd_kernel<float><<<static_cast<unsigned>(blocks), block, smem, s>>>(static_cast<const float*>(x), static_cast<float*>(y),
                                                                                   stats, frames, h, w, ho, wo, c);
  else
    subsample2_fwd_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), block, smem, s>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), stats, frames, h, w, ho, wo, c);
  return launch_status();
}

extern "C" int ehgr_subsample2_bwd(const void* g, void* gx, long long frames, int h, int w, int c, int dtype,
                                   ehgr_stream_t stream) {
  if (int st = pool_like_check(g, gx, frames, h, w, c, dtype)) return st;
  if (frames == 0) return EHGR_OK;
  const int ho = half_up(h), wo = half_up(w);
  const unsigned grid = stream_grid(frames * h * w * (c / (16 / esize_of(dtype))));
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32)
    subsample2_bwd_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(g), static_cast<float*>(gx), frames, h, w, ho, wo, c);
  else
    subsample2_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(g), static_cast<__nv_bfloat16*>(gx),
                                                              frames, h, w, ho, wo, c);
  return launch_status();
}

extern "C" int ehgr_bn_add_relu(const void* raw, const float* scale, const float* shift, const void* addend, void* out,
                                long long m, int c, int dtype, ehgr_stream_t stream) {
  const int es = esize_of(dtype);
  if (es == 0) return EHGR_E_DTYPE;
  if (!raw || !scale || !shift || !addend || !out) return EHGR_E_NULL;
  if (m < 0 || c <= 0 || (c % (16 / es))) return EHGR_E_SHAPE;
  if (!aligned_to(raw, 16) || !aligned_to(addend, 16) || !aligned_to(out, 16) || !aligned_to(scale, 16) || !aligned_to(shift, 16))
    return EHGR_E_ALIGN;
  if (m == 0) return EHGR_OK;
  const unsigned grid = stream_grid(m * (c / (16 / es)));
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32)
    bn_add_relu_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(raw), scale, shift, static_cast<const float*>(addend),
                                                   static_cast<float*>(out), m, c);
  else
    bn_add_relu_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(raw), scale, shift,
                                                           static_cast<const __nv_bfloat16*>(addend),
                                                           static_cast<__nv_bfloat16*>(out), m, c);
  return launch_status();
}

extern "C" int ehgr_relu_bwd(const void* g, const void* out, void* gz, long long n, int dtype, ehgr_stream_t stream) {
  const int es = esize_of(dtype);
  if (es == 0) return EHGR_E_DTYPE;
  if (!g || !out || !gz) return EHGR_E_NULL;
  if (n < 0 || (n % (16 / es))) return EHGR_E_SHAPE;
  if (!aligned_to(g, 16) || !aligned_to(out, 16) || !aligned_to(gz, 16)) return EHGR_E_ALIGN;
  if (n == 0) return EHGR_OK;
  const long long nvec = n / (16 / es);
  const unsigned grid = stream_grid(nvec);
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32)
    relu_bwd_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(g), static_cast<const float*>(out), static_cast<float*>(gz), nvec);
  else
    relu_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(g), static_cast<const __nv_bfloat16*>(out),
                                                        static_cast<__nv_bfloat16*>(gz), nvec);
  return launch_status();
}
