// stem.cu — K9: the network's first layer, dense 3x3 stride-2 pad-1 conv 3 -> Cout, reading the NCHW
// input the caller supplies (models/models.py:336: input.view(-1, 3, H, W)) and writing the NHWC raw
// activation + batch statistics the fused chain works on; and its weight gradient.
// Reference: conv_bn(3, 32, 2), archs/mobilenet_v2.py:7-12,90.
#include "tma.cuh"
#include "bnfin.cuh"

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

// ------------------------------------------------------------------------------------------------
// Fast path for Cout == 32 (every MobileNetV2 the reference builds: width_mult = 1).
//
// A CTA works on a band of R output rows of one frame.  The 2R+1 input rows of the three colour planes
// are staged into shared memory de-interleaved by column parity,
//     E[ci][r][j] = x[ci][hi0+r][2j]      O[ci][r][j] = x[ci][hi0+r][2j-1]   (zero outside the image)
// so that the stride-2 taps of neighbouring output columns are stride-1 (conflict-free) reads:
//     kw = 0 -> O[wo], kw = 1 -> E[wo], kw = 2 -> O[wo+1].
// ------------------------------------------------------------------------------------------------
constexpr int kStemC = 32;

template <typename X>
__device__ __forceinline__ void stem_stage_band(const X* __restrict__ x, float* __restrict__ planes, long long nt,
                                                int hi0, int nrows, int h, int w, int wp, int tid, int nthreads) {
  // planes: [3][nrows][2][wp] (E then O per row); index c2 in [0, 2wp): col = c2 - 1, c2 even -> O[c2/2], odd -> E[c2/2]
  const int per_row = 2 * wp;
  const size_t plane = static_cast<size_t>(h) * w;
  const int nwarps = nthreads >> 5, warp = tid >> 5, lane = tid & 31;
  for (int rr = warp; rr < 3 * nrows; rr += nwarps) {      // one warp per (plane, row): no per-element divisions
    const int ci = rr / nrows, r = rr - ci * nrows;
    const int hi = hi0 + r;
    const bool row_ok = hi >= 0 && hi < h;
    const X* src = x + (static_cast<size_t>(nt) * 3 + ci) * plane + static_cast<size_t>(row_ok ? hi : 0) * w;
    float* dst = planes + static_cast<size_t>(rr) * per_row;
#pragma unroll 4
    for (int c2 = lane; c2 < per_row; c2 += 32) {
      const int col = c2 - 1;
      float v = 0.f;
      if (row_ok && col >= 0 && col < w) v = ld_x<X>(src + col);
      dst[((c2 & 1) ? 0 : wp) + (c2 >> 1)] = v;
    }
  }
}

// 32 values per lane -> lane L receives the warp-wide sum of element L (31 shuffles)
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = lane & s;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

template <typename T>
__device__ __forceinline__ void stem_store32(T* __restrict__ p, const float (&a)[32]);
template <>
__device__ __forceinline__ void stem_store32<float>(float* __restrict__ p, const float (&a)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
}
template <>
__device__ __forceinline__ void stem_store32<__nv_bfloat16>(__nv_bfloat16* __restrict__ p, const float (&a)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    reinterpret_cast<uint4*>(p)[i] = make_uint4(pack_bf16x2(a[8 * i], a[8 * i + 1]), pack_bf16x2(a[8 * i + 2], a[8 * i + 3]),
                                                pack_bf16x2(a[8 * i + 4], a[8 * i + 5]), pack_bf16x2(a[8 * i + 6], a[8 * i + 7]));
}

// thread = one output column, two vertically adjacent output pixels x 32 channels in registers; the
// 27 x 32 weights are broadcast reads from shared memory, shared by the two pixels.
template <typename X, typename T>
__global__ void __launch_bounds__(128)
stem_fwd32_kernel(const X* __restrict__ x, const float* __restrict__ wgt, T* __restrict__ out,
                  double* __restrict__ stats, StemGeom g, int R, int bands, int wp) {
  extern __shared__ __align__(16) float smem[];
  float* ws = smem;                      // [27][32]
  float* s_stat = ws + 27 * kStemC;      // [64]
  float* planes = s_stat + 2 * kStemC;   // [3][2R+1][2][wp]
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 27 * kStemC; i += blockDim.x) {
    const int tap = i / kStemC, co = i - tap * kStemC;
    ws[i] = wgt[co * 27 + tap];
  }
  if (tid < 2 * kStemC) s_stat[tid] = 0.f;
  float st_sum = 0.f, st_sq = 0.f;       // this lane's channel (= lane)
  const int nrows = 2 * R + 1;
  const long long items = static_cast<long long>(g.nt) * bands;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const long long nt = item / bands;
    const int ho0 = static_cast<int>(item - nt * bands) * R;
    __syncthreads();
    stem_stage_band<X>(x, planes, nt, 2 * ho0 - 1, nrows, g.h, g.w, wp, tid, blockDim.x);
    __syncthreads();
    const int rmax = min(R, g.ho - ho0);
    const int wo_end = (g.wo + 31) / 32 * 32;          // whole warps run the shuffles
    for (int wo = tid; wo < wo_end; wo += blockDim.x) {
      const bool col_ok = wo < g.wo;
      const int wc = col_ok ? wo : 0;
      for (int oh = 0; oh < rmax; oh += 2) {
        const bool two = oh + 1 < rmax;
        float a0[32], a1[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) a0[i] = a1[i] = 0.f;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const float* row0 = planes + ((ci * nrows + 2 * oh + kh) * 2) * wp;   // E row, O row = +wp
            const float* row1 = row0 + 4 * wp;                                      // two input rows further down
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const int off = (kw == 1 ? 0 : wp) + wc + (kw == 2 ? 1 : 0);
              const float x0 = row0[off];
              const float x1 = two ? row1[off] : 0.f;
              const float4* wv = reinterpret_cast<const float4*>(ws + (ci * 9 + kh * 3 + kw) * kStemC);
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) {
                const float4 w4 = wv[q4];
                a0[4 * q4 + 0] = fmaf(x0, w4.x, a0[4 * q4 + 0]); a1[4 * q4 + 0] = fmaf(x1, w4.x, a1[4 * q4 + 0]);
                a0[4 * q4 + 1] = fmaf(x0, w4.y, a0[4 * q4 + 1]); a1[4 * q4 + 1] = fmaf(x1, w4.y, a1[4 * q4 + 1]);
                a0[4 * q4 + 2] = fmaf(x0, w4.z, a0[4 * q4 + 2]); a1[4 * q4 + 2] = fmaf(x1, w4.z, a1[4 * q4 + 2]);
                a0[4 * q4 + 3] = fmaf(x0, w4.w, a0[4 * q4 + 3]); a1[4 * q4 + 3] = fmaf(x1, w4.w, a1[4 * q4 + 3]);
              }
            }
          }
        }
        if (col_ok) {
          const long long q0 = (nt * g.ho + ho0 + oh) * g.wo + wo;
          stem_store32<T>(out + q0 * kStemC, a0);
          if (two) stem_store32<T>(out + (q0 + g.wo) * kStemC, a1);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) a0[i] = a1[i] = 0.f;
        }
        if (stats) {
          float sq[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) { sq[i] = fmaf(a0[i], a0[i], a1[i] * a1[i]); a0[i] += a1[i]; }
          st_sum += warp_transpose_sum32(a0, lane);
          st_sq += warp_transpose_sum32(sq, lane);
        }
      }
    }
  }
  if (stats) {
    atomicAdd(&s_stat[lane], st_sum);
    atomicAdd(&s_stat[kStemC + lane], st_sq);
    __syncthreads();
    if (tid < 2 * kStemC) atomicAdd(&stats[tid], static_cast<double>(s_stat[tid]));
  }
}

// ------------------------------------------------------------------------------------------------
// TMA variant of the banded forward kernel (row pitch and base 16-byte aligned, W + 2 <= 256): the band
// [3 planes][2R+1 rows][W+2 columns] is one 4-D TMA box (halo rows / columns zero-filled by hardware),
// double-buffered so that the next band is in flight while this one is computed; FP32 math on register
// pairs (FFMA2).  Shared-memory reads of the stride-2 taps are 2-way bank conflicted, which is noise
// next to the 8 weight vectors read per tap.
// ------------------------------------------------------------------------------------------------
template <typename X>
__device__ __forceinline__ float lds_x(const X* p);
template <>
__device__ __forceinline__ float lds_x<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float lds_x<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename X, typename T, int R>
__global__ void __launch_bounds__(128, 2)
stem_fwd32_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const float* __restrict__ wgt, T* __restrict__ out,
                      double* __restrict__ stats, StemGeom g, int bands, int bw, int tile_bytes, BnFin fin) {
  constexpr int NR = 2 * R + 1;
  constexpr int kPad = 16 / static_cast<int>(sizeof(X));      // elements in 16 bytes
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* ws = reinterpret_cast<float*>(smem_raw + 2 * tile_bytes);   // [27][32]
  float* s_stat = ws + 27 * kStemC;                                   // [64]
  const uint32_t bar0 = tc::smem_u32(s_stat + 2 * kStemC);
  const uint32_t tile0 = tc::smem_u32(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) {
    tc::mbar_init(bar0, 1);
    tc::mbar_init(bar0 + 8, 1);
    tc::fence_mbar_init();
    tma::prefetch_map(&tm_x);
  }
  for (int i = tid; i < 27 * kStemC; i += blockDim.x) {
    const int tap = i / kStemC, co = i - tap * kStemC;
    ws[i] = wgt[co * 27 + tap];
  }
  if (tid < 2 * kStemC) s_stat[tid] = 0.f;
  float st_sum = 0.f, st_sq = 0.f;
  const long long items = static_cast<long long>(g.nt) * bands;
  auto issue = [&](int buf, long long item) {
    const long long nt = item / bands;
    const int ho0 = static_cast<int>(item - nt * bands) * R;
    tma::expect_tx(bar0 + 8 * buf, static_cast<uint32_t>(3 * NR * bw * sizeof(X)));
    // the box starts one 16-byte unit left of the image (innermost TMA coordinates stay 16-byte aligned);
    // column -1 of the image is element kPad-1 of a tile row
    tma::load_4d(tile0 + buf * tile_bytes, &tm_x, bar0 + 8 * buf, -kPad, 2 * ho0 - 1, 0, static_cast<int>(nt));
  };
  __syncthreads();
  if (tid == 0 && blockIdx.x < items) issue(0, blockIdx.x);
  int buf = 0;
  uint32_t phase[2] = {0, 0};
  for (long long item = blockIdx.x; item < items; item += gridDim.x, buf ^= 1) {
    const long long nt = item / bands;
    const int ho0 = static_cast<int>(item - nt * bands) * R;
    if (tid == 0 && item + gridDim.x < items) issue(buf ^ 1, item + gridDim.x);
    tc::mbar_wait(bar0 + 8 * buf, phase[buf]);
    phase[buf] ^= 1;
    const X* tile = reinterpret_cast<const X*>(smem_raw + buf * tile_bytes);
    const int rmax = min(R, g.ho - ho0);
    const int wo_end = (g.wo + 31) / 32 * 32;          // whole warps run the shuffles
    for (int wo = tid; wo < wo_end; wo += blockDim.x) {
      const bool col_ok = wo < g.wo;
      const int wc = col_ok ? wo : 0;
      for (int oh = 0; oh < rmax; oh += 2) {
        const bool two = oh + 1 < rmax;
        float2 a0[16], a1[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a0[i] = a1[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const X* row0 = tile + (ci * NR + 2 * oh + kh) * bw + 2 * wc + (kPad - 1);   // box origin is column -kPad
            const X* row1 = row0 + 2 * bw;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const float x0 = lds_x<X>(row0 + kw);
              const float x1 = two ? lds_x<X>(row1 + kw) : 0.f;
              const float2 xx0 = make_float2(x0, x0), xx1 = make_float2(x1, x1);
              const float4* wv = reinterpret_cast<const float4*>(ws + (ci * 9 + kh * 3 + kw) * kStemC);
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) {
                const float4 w4 = wv[q4];
                const float2 wlo = make_float2(w4.x, w4.y), whi = make_float2(w4.z, w4.w);
                a0[2 * q4] = __ffma2_rn(xx0, wlo, a0[2 * q4]);
                a0[2 * q4 + 1] = __ffma2_rn(xx0, whi, a0[2 * q4 + 1]);
                a1[2 * q4] = __ffma2_rn(xx1, wlo, a1[2 * q4]);
                a1[2 * q4 + 1] = __ffma2_rn(xx1, whi, a1[2 * q4 + 1]);
              }
            }
          }
        }
        float f0[32], f1[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          f0[2 * i] = col_ok ? a0[i].x : 0.f; f0[2 * i + 1] = col_ok ? a0[i].y : 0.f;
          f1[2 * i] = col_ok ? a1[i].x : 0.f; f1[2 * i + 1] = col_ok ? a1[i].y : 0.f;
        }
        if (col_ok) {
          const long long q0 = (nt * g.ho + ho0 + oh) * g.wo + wo;
          stem_store32<T>(out + q0 * kStemC, f0);
          if (two) stem_store32<T>(out + (q0 + g.wo) * kStemC, f1);
        }
        if (stats) {
          float sq[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) { sq[i] = fmaf(f0[i], f0[i], f1[i] * f1[i]); f0[i] += f1[i]; }
          st_sum += warp_transpose_sum32(f0, lane);
          st_sq += warp_transpose_sum32(sq, lane);
        }
      }
    }
    tc::fence_proxy_async();
    __syncthreads();                                    // band consumed: its buffer may be refilled
  }
  if (stats) {
    atomicAdd(&s_stat[lane], st_sum);
    atomicAdd(&s_stat[kStemC + lane], st_sq);
    __syncthreads();
    if (tid < 2 * kStemC) atomicAdd(&stats[tid], static_cast<double>(s_stat[tid]));
  }
  bn_finalize_if_last(fin, stats, kStemC, stats != nullptr && tid < 2 * kStemC);
}

// Weight gradient.  12 warps = 4 channel groups (8 output channels) x 3 input planes; a warp keeps its
// 9 taps x 8 channels in registers over every band the CTA visits, lanes = output pixels of the band.
// rowop(dy) is evaluated once per element while the band is staged (as fp32, [quad][pixel][4]).
template <typename X, typename T>
__global__ void __launch_bounds__(384)
stem_wgrad32_kernel(RowOp dy, const X* __restrict__ x, float* __restrict__ dwgt, StemGeom g, int R, int bands, int wp) {
  constexpr int V = VecOf<T>::N;
  constexpr int QV = V / 4;                    // quads per vector
  extern __shared__ __align__(16) float smem[];
  const int nrows = 2 * R + 1;
  const int npix = R * g.wo;
  float* dtile = smem;                                       // [8][npix][4]  (16-byte aligned: float4 accesses)
  float* planes = dtile + 8 * npix * 4;                      // [3][nrows][2][wp]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cog = warp & 3, ci = warp >> 2;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
  const int vec_per_row = kStemC / V;
  const int cv = tid % vec_per_row;            // blockDim % vec_per_row == 0: fixed channel vector per thread
  const long long items = static_cast<long long>(g.nt) * bands;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const long long nt = item / bands;
    const int ho0 = static_cast<int>(item - nt * bands) * R;
    const int rmax = min(R, g.ho - ho0);
    const int np = rmax * g.wo;
    __syncthreads();
    stem_stage_band<X>(x, planes, nt, 2 * ho0 - 1, nrows, g.h, g.w, wp, tid, blockDim.x);
    {
      RowLoader<T, V> ld;            // coefficients live only while staging (the tap loop needs the registers)
      ld.init(dy, cv * V, kStemC);
      const long long q0 = (nt * g.ho + ho0) * g.wo;
      const int step = blockDim.x / vec_per_row;
#pragma unroll 1
      for (int p0 = tid / vec_per_row; p0 < np; p0 += 2 * step) {
        typename RowLoader<T, V>::Raw raw[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) if (p0 + j * step < np) raw[j] = ld.fetch(dy, q0 + p0 + j * step);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int p = p0 + j * step;
          if (p < np) {
            float v[V];
            ld.finish(dy, raw[j], v);
#pragma unroll
            for (int qd = 0; qd < QV; ++qd)
              *reinterpret_cast<float4*>(dtile + (static_cast<size_t>(cv * QV + qd) * npix + p) * 4) =
                  make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
          }
        }
      }
    }
    __syncthreads();
    for (int oh = 0; oh < rmax; ++oh) {
      for (int wo = lane; wo < g.wo; wo += 32) {
        const int p = oh * g.wo + wo;
        const float4 d0 = *reinterpret_cast<const float4*>(dtile + (static_cast<size_t>(cog * 2) * npix + p) * 4);
        const float4 d1 = *reinterpret_cast<const float4*>(dtile + (static_cast<size_t>(cog * 2 + 1) * npix + p) * 4);
        const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const float* row = planes + ((ci * nrows + 2 * oh + kh) * 2) * wp;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float xv = row[(kw == 1 ? 0 : wp) + wo + (kw == 2 ? 1 : 0)];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[kh * 3 + kw][i] = fmaf(d[i], xv, acc[kh * 3 + kw][i]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = acc[t][i];
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
      if (lane == 0) atomicAdd(&dwgt[(cog * 8 + i) * 27 + ci * 9 + t], v);
    }
}

// TMA variant of the weight-gradient kernel: the input band arrives as one TMA box while the CTA stages
// rowop(dy); the tap loop runs on register pairs (FFMA2).  Same warp roles as above.
template <typename X, typename T, int R>
__global__ void __launch_bounds__(384)
stem_wgrad32_tma_kernel(RowOp dy, const __grid_constant__ CUtensorMap tm_x, float* __restrict__ dwgt, StemGeom g, int bands,
                        int bw, int tile_bytes) {
  constexpr int V = VecOf<T>::N;
  constexpr int QV = V / 4;
  constexpr int NR = 2 * R + 1;
  constexpr int kPad = 16 / static_cast<int>(sizeof(X));
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int npix = R * g.wo;
  const X* tile = reinterpret_cast<const X*>(smem_raw);
  float* dtile = reinterpret_cast<float*>(smem_raw + tile_bytes);            // [8][npix][4]
  const uint32_t bar = tc::smem_u32(dtile + 8 * npix * 4);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cog = warp & 3, ci = warp >> 2;
  if (tid == 0) {
    tc::mbar_init(bar, 1);
    tc::fence_mbar_init();
    tma::prefetch_map(&tm_x);
  }
  float2 acc[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[t][i] = make_float2(0.f, 0.f);
  const int vec_per_row = kStemC / V;
  const int cv = tid % vec_per_row;
  const long long items = static_cast<long long>(g.nt) * bands;
  uint32_t phase = 0;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const long long nt = item / bands;
    const int ho0 = static_cast<int>(item - nt * bands) * R;
    const int rmax = min(R, g.ho - ho0);
    const int np = rmax * g.wo;
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tma::expect_tx(bar, static_cast<uint32_t>(3 * NR * bw * sizeof(X)));
      tma::load_4d(tc::smem_u32(smem_raw), &tm_x, bar, -kPad, 2 * ho0 - 1, 0, static_cast<int>(nt));
    }
    {
      RowLoader<T, V> ld;
      ld.init(dy, cv * V, kStemC);
      const long long q0 = (nt * g.ho + ho0) * g.wo;
      const int step = blockDim.x / vec_per_row;
#pragma unroll 1
      for (int p0 = tid / vec_per_row; p0 < np; p0 += 2 * step) {
        typename RowLoader<T, V>::Raw raw[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) if (p0 + j * step < np) raw[j] = ld.fetch(dy, q0 + p0 + j * step);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int p = p0 + j * step;
          if (p < np) {
            float v[V];
            ld.finish(dy, raw[j], v);
#pragma unroll
            for (int qd = 0; qd < QV; ++qd)
              *reinterpret_cast<float4*>(dtile + (static_cast<size_t>(cv * QV + qd) * npix + p) * 4) =
                  make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
          }
        }
      }
    }
    tc::mbar_wait(bar, phase);
    phase ^= 1;
    __syncthreads();
    for (int oh = 0; oh < rmax; ++oh) {
      for (int wo = lane; wo < g.wo; wo += 32) {
        const int p = oh * g.wo + wo;
        const float4 d0 = *reinterpret_cast<const float4*>(dtile + (static_cast<size_t>(cog * 2) * npix + p) * 4);
        const float4 d1 = *reinterpret_cast<const float4*>(dtile + (static_cast<size_t>(cog * 2 + 1) * npix + p) * 4);
        const float2 d[4] = {make_float2(d0.x, d0.y), make_float2(d0.z, d0.w), make_float2(d1.x, d1.y), make_float2(d1.z, d1.w)};
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const X* row = tile + (ci * NR + 2 * oh + kh) * bw + 2 * wo + (kPad - 1);
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float xv = lds_x<X>(row + kw);
            const float2 xx = make_float2(xv, xv);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[kh * 3 + kw][i] = __ffma2_rn(d[i], xx, acc[kh * 3 + kw][i]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = (i & 1) ? acc[t][i >> 1].y : acc[t][i >> 1].x;
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
      if (lane == 0) atomicAdd(&dwgt[(cog * 8 + i) * 27 + ci * 9 + t], v);
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

template <typename X, typename T>
static void stem_fwd32_go(const void* x, const float* w, void* out, double* stats, const StemGeom& g, cudaStream_t s) {
  const int R = 4, bands = static_cast<int>(cdiv(g.ho, R)), wp = g.wo + 1;
  const size_t smem = (static_cast<size_t>(29) * kStemC + 3 * (2 * R + 1) * 2 * wp) * sizeof(float);
  ensure_smem(stem_fwd32_kernel<X, T>, 200 * 1024);
  const long long items = static_cast<long long>(g.nt) * bands;
  const unsigned grid = static_cast<unsigned>(std::min<long long>(items, 4LL * kNumSMs));
  stem_fwd32_kernel<X, T><<<grid, 128, smem, s>>>(static_cast<const X*>(x), w, static_cast<T*>(out), stats, g, R, bands, wp);
}

template <typename X, typename T>
static void stem_wgrad32_go(const RowOp& dy, const void* x, float* dw, const StemGeom& g, cudaStream_t s) {
  const int R = 2, bands = static_cast<int>(cdiv(g.ho, R)), wp = g.wo + 1;
  const size_t smem = (static_cast<size_t>(3) * (2 * R + 1) * 2 * wp + static_cast<size_t>(8) * R * g.wo * 4) * sizeof(float);
  ensure_smem(stem_wgrad32_kernel<X, T>, 200 * 1024);
  const long long items = static_cast<long long>(g.nt) * bands;
  const unsigned grid = static_cast<unsigned>(std::min<long long>(items, 1LL * kNumSMs));
  stem_wgrad32_kernel<X, T><<<grid, 384, smem, s>>>(dy, static_cast<const X*>(x), dw, g, R, bands, wp);
}

// x: [NT, 3, H, W] -> 4-D map (W, H, 3, NT); box = (bw columns from -1, 2R+1 rows, 3 planes, 1 frame)
template <typename X>
static bool stem_tma_map(CUtensorMap* tm, const void* x, const StemGeom& g, int nrows, int* bw_out) {
  const int es = static_cast<int>(sizeof(X));
  const int bw = (g.w + 1 + 2 * (16 / es) - 1) / (16 / es) * (16 / es);   // columns -16/es .. w, rounded to 16 bytes
  if (bw > 256 || (static_cast<long long>(g.w) * es) % 16 != 0 || !aligned_to(x, 16)) return false;
  const unsigned long long dims[4] = {static_cast<unsigned long long>(g.w), static_cast<unsigned long long>(g.h), 3ULL,
                                      static_cast<unsigned long long>(g.nt)};
  const unsigned long long strides[3] = {static_cast<unsigned long long>(g.w) * es, static_cast<unsigned long long>(g.h) * g.w * es,
                                         3ULL * g.h * g.w * es};
  const unsigned box[4] = {static_cast<unsigned>(bw), static_cast<unsigned>(nrows), 3u, 1u};
  *bw_out = bw;
  return tma::make_map_4d(tm, es, x, dims, strides, box) == EHGR_OK;
}

template <typename X, typename T>
static bool stem_fwd32_tma_go(const void* x, const float* w, void* out, double* stats, const StemGeom& g, cudaStream_t s) {
  constexpr int R = 4;
  CUtensorMap tm;
  int bw = 0;
  if (!stem_tma_map<X>(&tm, x, g, 2 * R + 1, &bw)) return false;
  const int bands = static_cast<int>(cdiv(g.ho, R));
  const int tile_bytes = static_cast<int>((3 * (2 * R + 1) * bw * sizeof(X) + 127) / 128 * 128);
  const size_t smem = 2 * static_cast<size_t>(tile_bytes) + (29 * kStemC) * sizeof(float) + 16;
  ensure_smem(stem_fwd32_tma_kernel<X, T, R>, smem);
  const long long items = static_cast<long long>(g.nt) * bands;
  const unsigned grid = static_cast<unsigned>(std::min<long long>(items, 3LL * kNumSMs));
  stem_fwd32_tma_kernel<X, T, R><<<grid, 128, smem, s>>>(tm, w, static_cast<T*>(out), stats, g, bands, bw, tile_bytes, take_fin());
  return true;
}

template <typename X, typename T>
static bool stem_wgrad32_tma_go(const RowOp& dy, const void* x, float* dw, const StemGeom& g, cudaStream_t s) {
  constexpr int R = 2;
  CUtensorMap tm;
  int bw = 0;
  if (!stem_tma_map<X>(&tm, x, g, 2 * R + 1, &bw)) return false;
  const int bands = static_cast<int>(cdiv(g.ho, R));
  const int tile_bytes = static_cast<int>((3 * (2 * R + 1) * bw * sizeof(X) + 127) / 128 * 128);
  const size_t smem = static_cast<size_t>(tile_bytes) + static_cast<size_t>(8) * R * g.wo * 4 * sizeof(float) + 16;
  if (smem > 200 * 1024) return false;
  ensure_smem(stem_wgrad32_tma_kernel<X, T, R>, smem);
  const long long items = static_cast<long long>(g.nt) * bands;
  const unsigned grid = static_cast<unsigned>(std::min<long long>(items, 1LL * kNumSMs));
  stem_wgrad32_tma_kernel<X, T, R><<<grid, 384, smem, s>>>(dy, tm, dw, g, bands, bw, tile_bytes);
  return true;
}

// shared-memory budget of the banded kernels (image width bound)
static bool stem32_fits(const StemGeom& g) { return g.cout == kStemC && g.wo <= 1024; }

template <typename X>
static int stem_fwd_launch(const void* x, const float* w, void* out, double* stats, const StemGeom& g, int dtype,
                           cudaStream_t s) {
  if (stem32_fits(g)) {
    if (dtype == EHGR_F32) {
      if (!stem_fwd32_tma_go<X, float>(x, w, out, stats, g, s)) stem_fwd32_go<X, float>(x, w, out, stats, g, s);
    } else {
      if (!stem_fwd32_tma_go<X, __nv_bfloat16>(x, w, out, stats, g, s)) stem_fwd32_go<X, __nv_bfloat16>(x, w, out, stats, g, s);
    }
    return launch_status();
  }
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
  if (stem32_fits(g) && dy.mode != EHGR_ROW_GATE) {
    if (dtype == EHGR_F32) {
      if (!stem_wgrad32_tma_go<X, float>(dy, x, dw, g, s)) stem_wgrad32_go<X, float>(dy, x, dw, g, s);
    } else {
      if (!stem_wgrad32_tma_go<X, __nv_bfloat16>(dy, x, dw, g, s)) stem_wgrad32_go<X, __nv_bfloat16>(dy, x, dw, g, s);
    }
    return launch_status();
  }
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
  return ehgr_stem_fwd_bn(x, w, out, stats, nt, h, wd, cout, x_dtype, dtype, nullptr, stream);
}

extern "C" int ehgr_stem_fwd_bn(const void* x, const float* w, void* out, double* stats, int nt, int h, int wd,
                                int cout, int x_dtype, int dtype, const ehgr_bnfin* fin, ehgr_stream_t stream) {
  if (fin) {
    if (!fin->scale || !fin->shift || !fin->counter) return EHGR_E_NULL;
    if (fin->training && (!stats || fin->count <= 0)) return EHGR_E_NULL;
    if (!fin->training && (!fin->running_mean || !fin->running_var)) return EHGR_E_NULL;
  }
  StemGeom g;
  if (esize_of(x_dtype) == 0 || esize_of(dtype) == 0) return EHGR_E_DTYPE;
  if (!x || !w || !out) return EHGR_E_NULL;
  if (int st = stem_geom(g, nt, h, wd, cout)) return st;
  if (!aligned_to(out, 16) || !aligned_to(x, esize_of(x_dtype))) return EHGR_E_ALIGN;
  if (g.n_out == 0) return EHGR_OK;
  cudaStream_t s = as_stream(stream);
  fin_slot().fin = fin;            // the TMA band kernel finalises in its last CTA; the other variants leave it parked
  const int st = x_dtype == EHGR_F32 ? stem_fwd_launch<float>(x, w, out, stats, g, dtype, s)
                                     : stem_fwd_launch<__nv_bfloat16>(x, w, out, stats, g, dtype, s);
  return finish_fin(stats, cout, s, st);
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
