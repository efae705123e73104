// bn.cu — K14 BatchNorm2d bookkeeping around the lazily-applied normalisation, and the generic
// "materialise a row operand" elementwise kernel (BatchNorm apply + residual add, K11).
//
// nn.BatchNorm2d semantics (what the reference's modules do, archs/mobilenet_v2.py:10,17,41,...):
//   training: y = (x - mean_b) / sqrt(var_b + eps) * gamma + beta, var_b biased;
//             running = (1-momentum)*running + momentum*{mean_b, var_b * n/(n-1)}
//   eval    : y = (x - running_mean) / sqrt(running_var + eps) * gamma + beta
// Backward of conv -> BN(train) -> [ReLU6]:  with dz the (masked) gradient at the BN output,
//   d raw = scale * (dz - mean(dz) - xhat * mean(dz*xhat)),   dgamma = sum dz*xhat,  dbeta = sum dz
// which is written  ca*dz + cb*raw + cc  so that consumers can apply it while loading (rowop.cuh).
#include <type_traits>

#include "rowop.cuh"
#include "bnfin.cuh"

namespace ehgr {

__global__ void bn_finalize_fin_kernel(BnFin f, const double* __restrict__ stats, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) bn_finalize_channel(f, stats, c, C);
}

int bn_finalize_standalone(const BnFin& f, const double* stats, int c, cudaStream_t s) {
  bn_finalize_fin_kernel<<<static_cast<unsigned>(cdiv(c, 128)), 128, 0, s>>>(f, stats, c);
  return launch_status();
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rmean,
                                   float* __restrict__ rvar, float momentum, float eps, int training,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_o,
                                   float* __restrict__ invstd_o, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, invstd;
  if (training) {
    const double m = stats[c] / count;
    double var = stats[C + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean = static_cast<float>(m);
    invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    if (rmean) {
      const double unbiased = count > 1.0 ? var * (count / (count - 1.0)) : var;
      rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * static_cast<float>(unbiased);
    }
  } else {
    mean = rmean[c];
    invstd = 1.0f / sqrtf(rvar[c] + eps);
  }
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float sc = g * invstd;
  scale[c] = sc;
  shift[c] = b - mean * sc;
  if (mean_o) mean_o[c] = mean;
  if (invstd_o) invstd_o[c] = invstd;
}

// sums[c] += sum_m mask*g ; sums[C+c] += sum_m mask*g*raw.  block = (bx channel vectors, P rows),
// persistent grid-stride over rows; a thread keeps its channel vector(s) for its whole life.
template <typename T, bool relu6>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const T* __restrict__ g, const T* __restrict__ raw, const float* __restrict__ scale,
                     const float* __restrict__ shift, double* __restrict__ sums, long long M, int C, BnBwd fin, float hi) {
  constexpr int V = VecOf<T>::N;
  extern __shared__ float smem[];  // [2C]
  const int nthreads = blockDim.x * blockDim.y;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < 2 * C; i += nthreads) smem[i] = 0.f;
  __syncthreads();
  const long long row_stride = static_cast<long long>(gridDim.x) * blockDim.y;
  for (int cv = threadIdx.x; cv < C / V; cv += blockDim.x) {
    const int c0 = cv * V;
    float s[V], b[V], a1[V], a2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { a1[i] = a2[i] = 0.f; s[i] = 1.f; b[i] = 1.f; }
    if (relu6) { load_vec<float, V>(scale + c0, s); load_vec<float, V>(shift + c0, b); }
    constexpr int U = 4;                       // rows fetched per batch: 2U 16-byte loads in flight per thread
#pragma unroll 1
    for (long long m0 = static_cast<long long>(blockIdx.x) * blockDim.y + threadIdx.y; m0 < M; m0 += U * row_stride) {
      uint4 gq[U], rq[U];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const long long m = m0 + j * row_stride;
        gq[j] = rq[j] = make_uint4(0, 0, 0, 0);
        if (m < M) {
          gq[j] = *reinterpret_cast<const uint4*>(g + m * C + c0);
          rq[j] = *reinterpret_cast<const uint4*>(raw + m * C + c0);
        }
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        if constexpr (sizeof(T) == 2) {
          // bf16: register-pair arithmetic (FFMA2 / FADD2), accumulators indexed [2k], [2k+1]
          const uint32_t gw[4] = {gq[j].x, gq[j].y, gq[j].z, gq[j].w}, rw[4] = {rq[j].x, rq[j].y, rq[j].z, rq[j].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float2 gg = bf2_to_f2(gw[k]);
            const float2 rr = bf2_to_f2(rw[k]);
            if (relu6) {
              const float2 z = __ffma2_rn(rr, make_float2(s[2 * k], s[2 * k + 1]), make_float2(b[2 * k], b[2 * k + 1]));
              if (!(z.x > 0.f && z.x < hi)) gg.x = 0.f;
              if (!(z.y > 0.f && z.y < hi)) gg.y = 0.f;
            }
            const float2 n1 = __fadd2_rn(make_float2(a1[2 * k], a1[2 * k + 1]), gg);
            const float2 n2 = __ffma2_rn(gg, rr, make_float2(a2[2 * k], a2[2 * k + 1]));
            a1[2 * k] = n1.x; a1[2 * k + 1] = n1.y;
            a2[2 * k] = n2.x; a2[2 * k + 1] = n2.y;
          }
          continue;
        }
        float gv[V], rv[V];
        load_vec<T, V>(reinterpret_cast<const T*>(&gq[j]), gv);
        load_vec<T, V>(reinterpret_cast<const T*>(&rq[j]), rv);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float gg = gv[i];
          if (relu6) {
            const float z = fmaf(rv[i], s[i], b[i]);
            if (!(z > 0.f && z < hi)) gg = 0.f;
          }
          a1[i] += gg;
          a2[i] = fmaf(gg, rv[i], a2[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) { atomicAdd(&smem[c0 + i], a1[i]); atomicAdd(&smem[C + c0 + i], a2[i]); }
  }
  __syncthreads();
  for (int i = tid; i < 2 * C; i += nthreads) atomicAdd(&sums[i], static_cast<double>(smem[i]));
  if (fin.counter) {                                       // 2-D block: the helper's thread 0 / stride are 1-D
    __shared__ int s_last;
    if (tid < 2 * C) __threadfence();                      // the threads that issued this CTA's atomics
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(fin.counter, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (s_last) {
      __threadfence();
      for (int c = tid; c < C; c += nthreads) bn_bwd_finalize_channel(fin, sums, c, C);
      if (tid == 0) *fin.counter = 0u;
    }
  }
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                       const float* __restrict__ mean, const float* __restrict__ invstd, int training,
                                       float* __restrict__ ca, float* __restrict__ cb, float* __restrict__ cc,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double sdz = sums[c], sdzr = sums[C + c];
  const double mu = mean[c], is = invstd[c], g = gamma ? gamma[c] : 1.0;
  const double sdzx = (sdzr - mu * sdz) * is;  // sum dz * xhat
  if (dgamma) dgamma[c] = static_cast<float>(sdzx);
  if (dbeta) dbeta[c] = static_cast<float>(sdz);
  const double sc = g * is;
  if (training) {
    const double k1 = sdz / count, k2 = sdzx / count;
    ca[c] = static_cast<float>(sc);
    cb[c] = static_cast<float>(-sc * k2 * is);
    cc[c] = static_cast<float>(-sc * (k1 - mu * is * k2));
  } else {
    ca[c] = static_cast<float>(sc);
    cb[c] = 0.f;
    cc[c] = 0.f;
  }
}

// out = rowop(a) (+ addend).  block = (bx channel vectors, P rows): a thread keeps ONE channel vector
// (operand coefficients in registers) and strides over rows, four rows fetched per batch.
// kMode: the operand mode as a compile-time constant (the host dispatches on it): the row loader's other
// branches and their registers disappear, which is what bounds the occupancy of this streaming kernel.
template <typename T, int kMode>
__global__ void __launch_bounds__(256, 2)
row_apply_kernel(RowOp a_in, const T* __restrict__ addend, T* __restrict__ out, long long M, int C, int cv_total) {
  constexpr int V = VecOf<T>::N;
  constexpr int U = 4;
  using Ld = RowLoader<T, V, kMode == EHGR_ROW_BNBWD, kMode == EHGR_ROW_GATE>;
  RowOp a = a_in;
  a.mode = kMode;
  const long long row_stride = static_cast<long long>(gridDim.x) * blockDim.y;
  for (int cv = threadIdx.x; cv < cv_total; cv += blockDim.x) {
    const int c0 = cv * V;
    Ld ld;
    ld.init(a, c0, C);
    for (long long m0 = static_cast<long long>(blockIdx.x) * blockDim.y + threadIdx.y; m0 < M; m0 += U * row_stride) {
      typename Ld::Raw raw[U];
      uint4 ad[U];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const long long m = m0 + j * row_stride;
        if (m < M) {
          raw[j] = ld.fetch(a, m);
          if (addend) ad[j] = *reinterpret_cast<const uint4*>(addend + m * C + c0);
        }
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const long long m = m0 + j * row_stride;
        if (m < M) {
          if constexpr (sizeof(T) == 2) {
            if (!addend) {                       // packed path: FFMA2 arithmetic, result re-packed directly
              *reinterpret_cast<uint4*>(out + m * C + c0) = ld.finish_packed(a, raw[j]);
              continue;
            }
          }
          float v[V];
          ld.finish(a, raw[j], v);
          if (addend) {
            float av[V];
            load_vec<T, V>(reinterpret_cast<const T*>(&ad[j]), av);
#pragma unroll
            for (int k = 0; k < V; ++k) v[k] += av[k];
          }
          store_vec<T, V>(out + m * C + c0, v);
        }
      }
    }
  }
}

}  // namespace ehgr

using namespace ehgr;

extern "C" int ehgr_bn_finalize(const double* stats, long long count, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, float momentum, float eps, int training,
                                float* scale, float* shift, float* mean, float* invstd, int c,
                                ehgr_stream_t stream) {
  if (!scale || !shift) return EHGR_E_NULL;
  if (training && (!stats || count <= 0)) return EHGR_E_NULL;
  if (!training && (!running_mean || !running_var)) return EHGR_E_NULL;
  if (c <= 0) return EHGR_E_SHAPE;
  bn_finalize_kernel<<<static_cast<unsigned>(cdiv(c, 128)), 128, 0, as_stream(stream)>>>(
      stats, static_cast<double>(count), gamma, beta, running_mean, running_var, momentum, eps, training, scale,
      shift, mean, invstd, c);
  return launch_status();
}

extern "C" int ehgr_bn_bwd_reduce(const void* g, const void* raw, const float* scale, const float* shift,
                                  int relu6, double* sums, long long m, int c, int dtype, ehgr_stream_t stream) {
  return ehgr_bn_bwd_reduce_fin(g, raw, scale, shift, relu6, sums, m, c, dtype, nullptr, stream);
}

extern "C" int ehgr_bn_bwd_reduce_fin(const void* g, const void* raw, const float* scale, const float* shift,
                                      int relu6, double* sums, long long m, int c, int dtype, const ehgr_bnbwd* fin,
                                      ehgr_stream_t stream) {
  const int es = esize_of(dtype);
  BnBwd f{};
  if (fin) {
    f = *fin;
    if (!f.mean || !f.invstd || !f.ca || !f.cb || !f.cc || !f.counter) return EHGR_E_NULL;
    if (f.count <= 0) return EHGR_E_SHAPE;
  }
  if (es == 0) return EHGR_E_DTYPE;
  if (!g || !raw || !sums || (relu6 && (!scale || !shift))) return EHGR_E_NULL;
  const int V = 16 / es;
  if (m < 0 || c <= 0 || (c % V)) return EHGR_E_SHAPE;
  if (!aligned_to(g, 16) || !aligned_to(raw, 16)) return EHGR_E_ALIGN;
  if (m == 0) return EHGR_OK;
  const int cv = c / V;
  int bx = cv;
  while (bx > 256) bx = (bx + 1) / 2;
  const dim3 block(bx, std::max(1, 256 / bx));
  // every CTA ends with 2C double atomics: cap the CTA count for wide layers (they have few rows anyway)
  const long long cap = std::max<long long>(kNumSMs, std::min<long long>(4LL * kNumSMs, 300000LL / (2 * c)));   // one resident wave
  const long long blocks = std::max(1LL, std::min(cdiv(m, 4LL * block.y), cap));
  const size_t smem = static_cast<size_t>(2) * c * sizeof(float);
  cudaStream_t s = as_stream(stream);
  auto go = [&](auto tag, auto rtag) {
    using T = decltype(tag);
    bn_bwd_reduce_kernel<T, decltype(rtag)::value><<<static_cast<unsigned>(blocks), block, smem, s>>>(
        static_cast<const T*>(g), static_cast<const T*>(raw), scale, shift, sums, m, c, f, relu6 == 2 ? INFINITY : 6.f);
  };
  if (dtype == EHGR_F32) {
    if (relu6) go(float{}, std::true_type{}); else go(float{}, std::false_type{});
  } else {
    if (relu6) go(__nv_bfloat16{}, std::true_type{}); else go(__nv_bfloat16{}, std::false_type{});
  }
  return launch_status();
}

extern "C" int ehgr_bn_bwd_finalize(const double* sums, long long count, const float* gamma, const float* mean,
                                    const float* invstd, int training, float* ca, float* cb, float* cc,
                                    float* dgamma, float* dbeta, int c, ehgr_stream_t stream) {
  if (!sums || !mean || !invstd || !ca || !cb || !cc) return EHGR_E_NULL;
  if (c <= 0 || count <= 0) return EHGR_E_SHAPE;
  bn_bwd_finalize_kernel<<<static_cast<unsigned>(cdiv(c, 128)), 128, 0, as_stream(stream)>>>(
      sums, static_cast<double>(count), gamma, mean, invstd, training, ca, cb, cc, dgamma, dbeta, c);
  return launch_status();
}

extern "C" int ehgr_row_apply(const ehgr_rowop* a, const void* addend, void* out, long long m, int c, int dtype,
                              ehgr_stream_t stream) {
  const int es = esize_of(dtype);
  if (es == 0) return EHGR_E_DTYPE;
  if (!out) return EHGR_E_NULL;
  if (int st = validate_rowop(a, es)) return st;
  const int V = 16 / es;
  if (m < 0 || c <= 0 || (c % V)) return EHGR_E_SHAPE;
  if (!aligned_to(out, 16) || (addend && !aligned_to(addend, 16))) return EHGR_E_ALIGN;
  if (m == 0) return EHGR_OK;
  const int cv = c / V;
  int bx = cv;
  while (bx > 256) bx = (bx + 1) / 2;
  const dim3 block(bx, std::max(1, 256 / bx));
  const long long blocks = std::max(1LL, std::min(cdiv(m, 4LL * block.y), 4LL * kNumSMs));
  cudaStream_t s = as_stream(stream);
  auto launch = [&](auto tag, auto mode_tag) {
    using T = decltype(tag);
    row_apply_kernel<T, decltype(mode_tag)::value><<<static_cast<unsigned>(blocks), block, 0, s>>>(
        *a, static_cast<const T*>(addend), static_cast<T*>(out), m, c, cv);
  };
  auto by_mode = [&](auto tag) {
    switch (a->mode) {
      case EHGR_ROW_PLAIN: launch(tag, std::integral_constant<int, EHGR_ROW_PLAIN>{}); break;
      case EHGR_ROW_AFFINE: launch(tag, std::integral_constant<int, EHGR_ROW_AFFINE>{}); break;
      case EHGR_ROW_SHIFT: launch(tag, std::integral_constant<int, EHGR_ROW_SHIFT>{}); break;
      case EHGR_ROW_BNBWD: launch(tag, std::integral_constant<int, EHGR_ROW_BNBWD>{}); break;
      default: launch(tag, std::integral_constant<int, EHGR_ROW_GATE>{}); break;
    }
  };
  if (dtype == EHGR_F32) by_mode(float{}); else by_mode(__nv_bfloat16{});
  return launch_status();
}
