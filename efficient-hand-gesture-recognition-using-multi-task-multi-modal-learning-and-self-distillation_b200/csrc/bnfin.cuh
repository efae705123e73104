// bnfin.cuh — BatchNorm bookkeeping folded into the PRODUCING kernel ("last CTA done").
//
// Every kernel that produces a raw activation accumulates its per-channel batch statistics; instead of a separate
// one-block launch per layer (52 + 52 launches per step, ~0.8 ms), the CTA that arrives LAST at a per-layer ticket
// counter reads the finished statistics and runs the finalisation itself:
//   forward  (ehgr_bnfin) : statistics -> scale / shift / mean / invstd, running-statistics update (nn.BatchNorm2d)
//   backward (ehgr_bnbwd) : sums -> ca / cb / cc (the BNBWD row operand's coefficients), d(gamma), d(beta)
// Protocol: each thread __threadfence()s its atomics, the CTA syncs, thread 0 takes a ticket; the CTA that draws
// gridDim.x - 1 sees every other CTA's atomics (fence + atomic ticket), finalises and resets the counter to 0
// (the counter is zero before every launch; CUDA-graph replays rely on the reset).
#pragma once
#include "common.cuh"

namespace ehgr {

using BnFin = ehgr_bnfin;
using BnBwd = ehgr_bnbwd;

__device__ __forceinline__ double ldcg_f64(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ void bn_finalize_channel(const BnFin& f, const double* stats, int c, int C) {
  float mean, invstd;
  if (f.training) {
    const double count = static_cast<double>(f.count);
    const double m = ldcg_f64(stats + c) / count;
    double var = ldcg_f64(stats + C + c) / count - m * m;
    if (var < 0.0) var = 0.0;
    mean = static_cast<float>(m);
    invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(f.eps)));
    if (f.running_mean) {
      const double unbiased = count > 1.0 ? var * (count / (count - 1.0)) : var;
      f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * mean;
      f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * static_cast<float>(unbiased);
    }
  } else {
    mean = f.running_mean[c];
    invstd = 1.0f / sqrtf(f.running_var[c] + f.eps);
  }
  const float g = f.gamma ? f.gamma[c] : 1.f, b = f.beta ? f.beta[c] : 0.f;
  const float sc = g * invstd;
  f.scale[c] = sc;
  f.shift[c] = b - mean * sc;
  if (f.mean) f.mean[c] = mean;
  if (f.invstd) f.invstd[c] = invstd;
}

__device__ __forceinline__ void bn_bwd_finalize_channel(const BnBwd& f, const double* sums, int c, int C) {
  const double count = static_cast<double>(f.count);
  const double sdz = ldcg_f64(sums + c), sdzr = ldcg_f64(sums + C + c);
  const double mu = f.mean[c], is = f.invstd[c], g = f.gamma ? f.gamma[c] : 1.0;
  const double sdzx = (sdzr - mu * sdz) * is;  // sum dz * xhat
  if (f.dgamma) f.dgamma[c] = static_cast<float>(sdzx);
  if (f.dbeta) f.dbeta[c] = static_cast<float>(sdz);
  const double sc = g * is;
  if (f.training) {
    const double k1 = sdz / count, k2 = sdzx / count;
    f.ca[c] = static_cast<float>(sc);
    f.cb[c] = static_cast<float>(-sc * k2 * is);
    f.cc[c] = static_cast<float>(-sc * (k1 - mu * is * k2));
  } else {
    f.ca[c] = static_cast<float>(sc);
    f.cb[c] = 0.f;
    f.cc[c] = 0.f;
  }
}

// Call from ALL threads of the CTA after the CTA's last statistics atomic.  True (in every thread) in the CTA that
// arrives last; that CTA may read everything the other CTAs accumulated.  `did_atomics`: this thread issued some of
// the CTA's statistics atomics — only those threads need the (expensive: it drains the thread's outstanding stores)
// device-scope fence before the ticket.
__device__ __forceinline__ bool cta_arrives_last(unsigned int* counter, bool did_atomics) {
  __shared__ int s_last;
  if (did_atomics) __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  const bool last = s_last != 0;
  if (last) __threadfence();
  return last;
}

// the whole epilogue: no-op when fin.counter is NULL
__device__ __forceinline__ void bn_finalize_if_last(const BnFin& f, const double* stats, int C, bool did_atomics = true) {
  if (!f.counter) return;
  if (!cta_arrives_last(f.counter, did_atomics)) return;
  for (int c = threadIdx.x; c < C; c += blockDim.x) bn_finalize_channel(f, stats, c, C);
  if (threadIdx.x == 0) *f.counter = 0u;
}
__device__ __forceinline__ void bn_bwd_finalize_if_last(const BnBwd& f, const double* sums, int C, bool did_atomics = true) {
  if (!f.counter) return;
  if (!cta_arrives_last(f.counter, did_atomics)) return;
  for (int c = threadIdx.x; c < C; c += blockDim.x) bn_bwd_finalize_channel(f, sums, c, C);
  if (threadIdx.x == 0) *f.counter = 0u;
}

// Host side: an entry point that accepts `fin` parks it here; a kernel launcher that implements the in-kernel
// finalisation takes it (take_fin) and passes it to its kernel; whatever is still parked when the entry point's
// producer has been launched is finalised by the stand-alone kernel (finish_fin).  thread_local: re-entrant.
struct FinSlot {
  const BnFin* fin = nullptr;
};
FinSlot& fin_slot();
inline BnFin take_fin() {
  FinSlot& s = fin_slot();
  BnFin f{};
  if (s.fin) { f = *s.fin; s.fin = nullptr; }
  return f;
}
int bn_finalize_standalone(const BnFin& f, const double* stats, int c, cudaStream_t s);
inline int finish_fin(const double* stats, int c, cudaStream_t s, int status) {
  FinSlot& slot = fin_slot();
  const BnFin* f = slot.fin;
  slot.fin = nullptr;
  if (status != EHGR_OK || !f) return status;
  return bn_finalize_standalone(*f, stats, c, s);
}

}  // namespace ehgr
