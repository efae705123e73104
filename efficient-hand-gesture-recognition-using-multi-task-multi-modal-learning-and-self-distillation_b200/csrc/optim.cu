// optim.cu — N1: the optimiser step of the training loops (torch.optim.SGD with momentum and weight decay over
// the policy groups of get_optim_policies, train_mtmm.py:576-585 / utils.py:39-46) as ONE kernel over flat
// buffers:   d = g + wd*decay_mult[k]*p;   buf = momentum*buf + d;   p -= lr*lr_mult[k]*buf
// (dampening 0, no Nesterov — the reference's configuration; a zero-initialised momentum buffer reproduces
// torch's first-step rule buf = d).  k = group code of the element (one byte per element); the base learning
// rate is read from DEVICE memory so that a schedule (adjust_learning_rate) never invalidates a captured graph.
//
// EMA (the reference's EMAWrapper, train_mtmm.py:110-128, updated after every optimiser step, :245):
//   ema = decay * ema + (1 - decay) * value   over EVERY state_dict entry.  For the parameters it is folded into the
// SGD kernel (the freshly updated parameter is still in registers); floating-point buffers (BatchNorm running
// statistics) and the int64 num_batches_tracked counters have one small kernel each.  The three roundings of the
// reference expression (two products, one sum, each a separate fp32 op) are reproduced: bit-identical results.
#include "common.cuh"

namespace ehgr {

__global__ void __launch_bounds__(256)
sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, const uint8_t* __restrict__ code,
                const float* __restrict__ lr_mult, const float* __restrict__ decay_mult, int n_groups,
                const float* __restrict__ lr_dev, float momentum, float weight_decay, long long n,
                float* __restrict__ ema, float ema_decay, float ema_rest, __nv_bfloat16* __restrict__ p16) {
  __shared__ float s_lr[64], s_wd[64];
  const float lr = *lr_dev;
  for (int i = threadIdx.x; i < n_groups; i += blockDim.x) {
    s_lr[i] = lr * lr_mult[i];
    s_wd[i] = weight_decay * decay_mult[i];
  }
  __syncthreads();
  const long long n4 = n / 4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 bv = reinterpret_cast<float4*>(buf)[i];
    const uchar4 cv = reinterpret_cast<const uchar4*>(code)[i];
    float* pp = reinterpret_cast<float*>(&pv);
    const float* gp = reinterpret_cast<const float*>(&gv);
    float* bp = reinterpret_cast<float*>(&bv);
    const uint8_t cs[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (cs[k] == 255) continue;                            // padding / frozen element
      const float d = __fadd_rn(gp[k], __fmul_rn(s_wd[cs[k]], pp[k]));
      bp[k] = __fadd_rn(__fmul_rn(momentum, bp[k]), d);
      pp[k] = __fsub_rn(pp[k], __fmul_rn(s_lr[cs[k]], bp[k]));
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(buf)[i] = bv;
    if (p16) {                                               // bf16 mirror of the parameters for the tensor-core kernels
      const __nv_bfloat162 lo = __floats2bfloat162_rn(pp[0], pp[1]), hi = __floats2bfloat162_rn(pp[2], pp[3]);
      reinterpret_cast<uint2*>(p16)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
    if (ema) {                                               // padding elements: 0 stays 0
      float4 ev = reinterpret_cast<float4*>(ema)[i];
      float* ep = reinterpret_cast<float*>(&ev);
#pragma unroll
      for (int k = 0; k < 4; ++k) ep[k] = __fadd_rn(__fmul_rn(ema_decay, ep[k]), __fmul_rn(ema_rest, pp[k]));
      reinterpret_cast<float4*>(ema)[i] = ev;
    }
  }
}

__global__ void __launch_bounds__(256)
ema_f32_kernel(float* __restrict__ ema, const float* __restrict__ x, float decay, float rest, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    ema[i] = __fadd_rn(__fmul_rn(decay, ema[i]), __fmul_rn(rest, x[i]));
}

// int64 entries: the reference evaluates python_float * int64_tensor in the default float32 and copy_()s the sum back
// into the int64 tensor (truncation toward zero)
__global__ void __launch_bounds__(256)
ema_i64_kernel(long long* __restrict__ ema, const long long* __restrict__ x, float decay, float rest, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    ema[i] = static_cast<long long>(__fadd_rn(__fmul_rn(decay, static_cast<float>(ema[i])), __fmul_rn(rest, static_cast<float>(x[i]))));
}

}  // namespace ehgr

using namespace ehgr;

static void ema_coefs(double decay, float* d, float* rest) {
  *d = static_cast<float>(decay);              // python scalars reach the fp32 kernels as (float)decay and (float)(1. - decay)
  *rest = static_cast<float>(1.0 - decay);
}

extern "C" int ehgr_sgd_step(float* p, const float* g, float* buf, const void* code, const float* lr_mult,
                             const float* decay_mult, int n_groups, const float* lr_dev, float momentum, float weight_decay,
                             long long n, float* ema, double ema_decay, void* p16, ehgr_stream_t stream) {
  if (!p || !g || !buf || !code || !lr_mult || !decay_mult || !lr_dev) return EHGR_E_NULL;
  if (n < 0 || (n % 4) || n_groups <= 0 || n_groups > 64) return EHGR_E_SHAPE;
  if (!aligned_to(p, 16) || !aligned_to(g, 16) || !aligned_to(buf, 16) || !aligned_to(code, 4) || (ema && !aligned_to(ema, 16)) ||
      (p16 && !aligned_to(p16, 8)))
    return EHGR_E_ALIGN;
  float ed = 0.f, er = 0.f;
  ema_coefs(ema_decay, &ed, &er);
  if (n == 0) return EHGR_OK;
  const unsigned blocks = static_cast<unsigned>(std::max(1LL, std::min(cdiv(n / 4, 256), 8LL * kNumSMs)));
  sgd_step_kernel<<<blocks, 256, 0, as_stream(stream)>>>(p, g, buf, static_cast<const uint8_t*>(code), lr_mult, decay_mult,
                                                          n_groups, lr_dev, momentum, weight_decay, n, ema, ed, er,
                                                          static_cast<__nv_bfloat16*>(p16));
  return launch_status();
}

extern "C" int ehgr_ema_update(void* ema, const void* x, long long n, double decay, int is_int64, ehgr_stream_t stream) {
  if (!ema || !x) return EHGR_E_NULL;
  if (n < 0) return EHGR_E_SHAPE;
  if (!aligned_to(ema, is_int64 ? 8 : 4) || !aligned_to(x, is_int64 ? 8 : 4)) return EHGR_E_ALIGN;
  if (n == 0) return EHGR_OK;
  float ed, er;
  ema_coefs(decay, &ed, &er);
  const unsigned blocks = static_cast<unsigned>(std::max(1LL, std::min(cdiv(n, 256), 4LL * kNumSMs)));
  if (is_int64)
    ema_i64_kernel<<<blocks, 256, 0, as_stream(stream)>>>(static_cast<long long*>(ema), static_cast<const long long*>(x), ed, er, n);
  else
    ema_f32_kernel<<<blocks, 256, 0, as_stream(stream)>>>(static_cast<float*>(ema), static_cast<const float*>(x), ed, er, n);
  return launch_status();
}
