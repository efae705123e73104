// optim.cu — N1: the optimiser step of the training loops (torch.optim.SGD with momentum and weight decay over
// the policy groups of get_optim_policies, train_mtmm.py:576-585 / utils.py:39-46) as ONE kernel over flat
// buffers:   d = g + wd*decay_mult[k]*p;   buf = momentum*buf + d;   p -= lr*lr_mult[k]*buf
// (dampening 0, no Nesterov — the reference's configuration; a zero-initialised momentum buffer reproduces
// torch's first-step rule buf = d).  k = group code of the element (one byte per element); the base learning
// rate is read from DEVICE memory so that a schedule (adjust_learning_rate) never invalidates a captured graph.
#include "common.cuh"

namespace ehgr {

__global__ void __launch_bounds__(256)
sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, const uint8_t* __restrict__ code,
                const float* __restrict__ lr_mult, const float* __restrict__ decay_mult, int n_groups,
                const float* __restrict__ lr_dev, float momentum, float weight_decay, long long n) {
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
  }
}

}  // namespace ehgr

using namespace ehgr;

extern "C" int ehgr_sgd_step(float* p, const float* g, float* buf, const void* code, const float* lr_mult,
                             const float* decay_mult, int n_groups, const float* lr_dev, float momentum, float weight_decay,
                             long long n, ehgr_stream_t stream) {
  if (!p || !g || !buf || !code || !lr_mult || !decay_mult || !lr_dev) return EHGR_E_NULL;
  if (n < 0 || (n % 4) || n_groups <= 0 || n_groups > 64) return EHGR_E_SHAPE;
  if (!aligned_to(p, 16) || !aligned_to(g, 16) || !aligned_to(buf, 16) || !aligned_to(code, 4)) return EHGR_E_ALIGN;
  if (n == 0) return EHGR_OK;
  const unsigned blocks = static_cast<unsigned>(std::max(1LL, std::min(cdiv(n / 4, 256), 8LL * kNumSMs)));
  sgd_step_kernel<<<blocks, 256, 0, as_stream(stream)>>>(p, g, buf, static_cast<const uint8_t*>(code), lr_mult, decay_mult,
                                                          n_groups, lr_dev, momentum, weight_decay, n);
  return launch_status();
}
