// loss.cu — K12 / K13: the two loss heads, forward AND backward in a single launch each.
//
// K12 MTMM (train_mtmm.py:223-231):  loss = CE(logits, y) + w_d * MSE(pred, bilinear56(depth))
//   The 224->56 bilinear resize with align_corners=False samples at 4i+1.5, i.e. it is the mean of the
//   2x2 centre pixels (rows/cols 4i+1, 4i+2) of every 4x4 cell — read directly, no resized tensor.
// K13 SD (train_sd.py:178-193,227-265):
//   L = (1-a) * sum_{i=0..3} CE(z_i, y) + a * T^2 * sum_{i=1..3} KD(z_i, softmax(z_0/T).detach())
//       + b * sum_{i=1..3} sum (f_i - f_0.detach())^2 * [(f_i > 0) | (f_0 > 0)]
// The reference computes these with ~5 / ~40 small kernels and 4 / 17 .item() synchronisations.
#include "rowop.cuh"

namespace ehgr {

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.f;
  if (warp == 0) r = warp_sum(r);
  __syncthreads();
  if (threadIdx.x == 0) scratch[0] = r;
  __syncthreads();
  r = scratch[0];
  return r;
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  v = warp_max(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? scratch[threadIdx.x] : -INFINITY;
  if (warp == 0) r = warp_max(r);
  __syncthreads();
  if (threadIdx.x == 0) scratch[0] = r;
  __syncthreads();
  r = scratch[0];
  return r;
}

// log-sum-exp of z[0..K) * inv_t over the block (every thread gets the result)
__device__ __forceinline__ float block_lse(const float* __restrict__ z, int K, float inv_t, float* scratch) {
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < K; j += blockDim.x) mx = fmaxf(mx, z[j] * inv_t);
  mx = block_max(mx, scratch);
  float s = 0.f;
  for (int j = threadIdx.x; j < K; j += blockDim.x) s += expf(z[j] * inv_t - mx);
  s = block_sum(s, scratch);
  return mx + logf(s);
}

template <typename T>
__global__ void __launch_bounds__(128)
mtmm_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, const T* __restrict__ pred,
                 const float* __restrict__ gt, float wd, float* __restrict__ loss_out, float* __restrict__ dlogits,
                 float* __restrict__ dpred, int N, int K, long long cells, int ph, int pw, int ce_blocks) {
  __shared__ float scratch[32];
  if (static_cast<int>(blockIdx.x) < ce_blocks) {
    const int n = blockIdx.x;
    const float* z = logits + static_cast<size_t>(n) * K;
    const float lse = block_lse(z, K, 1.f, scratch);
    const long long y64 = labels[n];
    // A label outside [0, K) never indexes memory: the row's loss and gradient become NaN (the reference's
    // nn.CrossEntropyLoss raises a device assert there).  ignore_index is not supported.
    const bool bad = y64 < 0 || y64 >= K;
    const int y = bad ? 0 : static_cast<int>(y64);
    const float inv_n = 1.f / static_cast<float>(N);
    for (int j = threadIdx.x; j < K; j += blockDim.x)
      dlogits[static_cast<size_t>(n) * K + j] = bad ? NAN : (expf(z[j] - lse) - (j == y ? 1.f : 0.f)) * inv_n;
    if (threadIdx.x == 0) {
      const float ce = bad ? NAN : (lse - z[y]) * inv_n;
      atomicAdd(&loss_out[1], ce);
      atomicAdd(&loss_out[0], ce);
    }
    return;
  }
  // depth part: grid-stride over the ph*pw cells of every frame
  const int gw = 4 * pw;
  const float inv = 1.f / static_cast<float>(cells);
  float part = 0.f;
  const long long stride = static_cast<long long>(gridDim.x - ce_blocks) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x - ce_blocks) * blockDim.x + threadIdx.x; i < cells; i += stride) {
    const int j = static_cast<int>(i % pw);
    const long long r = i / pw;
    const int ii = static_cast<int>(r % ph);
    const long long fr = r / ph;
    const float* g = gt + (fr * (4LL * ph) + (4 * ii + 1)) * gw + 4 * j + 1;
    const float2 r0 = make_float2(__ldg(g), __ldg(g + 1));
    const float2 r1 = make_float2(__ldg(g + gw), __ldg(g + gw + 1));
    // same association as bilinear interpolation: blend columns, then rows, weights 0.5
    const float target = 0.5f * (0.5f * r0.x + 0.5f * r0.y) + 0.5f * (0.5f * r1.x + 0.5f * r1.y);
    float p;
    if constexpr (sizeof(T) == 4) p = reinterpret_cast<const float*>(pred)[i];
    else p = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(pred)[i]);
    const float d = p - target;
    part = fmaf(d, d, part);
    dpred[i] = 2.f * d * inv * wd;
  }
  part = block_sum(part, scratch);
  if (threadIdx.x == 0) {
    const float mse = part * inv;
    atomicAdd(&loss_out[2], mse);
    atomicAdd(&loss_out[0], wd * mse);
  }
}

struct SdPtrs {
  const float* logits[4];
  const float* feats[4];
  float* dlogits[4];
  float* dfeats[3];
};

__global__ void __launch_bounds__(128)
sd_loss_kernel(SdPtrs p, const long long* __restrict__ labels, float alpha, float beta, float temp,
               float* __restrict__ terms, int N, int K, long long feat_elems, int row_blocks) {
  __shared__ float scratch[32];
  if (static_cast<int>(blockIdx.x) < row_blocks) {
    const int n = blockIdx.x;
    const long long y64 = labels[n];
    const bool bad = y64 < 0 || y64 >= K;           // out-of-range label: NaN loss / gradient, no out-of-bounds read
    const int y = bad ? 0 : static_cast<int>(y64);
    const float inv_n = bad ? NAN : 1.f / static_cast<float>(N), inv_t = 1.f / temp;
    const float* z0 = p.logits[0] + static_cast<size_t>(n) * K;
    const float lse0 = block_lse(z0, K, 1.f, scratch);
    const float lse0t = block_lse(z0, K, inv_t, scratch);
    for (int j = threadIdx.x; j < K; j += blockDim.x)
      p.dlogits[0][static_cast<size_t>(n) * K + j] = (1.f - alpha) * inv_n * (expf(z0[j] - lse0) - (j == y ? 1.f : 0.f));
    if (threadIdx.x == 0) {
      const float ce = (lse0 - z0[y]) * inv_n;
      atomicAdd(&terms[1], ce);
      atomicAdd(&terms[0], (1.f - alpha) * ce);
    }
    for (int i = 1; i < 4; ++i) {
      const float* z = p.logits[i] + static_cast<size_t>(n) * K;
      const float lse = block_lse(z, K, 1.f, scratch);
      const float lset = block_lse(z, K, inv_t, scratch);
      float kd = 0.f;
      for (int j = threadIdx.x; j < K; j += blockDim.x) {
        const float soft = expf(z0[j] * inv_t - lse0t);        // teacher softmax(z0/T), detached
        const float ls = z[j] * inv_t - lset;                    // student log_softmax(z/T)
        kd -= ls * soft;
        const float g_ce = expf(z[j] - lse) - (j == y ? 1.f : 0.f);
        const float g_kd = (expf(ls) - soft) * inv_t;
        p.dlogits[i][static_cast<size_t>(n) * K + j] = inv_n * ((1.f - alpha) * g_ce + alpha * temp * temp * g_kd);
      }
      kd = block_sum(kd, scratch);
      if (threadIdx.x == 0) {
        const float ce = (lse - z[y]) * inv_n;
        const float kdm = kd * inv_n * temp * temp;
        atomicAdd(&terms[1 + i], ce);
        atomicAdd(&terms[4 + i], kdm);
        atomicAdd(&terms[0], (1.f - alpha) * ce + alpha * kdm);
      }
    }
    return;
  }
  float part[3] = {0.f, 0.f, 0.f};
  const long long stride = static_cast<long long>(gridDim.x - row_blocks) * blockDim.x;
  for (long long e = static_cast<long long>(blockIdx.x - row_blocks) * blockDim.x + threadIdx.x; e < feat_elems; e += stride) {
    const float f0 = p.feats[0][e];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float f = p.feats[i + 1][e];
      const float d = f - f0;
      const float m = (f > 0.f || f0 > 0.f) ? 1.f : 0.f;
      part[i] = fmaf(d * d, m, part[i]);
      p.dfeats[i][e] = 2.f * beta * d * m;
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float s = block_sum(part[i], scratch);
    if (threadIdx.x == 0 && s != 0.f) {
      atomicAdd(&terms[8 + i], s);
      atomicAdd(&terms[0], beta * s);
    }
  }
}

}  // namespace ehgr

using namespace ehgr;

extern "C" int ehgr_mtmm_loss(const float* logits, const long long* labels, const void* pred, const float* depth_gt,
                              float depth_weight, float* loss_out, float* dlogits, float* dpred, int n, int k,
                              int frames, int ph, int pw, int dtype, ehgr_stream_t stream) {
  if (esize_of(dtype) == 0) return EHGR_E_DTYPE;
  // n == 0: the depth term alone (the MTMM+SD step adds it to the SD loss kernel's terms, train_mtmm_sd.py:240-293)
  if (!pred || !depth_gt || !loss_out || !dpred) return EHGR_E_NULL;
  if (n > 0 && (!logits || !labels || !dlogits)) return EHGR_E_NULL;
  if (n < 0 || (n > 0 && k <= 0) || frames <= 0 || ph <= 0 || pw <= 0) return EHGR_E_SHAPE;
  const long long cells = static_cast<long long>(frames) * ph * pw;
  const long long depth_blocks = std::min(cdiv(cells, 128 * 4), 8LL * kNumSMs);
  const unsigned grid = static_cast<unsigned>(n + depth_blocks);
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32)
    mtmm_loss_kernel<float><<<grid, 128, 0, s>>>(logits, labels, static_cast<const float*>(pred), depth_gt, depth_weight,
                                                 loss_out, dlogits, dpred, n, k, cells, ph, pw, n);
  else
    mtmm_loss_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(logits, labels, static_cast<const __nv_bfloat16*>(pred),
                                                         depth_gt, depth_weight, loss_out, dlogits, dpred, n, k, cells,
                                                         ph, pw, n);
  return launch_status();
}

extern "C" int ehgr_sd_loss(const float* const* logits, const float* const* feats, const long long* labels,
                            float alpha, float beta, float temperature, float* terms_out, float* const* dlogits,
                            float* const* dfeats, int n, int k, long long rows, int f, ehgr_stream_t stream) {
  if (!logits || !feats || !labels || !terms_out || !dlogits || !dfeats) return EHGR_E_NULL;
  if (n <= 0 || k <= 0 || rows <= 0 || f <= 0 || temperature <= 0.f) return EHGR_E_SHAPE;
  SdPtrs p;
  for (int i = 0; i < 4; ++i) {
    if (!logits[i] || !feats[i] || !dlogits[i]) return EHGR_E_NULL;
    p.logits[i] = logits[i];
    p.feats[i] = feats[i];
    p.dlogits[i] = dlogits[i];
  }
  for (int i = 0; i < 3; ++i) {
    if (!dfeats[i]) return EHGR_E_NULL;
    p.dfeats[i] = dfeats[i];
  }
  const long long elems = rows * f;
  const long long fb = std::min(cdiv(elems, 128 * 4), 8LL * kNumSMs);
  sd_loss_kernel<<<static_cast<unsigned>(n + fb), 128, 0, as_stream(stream)>>>(p, labels, alpha, beta, temperature,
                                                                              terms_out, n, k, elems, n);
  return launch_status();
}
