// input.cu — N4: the input pipeline's last two steps on the device.  The reference converts decoded
// uint8 frames to float and normalises them on the CPU (ToTorchFormatTensor(div=True) then GroupNormalize,
// models/spatial_transforms.py:66-80,489-503) and ships fp32 to the GPU; here the uint8 frames travel
// (4x fewer H2D bytes) and one streaming kernel evaluates the same two fp32 operations in the same
// order,  y = (float(x) / 255 - mean[c]) / std[c],  so the result is bit-identical to the CPU tensors.
#include "rowop.cuh"

namespace ehgr {

template <typename T>
__global__ void __launch_bounds__(256)
normalize_u8_kernel(const uint8_t* __restrict__ src, T* __restrict__ dst, long long n_planes, int channels, long long plane,
                    const float* __restrict__ mean, const float* __restrict__ stdv, float div) {
  const long long vec_per_plane = (plane % 16 == 0) ? plane / 16 : 0;   // 16-byte loads need 16-pixel-aligned planes
  const long long total = n_planes * vec_per_plane;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pl = i / vec_per_plane;
    const int c = static_cast<int>(pl % channels);
    const float m = mean ? mean[c] : 0.f, s = stdv ? stdv[c] : 1.f;
    const long long off = pl * plane + (i - pl * vec_per_plane) * 16;
    const uint4 raw = *reinterpret_cast<const uint4*>(src + off);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float x = static_cast<float>((w[q] >> (8 * b)) & 0xffu);
        v[b] = __fdiv_rn(__fsub_rn(__fdiv_rn(x, div), m), s);
      }
      store_vec<T, 4>(dst + off + 4 * q, v);
    }
  }
  // ragged tail of each plane (plane % 16 pixels), one thread per pixel
  const long long tail = plane - vec_per_plane * 16;
  if (tail > 0) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_planes * tail;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const long long pl = i / tail;
      const int c = static_cast<int>(pl % channels);
      const long long off = pl * plane + vec_per_plane * 16 + (i - pl * tail);
      const float m = mean ? mean[c] : 0.f, s = stdv ? stdv[c] : 1.f;
      const float y = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(src[off]), div), m), s);
      dst[off] = static_cast<T>(y);
    }
  }
}

}  // namespace ehgr

using namespace ehgr;

extern "C" int ehgr_normalize_u8(const void* src, void* dst, long long n_planes, int channels, long long plane,
                                 const float* mean, const float* stdv, float div, int dst_dtype, ehgr_stream_t stream) {
  if (esize_of(dst_dtype) == 0) return EHGR_E_DTYPE;
  if (!src || !dst) return EHGR_E_NULL;
  if (n_planes < 0 || channels <= 0 || plane <= 0 || (n_planes % channels) || div == 0.f) return EHGR_E_SHAPE;
  if (!aligned_to(src, 16) || !aligned_to(dst, 16)) return EHGR_E_ALIGN;
  if (n_planes == 0) return EHGR_OK;
  const long long work = plane % 16 == 0 ? n_planes * (plane / 16) : n_planes * plane;
  const unsigned blocks = static_cast<unsigned>(std::max(1LL, std::min(cdiv(work, 256), 16LL * kNumSMs)));
  cudaStream_t s = as_stream(stream);
  if (dst_dtype == EHGR_F32)
    normalize_u8_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const uint8_t*>(src), static_cast<float*>(dst), n_planes,
                                                      channels, plane, mean, stdv, div);
  else
    normalize_u8_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(static_cast<const uint8_t*>(src),
                                                              static_cast<__nv_bfloat16*>(dst), n_planes, channels, plane,
                                                              mean, stdv, div);
  return launch_status();
}
