// common.cuh — shared device/host helpers for libehgr_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/ehgr_b200.h"

namespace ehgr {

extern std::atomic<long long> g_launches;

// Every launch goes through this so that `ehgr_launch_count()` is an honest count.
inline int launch_status() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? EHGR_OK : static_cast<int>(e);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize, set once per kernel (not on every launch: the call costs
// host time and must not sit inside a CUDA-graph capture more often than needed); grows monotonically.
void ensure_dyn_smem(const void* func, int bytes);
template <typename K>
inline void ensure_smem(K kernel, size_t bytes) { ensure_dyn_smem(reinterpret_cast<const void*>(kernel), static_cast<int>(bytes)); }

inline bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
inline int esize_of(int dtype) { return dtype == EHGR_F32 ? 4 : dtype == EHGR_BF16 ? 2 : 0; }
inline cudaStream_t as_stream(ehgr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;  // B200

// ---- streaming (read-once / write-once) vector access -------------------------------------
template <int BYTES>
struct Vec;
template <>
struct Vec<16> { using type = uint4; };
template <>
struct Vec<8> { using type = uint2; };
template <>
struct Vec<4> { using type = uint32_t; };
template <>
struct Vec<2> { using type = uint16_t; };

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_stream(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint16_t ld_stream(const uint16_t* p) {
  uint16_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(uint4* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream(uint2* p, uint2 v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream(uint32_t* p, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream(uint16_t* p, uint16_t v) {
  asm volatile("st.global.L1::no_allocate.u16 [%0], %1;" ::"l"(p), "h"(v) : "memory");
}
__device__ __forceinline__ uint4 zero_of(uint4) { return make_uint4(0, 0, 0, 0); }
__device__ __forceinline__ uint2 zero_of(uint2) { return make_uint2(0, 0); }
__device__ __forceinline__ uint32_t zero_of(uint32_t) { return 0u; }
__device__ __forceinline__ uint16_t zero_of(uint16_t) { return 0; }

// ---- warp reductions -----------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace ehgr
