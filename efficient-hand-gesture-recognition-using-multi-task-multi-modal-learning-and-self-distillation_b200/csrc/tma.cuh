// tma.cuh — Tensor Memory Accelerator plumbing: host-side tensor maps (cuTensorMapEncodeTiled, resolved
// through the runtime's driver entry point so that libcuda is not a link dependency) and the device-side
// bulk tensor copy.  The maps describe NHWC bf16 activations [NT, H, W, C]; a box is one spatial tile of
// one channel chunk, out-of-image coordinates (the conv halo, ragged channel chunks) are zero-filled by
// the hardware.
#pragma once
#include <cuda.h>

#include "tc_common.cuh"

namespace ehgr {
namespace tma {

// 0 on success.  box = (box_c channels, box_w, box_h, 1 frame); innermost box bytes must be a multiple of 16.
int make_nhwc_bf16_map(CUtensorMap* out, const void* base, int nt, int h, int w, int c, int box_c, int box_w, int box_h);

// generic 4-D map (dims / box innermost first, strides of dims 1..3 in bytes), bf16 (elem_bytes 2) or fp32 (4)
int make_map_4d(CUtensorMap* out, int elem_bytes, const void* base, const unsigned long long (&dims)[4],
                const unsigned long long (&strides_bytes)[3], const unsigned (&box)[4], int swizzle_bytes = 0);   // 0 | 64 | 128

// row-major bf16 matrix [rows, cols] -> boxes of box_rows x 64 columns in the SWIZZLE_128B shared-memory layout
int make_map_2d_sw128(CUtensorMap* out, const void* base, unsigned long long cols, unsigned long long rows, unsigned box_rows);

__device__ __forceinline__ void expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// coordinates innermost first: channel, x, y, frame
__device__ __forceinline__ void load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int x, int y, int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c), "r"(x), "r"(y), "r"(n)
      : "memory");
}
// transaction bytes only (no arrival): the same thread arrives later, once its other copies are complete
__device__ __forceinline__ void expect_tx_only(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 2-D box (coordinates innermost first: channel, row)
__device__ __forceinline__ void load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int row) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c), "r"(row)
               : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

}  // namespace tma
}  // namespace ehgr
