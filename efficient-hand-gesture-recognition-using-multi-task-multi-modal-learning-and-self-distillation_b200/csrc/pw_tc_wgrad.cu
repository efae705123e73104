// pw_tc_wgrad.cu — weight gradient of the pointwise convolution on tcgen05 tensor cores.
#include "tc_common.cuh"

namespace ehgr {
namespace tc {

// ---------------------------------------------------------------------------------------------------
// weight gradient:  dw[N,K] += sum_m dy[m,n] * a[m,k]
//   D[128 n x BKc k] (TMEM, fp32) accumulates over ALL row tiles a CTA owns: no per-tile epilogue.
//   Both operands are MN-major (the reduction index is the row m): a thread's 16-byte vector of 8
//   consecutive channels of row m lands at (m%8)*16 + (m/8)*128 + (channel/8)*2048 — plain vector
//   stores, no transposition anywhere.  grid = (n_tiles*k_tiles) x splits; each split strides over
//   the row tiles; the epilogue adds the partial tile to dw with fp32 atomics.
//   5 warps: 0-3 produce (and run the epilogue at the end), 4 issues the MMAs.
// ---------------------------------------------------------------------------------------------------
constexpr int kWgStages = 2;
constexpr int kWgThreads = 288;   // 2 producer groups x 4 warps + 1 MMA warp
constexpr int kWgDyBytes = 128 * 128 * 2;   // [128 n][128 m] bf16

struct WgradArgs {
  RowOp dy, a;
  float* dw;
  long long M;
  int K, N;
  int BKc;        // k columns per output tile (multiple of 16, <= 256)
  int k_tiles, n_tiles, m_tiles, splits;
  int tmem_cols;
};

__global__ void __launch_bounds__(kWgThreads, 1) pw_wgrad_tc_kernel(WgradArgs p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int a_bytes = p.BKc * 128 * 2;
  const int stage_bytes = kWgDyBytes + a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * stage_bytes);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kWgStages), bar_done = smem_u32(bars + 2 * kWgStages);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWgStages + 1);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(bar_full + 8 * s, 128);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  // zero the operand ring once: padded channel groups are never written again
  for (int i = threadIdx.x; i < kWgStages * stage_bytes / 16; i += kWgThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), static_cast<uint32_t>(p.tmem_cols));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles = p.n_tiles * p.k_tiles;
  const int tile = blockIdx.x % tiles, split = blockIdx.x / tiles;
  const int n0 = (tile / p.k_tiles) * 128, k0 = (tile % p.k_tiles) * p.BKc;
  const int n_valid = min(128, p.N - n0), k_valid = min(p.BKc, p.K - k0);   // multiples of 8
  const int ng = n_valid >> 3, kg = k_valid >> 3;
  int my_tiles = 0;
  for (int mt = split; mt < p.m_tiles; mt += p.splits) ++my_tiles;

  if (warp < 8) {
    // two producer groups (128 threads each) alternate ring stages; per stage a thread fetches its
    // vectors in batches of four (loads only), then applies the row operand and stores.
    const int tid = threadIdx.x & 127, group = warp >> 2;
    const int r = tid & 7;
    using Ld = RowLoader<__nv_bfloat16, 8>;
    uint32_t it = 0;
    for (int mt = split; mt < p.m_tiles; mt += p.splits, ++it) {
      if (static_cast<int>(it & 1) != group) continue;
      const int s = it % kWgStages;
      uint8_t* dy_dst = smem + s * stage_bytes;
      uint8_t* a_dst = dy_dst + kWgDyBytes;
      const long long m0 = static_cast<long long>(mt) * 128;
      bool waited = false;
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const RowOp& op = pass ? p.a : p.dy;
        const int groups = pass ? kg : ng, c_base = pass ? k0 : n0, C = pass ? p.K : p.N;
        uint8_t* dst = pass ? a_dst : dy_dst;
#pragma unroll 1
        for (int v0 = tid; v0 < 128 * groups; v0 += 4 * 128) {
          Ld::Raw raw[4];
          bool live[4];
          int off[4], c0[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int v = v0 + j * 128;
            off[j] = -1;
            live[j] = false;
            if (v < 128 * groups) {
              const int g = (v >> 3) % groups, mg = (v >> 3) / groups;
              const long long m = m0 + mg * 8 + r;
              off[j] = g * 2048 + mg * 128 + r * 16;
              c0[j] = c_base + g * 8;
              live[j] = m < p.M;
              if (live[j]) {
                Ld ld;
                ld.c0 = c0[j];
                ld.C = C;
                raw[j] = ld.fetch(op, m);
              }
            }
          }
          if (!waited) { mbar_wait(bar_empty + 8 * s, ((it / kWgStages) & 1) ^ 1); waited = true; }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (off[j] >= 0) {
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = 0.f;
              if (live[j]) {
                Ld ld;
                ld.init(op, c0[j], C);
                ld.finish(op, raw[j], f);
              }
              *reinterpret_cast<uint4*>(dst + off[j]) = pack8(f);
            }
          }
        }
      }
      if (!waited) mbar_wait(bar_empty + 8 * s, ((it / kWgStages) & 1) ^ 1);
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * s);
    }
  }
  if (warp < 4) {
    // ---- epilogue (same warps): TMEM -> fp32 atomics into dw[N,K]
    if (my_tiles > 0) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
      const int q = warp & 3;
      const int n = n0 + q * 32 + lane;
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      for (int cc = 0; cc * 16 < k_valid; ++cc) {
        float v[16];
        tmem_ld16(t_base + cc * 16, v);
        if (n < p.N) {
          float* dst = p.dw + static_cast<size_t>(n) * p.K + k0 + cc * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (cc * 16 + i < k_valid) atomicAdd(dst + i, v[i]);
        }
      }
      tc_fence_before();
    }
  } else if (warp == 8 && lane == 0 && my_tiles > 0) {
    const uint32_t idesc = make_idesc(128, p.BKc, 1, 1);     // both operands MN-major
    uint32_t it = 0;
    for (int mt = split; mt < p.m_tiles; mt += p.splits, ++it) {
      const int s = it % kWgStages;
      mbar_wait(bar_full + 8 * s, (it / kWgStages) & 1);
      tc_fence_after();
      const uint32_t dy_addr = smem_base + s * stage_bytes;
      const uint32_t a_addr = dy_addr + kWgDyBytes;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        // 16 rows of m = two 8-row core matrices = 256 bytes; LBO (k groups) = 128, SBO (channel groups) = 2048
        const uint64_t da = make_desc(dy_addr + kk * 256, 128, 2048);
        const uint64_t db = make_desc(a_addr + kk * 256, 128, 2048);
        umma_bf16(tmem_base, da, db, idesc, (it | kk) ? 1u : 0u);
      }
      umma_commit(bar_empty + 8 * s);
    }
    umma_commit(bar_done);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

}  // namespace tc

bool pw_wgrad_tc_supported(const RowOp& dy, const RowOp& a, long long M, int K, int N, int dtype) {
  (void)dy; (void)a;
  if (dtype != EHGR_BF16) return false;
  if (K % 8 || N % 8 || K < 8 || N < 8) return false;
  if (M < 1 || M / 128 > 0x3fffffff) return false;
  return true;
}

int pw_wgrad_tc(const RowOp& dy, const RowOp& a, float* dw, long long M, int K, int N, cudaStream_t s) {
  tc::WgradArgs p;
  p.dy = dy; p.a = a; p.dw = dw;
  p.M = M; p.K = K; p.N = N;
  p.BKc = tc::pick_bn(K);
  p.k_tiles = (K + p.BKc - 1) / p.BKc;
  p.n_tiles = (N + 127) / 128;
  p.m_tiles = static_cast<int>(cdiv(M, 128));
  const int tiles = p.n_tiles * p.k_tiles;
  p.splits = std::max(1, std::min(p.m_tiles, kNumSMs / tiles));
  int cols = 32;
  while (cols < p.BKc) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = static_cast<size_t>(tc::kWgStages) * (tc::kWgDyBytes + p.BKc * 256) + 128;
  cudaFuncSetAttribute(tc::pw_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  tc::pw_wgrad_tc_kernel<<<static_cast<unsigned>(tiles * p.splits), tc::kWgThreads, smem, s>>>(p);
  return launch_status();
}

}  // namespace ehgr
