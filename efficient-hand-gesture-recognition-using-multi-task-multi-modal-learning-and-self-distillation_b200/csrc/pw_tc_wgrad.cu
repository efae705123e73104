// pw_tc_wgrad.cu — weight gradient of the pointwise convolution on tcgen05 tensor cores.
//
//   dw[N,K] += sum_m rowop(dy)[m,n] * rowop(a)[m,k]
//
//   D[128 n x BKc k] (TMEM, fp32) accumulates over ALL row tiles a CTA owns: no per-tile epilogue.
//   Both operands are MN-major (the reduction index is the row m): a lane's 16-byte vector of 8
//   consecutive channels of row m lands at (m%8)*16 + (m/8)*128 + (channel/8)*gs — plain vector stores,
//   no transposition anywhere; the same bytes a forward GEMM would stage, read through a different
//   descriptor.  A ring stage holds kMS = 32 rows (two K=16 MMA steps); grid = (n_tiles*k_tiles) x
//   splits, each split strides over the row chunks; the epilogue adds the partial tile to dw with fp32
//   atomics.  Warps: 0-6 producers (one ring stage each, lane = row, batched fetch), 7 MMA issuer,
//   0-3 double as the epilogue at the end.
#include <type_traits>

#include <cstring>

#include "tma.cuh"

namespace ehgr {
namespace tc {

constexpr int kMS = 32;                 // rows (reduction elements) per ring stage
constexpr int kWgProducers = 7;
constexpr int kWgMmaWarp = 7;
constexpr int kWgThreads = 256;
constexpr int kWgMaxStages = 12;
constexpr int kWgBarBytes = 256 + 7 * 128 + 128;  // (+ the kTma "box has landed" barriers)
//constexpr int kWgBarBytesOld = 256 + 7 * 128;        // mbarriers + TMEM slot, then one scratch word per stage row and producer warp
                                                  // (SHIFT operand with an odd fold, see wg_copy)
constexpr int kWgGroupStride = kMS * 16 + 16;   // bytes between 8-channel groups of a stage: padded by one 16-byte slot so
                                                // that lanes walking along the groups hit different banks (the
                                                // unpadded 512-byte stride made every 16-byte access an 8-way conflict)
constexpr int kWgBudget = 100 * 1024;           // two CTAs per SM: 16 warps instead of 8 hide the copy / transform latency

struct WgradArgs {
  RowOp dy, a;
  float* dw;
  long long M;
  int K, N;
  int BKc;        // k columns per output tile (multiple of 16, <= 256)
  int k_tiles, n_tiles, splits;
  int n_tile;     // n rows per output tile (<= 128, multiple of 8): N split EVENLY over the n tiles (144 -> 72 + 72, not 128 + 16)
  long long m_chunks;   // ceil(M / kMS)
  int tmem_cols;
  int n_stages, stage_bytes;
  int dy_blocks;  // kTma: 64-channel blocks of the dy part of a stage
  int stage_rows; // kTma: reduction rows of a stage (32; CONV3 by TMA: the box's pixels rounded up to 16 — the rest stays zero)
  int box_px, cv_hb, cv_nb, cv_tpf;   // CONV3 by TMA: a stage = one box of cv_nb frames x cv_hb image rows x the full width
  int a_sw64;     // CONV3 by TMA with 32-channel pixels: operand a in 32-channel blocks of 64-byte rows (SWIZZLE_64B)
};

// Asynchronous staging of one operand of a ring stage (32 rows x `groups` 8-channel groups): lane owns a
// fixed channel group g = lane % G (G = power of two >= groups) and the rows rsub, rsub + 32/G, ...;
// every 16-byte chunk is one cp.async straight into the MN-major core-matrix layout.  SHIFT is a gather
// (neighbouring frame or zero fill); AFFINE is applied in place afterwards by the lane that copied the
// chunk (coefficients of its fixed group live in registers for the whole kernel).
struct WgLane {
  int g, rsub, rstep;     // channel group, first row, row step
  int cls;                // SHIFT class of the group: 0 -> frame t+dir, 1 -> t-dir, 2 -> same frame, 3 -> straddles
  bool on;                // g < groups
};

template <int kMode>
__device__ __forceinline__ WgLane wg_lane(const RowOp& op, int c_base, int groups, int lane) {
  WgLane w;
  int lg = 0;
  while ((1 << lg) < groups) ++lg;
  w.g = lane & ((1 << lg) - 1);
  w.rsub = lane >> lg;
  w.rstep = 32 >> lg;
  w.on = w.g < groups;
  w.cls = 2;
  if (kMode == EHGR_ROW_SHIFT && w.on) {
    const int c = c_base + w.g * 8, fold = op.fold;
    const int cl = c < fold ? 0 : (c < 2 * fold ? 1 : 2);
    const int ch = c + 7 < fold ? 0 : (c + 7 < 2 * fold ? 1 : 2);
    w.cls = cl == ch ? cl : 3;
  }
  return w;
}

template <int kMode>
__device__ __forceinline__ void wg_copy(const RowOp& op, const WgLane& w, int c_base, int C, uint32_t dst_base, int gs,
                                        long long m_base, long long M, uint32_t scr32 = 0) {
  if (!w.on) return;
  const __nv_bfloat16* in1 = static_cast<const __nv_bfloat16*>(op.in1);
  const int c = c_base + w.g * 8;
  if constexpr (kMode == EHGR_ROW_CONV3) {
    // im2col gather (see rowop.cuh): the lane's column c = (tap, channel) is fixed, its rows walk the output grid
    const Conv3Tap tp = conv3_tap(op, c);
    const int Wo = op.cv_w, Ho = op.cv_h;
    const int mfirst = static_cast<int>(m_base) + w.rsub;
    int fr = mfirst / op.hw;
    const int rem = mfirst - fr * op.hw;
    int ho = rem / Wo, wo = rem - ho * Wo, m = mfirst;
    const int Mi = static_cast<int>(M);
#pragma unroll 2
    for (int row = w.rsub; row < kMS; row += w.rstep) {
      const int hs = ho + tp.dy, ws = wo + tp.dx;
      const bool live = m < Mi && static_cast<unsigned>(hs) < static_cast<unsigned>(Ho) &&
                        static_cast<unsigned>(ws) < static_cast<unsigned>(Wo);
      const uint32_t dst = dst_base + w.g * gs + (row >> 3) * 128 + (row & 7) * 16;
      cp_async16(dst, live ? in1 + conv3_src(op, fr, hs, ws) + tp.c : in1, live ? 16u : 0u);
      m += w.rstep;
      wo += w.rstep;
      while (wo >= Wo) { wo -= Wo; ++ho; }
      while (ho >= Ho) { ho -= Ho; ++fr; }
    }
    return;
  }
  int t0 = 0, rem0 = 0;
  const int dir = op.shift_dir < 0 ? -1 : 1;
  if (kMode == EHGR_ROW_SHIFT && w.cls != 2) {
    const long long f0 = m_base / op.hw;
    rem0 = static_cast<int>(m_base - f0 * op.hw);
    t0 = static_cast<int>(f0 % op.n_segment);
  }
  const long long step = static_cast<long long>(dir) * op.hw * C;
#pragma unroll 4
  for (int row = w.rsub; row < kMS; row += w.rstep) {
    const long long m = m_base + row;
    bool live = m < M;
    const uint32_t dst = dst_base + w.g * gs + (row >> 3) * 128 + (row & 7) * 16;
    const __nv_bfloat16* src = in1 + m * C + c;
    if (kMode == EHGR_ROW_SHIFT && w.cls != 2 && live) {
      int rem = rem0 + row, t = t0;
      while (rem >= op.hw) { rem -= op.hw; t = t + 1 == op.n_segment ? 0 : t + 1; }
      if (w.cls == 3) {
        // straddles a fold boundary: one asynchronous 4-byte copy per channel pair from the frame its class reads; a
        // pair split by an odd fold sends its upper channel to the scratch word (patched in by wg_shift_patch)
        const int fold = op.fold;
        const bool ok0 = t + dir >= 0 && t + dir < op.n_segment, ok1 = t - dir >= 0 && t - dir < op.n_segment;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int cq = c + 2 * q;
          const int c_lo = cq < fold ? 0 : (cq < 2 * fold ? 1 : 2), c_hi = cq + 1 < fold ? 0 : (cq + 1 < 2 * fold ? 1 : 2);
          const __nv_bfloat16* sq = src + 2 * q;
          const bool lv = c_lo == 2 || (c_lo == 0 ? ok0 : ok1);
          cp_async4(dst + 4 * q, lv ? sq + (c_lo == 0 ? step : c_lo == 1 ? -step : 0) : in1, lv ? 4u : 0u);
          if (c_hi != c_lo) {
            const bool lh = c_hi == 2 || (c_hi == 0 ? ok0 : ok1);
            cp_async4(scr32 + static_cast<uint32_t>(row) * 4u, lh ? sq + (c_hi == 0 ? step : c_hi == 1 ? -step : 0) : in1,
                      lh ? 4u : 0u);
          }
        }
        continue;
      }
      const int tt = w.cls == 0 ? t + dir : t - dir;
      live = tt >= 0 && tt < op.n_segment;
      src += w.cls == 0 ? step : -step;
    }
    cp_async16(dst, live ? src : in1, live ? 16u : 0u);
  }
}

// SHIFT, odd fold: channel `fold` of the pair (fold-1, fold) comes from the scratch word of its row
__device__ __forceinline__ void wg_shift_patch(const RowOp& op, const WgLane& w, int c_base, uint32_t dst_base, int gs,
                                               long long m_base, long long M, uint32_t scr32) {
  const int fold = op.fold;
  if (!w.on || !(fold & 1) || w.cls != 3) return;
  const int c = c_base + w.g * 8;
  if (!(c <= fold - 1 && fold < c + 8)) return;           // this lane's vector does not hold the split pair
  const uint32_t e = static_cast<uint32_t>(fold - c) * 2u;
#pragma unroll 4
  for (int row = w.rsub; row < kMS; row += w.rstep) {
    if (m_base + row < M) {
      uint16_t v;
      asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(scr32 + static_cast<uint32_t>(row) * 4u + 2u));
      asm volatile("st.shared.u16 [%0], %1;" ::"r"(dst_base + w.g * gs + (row >> 3) * 128 + (row & 7) * 16 + e), "h"(v) : "memory");
    }
  }
}

__device__ __forceinline__ void wg_affine_inplace(const RowOp& op, const WgLane& w,
                                                  const RowLoader<__nv_bfloat16, 8, false, false>& ld, uint32_t dst_base,
                                                  int gs, long long m_base, long long M) {
  if (!w.on) return;
#pragma unroll 4
  for (int row = w.rsub; row < kMS; row += w.rstep) {
    if (m_base + row < M) {
      const uint32_t dst = dst_base + w.g * gs + (row >> 3) * 128 + (row & 7) * 16;
      RowLoader<__nv_bfloat16, 8, false, false>::Raw raw;
      raw.a = lds128(dst);
      RowOp oc = op;
      oc.mode = EHGR_ROW_AFFINE;          // only called for AFFINE: a constant for the loader
      sts128(dst, ld.finish_packed(oc, raw));
    }
  }
}

// CONV3 with a lazy BatchNorm+activation: transform the live (in-grid) pixels of the lane's column in place
__device__ __forceinline__ void wg_conv3_affine_inplace(const RowOp& op, const WgLane& w, int c_base,
                                                        const RowLoader<__nv_bfloat16, 8, false, false>& ld, uint32_t dst_base,
                                                        int gs, long long m_base, long long M) {
  if (!w.on) return;
  const Conv3Tap tp = conv3_tap(op, c_base + w.g * 8);
  const int Wo = op.cv_w, Ho = op.cv_h;
  const int mfirst = static_cast<int>(m_base) + w.rsub;
  const int rem = mfirst % op.hw;
  int ho = rem / Wo, wo = rem - ho * Wo, m = mfirst;
  const int Mi = static_cast<int>(M);
  RowOp oc = op;
  oc.mode = EHGR_ROW_AFFINE;
#pragma unroll 2
  for (int row = w.rsub; row < kMS; row += w.rstep) {
    if (m < Mi && static_cast<unsigned>(ho + tp.dy) < static_cast<unsigned>(Ho) &&
        static_cast<unsigned>(wo + tp.dx) < static_cast<unsigned>(Wo)) {
      const uint32_t dst = dst_base + w.g * gs + (row >> 3) * 128 + (row & 7) * 16;
      RowLoader<__nv_bfloat16, 8, false, false>::Raw raw;
      raw.a = lds128(dst);
      sts128(dst, ld.finish_packed(oc, raw));
    }
    m += w.rstep;
    wo += w.rstep;
    while (wo >= Wo) { wo -= Wo; ++ho; }
    while (ho >= Ho) ho -= Ho;
  }
}

// kAsync: dy is PLAIN and a is PLAIN / AFFINE / SHIFT / CONV3 (what the fused chain issues); otherwise the register path.
// kAMode: mode of operand a as a compile-time constant (PLAIN / AFFINE / SHIFT: asynchronous path, dy PLAIN);
// -1: the generic register path.
// kTma (dy PLAIN, a PLAIN / AFFINE): both operands arrive as TMA boxes of 32 rows x 64 channels in the SWIZZLE_128B
//       layout — MN-major canonical form ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: a 64-channel block is 32 rows x
//       128 bytes (4 KB, LBO = 4096 to the next block), 8-row groups are 1 KB apart (SBO = 1024), a K=16 step = 2 KB.
//       One thread issues <= 6 boxes per stage; AFFINE transforms a's chunks in place (lane = one 8-channel column of
//       the tile for the whole kernel, coefficients in registers).
template <int kAMode, bool kTma>
__global__ void __launch_bounds__(kWgThreads, 2)
pw_wgrad_tc_kernel(const __grid_constant__ WgradArgs p, const __grid_constant__ CUtensorMap tm_dy,
                   const __grid_constant__ CUtensorMap tm_a) {
  constexpr bool kAsync = kAMode >= 0;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = kTma ? smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) : smem_raw;
  const int gs = kTma ? p.stage_rows * 128 : kWgGroupStride;   // bytes between channel groups (kTma: 64-channel blocks) of a stage
  const int dy_bytes = kTma ? p.dy_blocks * gs : 16 * gs;      // 128 n = 16 groups
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.n_stages * p.stage_bytes);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kWgMaxStages), bar_done = smem_u32(bars + 2 * kWgMaxStages);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWgMaxStages + 1);
  [[maybe_unused]] const uint32_t bar_landed = smem_u32(reinterpret_cast<uint8_t*>(bars) + 256 + 7 * 128);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);          // one arrival per producer warp (lane 0, after __syncwarp)
      mbar_init(bar_empty + 8 * s, 1);
      if (kTma) mbar_init(bar_landed + 8 * s, 1);
    }
    if (kTma) { tma::prefetch_map(&tm_dy); tma::prefetch_map(&tm_a); }
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  // zero the operand ring once: padded channel groups are never written again
  for (int i = threadIdx.x; i < p.n_stages * p.stage_bytes / 16; i += kWgThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == kWgMmaWarp) tmem_alloc(smem_u32(tmem_slot), static_cast<uint32_t>(p.tmem_cols));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles = p.n_tiles * p.k_tiles;
  const int tile = blockIdx.x % tiles, split = blockIdx.x / tiles;
  const int n0 = (tile / p.k_tiles) * p.n_tile, k0 = (tile % p.k_tiles) * p.BKc;
  const int n_valid = min(p.n_tile, p.N - n0), k_valid = min(p.BKc, p.K - k0);   // multiples of 8
  const int ng = n_valid >> 3, kg = k_valid >> 3;
  long long my_chunks = 0;
  if (split < p.m_chunks) my_chunks = (p.m_chunks - split + p.splits - 1) / p.splits;
  const int pw = p.n_stages < kWgProducers ? p.n_stages : kWgProducers;   // see pw_tc.cu: parity aliasing

  if (warp < pw) {
    if constexpr (kTma) {
      const int a_blocks = (k_valid + 63) >> 6, dy_blocks = (n_valid + 63) >> 6;
      const uint32_t box_bytes = static_cast<uint32_t>(kAMode == EHGR_ROW_CONV3 ? p.box_px : kMS) * 128u;
      const int a_blk = p.a_sw64 ? p.stage_rows * 64 : gs;          // bytes between operand a's channel blocks
      const int a_nblk = p.a_sw64 ? (k_valid + 31) >> 5 : a_blocks;
      const uint32_t bytes_dy = static_cast<uint32_t>(dy_blocks) * box_bytes;
      const uint32_t bytes_a = p.a_sw64 ? static_cast<uint32_t>(a_nblk * p.box_px) * 64u : static_cast<uint32_t>(a_blocks) * box_bytes;
      // AFFINE: lane = 8-channel column `lane` of the tile (block lane >> 3, chunk lane & 7), all 32 rows of a stage
      RowLoader<__nv_bfloat16, 8, false, false> ld_a;
      RowOp ac = p.a;
      ac.mode = EHGR_ROW_AFFINE;
      const bool col_on = kAMode == EHGR_ROW_AFFINE && lane * 8 < k_valid;
      if (col_on) ld_a.init(ac, k0 + lane * 8, p.K);
      int s = 0, turn = 0;
      uint32_t ph = 0;
      for (long long mc = split; mc < p.m_chunks; mc += p.splits, ++s, ++turn) {
        if (turn == pw) turn = 0;
        if (s == p.n_stages) { s = 0; ph ^= 1; }
        if (turn != warp) continue;
        const uint32_t dy_dst = smem_base + s * p.stage_bytes, a_dst = dy_dst + dy_bytes;
        const int row0 = static_cast<int>(mc * kMS);
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        if constexpr (kAMode == EHGR_ROW_CONV3) {
          // im2col by TMA: the stage's pixels are one box (cv_nb frames x cv_hb image rows x the full width): dy's rows
          // are that contiguous pixel range, a's 64-channel blocks are the same box shifted by each block's tap (the zero
          // padding is the out-of-bounds fill); rows past the box stay zero from the ring's initialisation
          if (lane == 0) {
            const int ft = static_cast<int>(mc / p.cv_tpf), h0 = static_cast<int>(mc - static_cast<long long>(ft) * p.cv_tpf) * p.cv_hb;
            const int nfr = ft * p.cv_nb;
            const int px0 = (nfr * p.a.cv_h + h0) * p.a.cv_w;
            tma::expect_tx(bar_full + 8 * s, bytes_dy + bytes_a);
            for (int b = 0; b < dy_blocks; ++b) tma::load_2d(dy_dst + b * gs, &tm_dy, bar_full + 8 * s, n0 + b * 64, px0);
            for (int b = 0; b < a_nblk; ++b) {
              const int k = k0 + b * (p.a_sw64 ? 32 : 64), tap = k / p.a.cv_cin, ty = tap / 3;
              tma::load_4d(a_dst + b * a_blk, &tm_a, bar_full + 8 * s, k - tap * p.a.cv_cin, tap - ty * 3 - 1, h0 + ty - 1, nfr);
            }
          }
          continue;
        }
        if (lane == 0) {
          if (kAMode == EHGR_ROW_PLAIN) {
            tma::expect_tx(bar_full + 8 * s, bytes_dy + bytes_a);      // the one arrival; every box completes its bytes here
          } else {
            tma::expect_tx_only(bar_full + 8 * s, bytes_dy);
            tma::expect_tx(bar_landed + 8 * s, bytes_a);
          }
          for (int b = 0; b < dy_blocks; ++b) tma::load_2d(dy_dst + b * 4096, &tm_dy, bar_full + 8 * s, n0 + b * 64, row0);
          const uint32_t bar_a = kAMode == EHGR_ROW_PLAIN ? bar_full + 8 * s : bar_landed + 8 * s;
          for (int b = 0; b < a_blocks; ++b) tma::load_2d(a_dst + b * 4096, &tm_a, bar_a, k0 + b * 64, row0);
        }
        if (kAMode != EHGR_ROW_PLAIN) {
          mbar_wait(bar_landed + 8 * s, ph);
          if (col_on) {
            const uint32_t base = a_dst + (lane >> 3) * 4096;
            const int c16 = lane & 7;
#pragma unroll 4
            for (int row = 0; row < kMS; ++row) {
              if (row0 + row < p.M) {
                const uint32_t dst = base + row * 128 + ((c16 ^ (row & 7)) << 4);
                RowLoader<__nv_bfloat16, 8, false, false>::Raw raw;
                raw.a = lds128(dst);
                sts128(dst, ld_a.finish_packed(ac, raw));
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_full + 8 * s);
        }
      }
    } else if constexpr (kAsync) {
      const WgLane wl_dy = wg_lane<EHGR_ROW_PLAIN>(p.dy, n0, ng, lane), wl_a = wg_lane<kAMode>(p.a, k0, kg, lane);
      RowLoader<__nv_bfloat16, 8, false, false> ld_a;
      if (kAMode == EHGR_ROW_AFFINE && wl_a.on) {
        RowOp ac = p.a;
        ac.mode = EHGR_ROW_AFFINE;
        ld_a.init(ac, k0 + wl_a.g * 8, p.K);
      }
      if (kAMode == EHGR_ROW_CONV3 && wl_a.on && p.a.scale) {
        RowOp ac = p.a;
        ac.mode = EHGR_ROW_AFFINE;
        ld_a.init(ac, conv3_tap(p.a, k0 + wl_a.g * 8).c, p.a.cv_cin);
      }
      int s = 0, turn = 0;
      uint32_t ph = 0;
      const uint32_t scr32 = smem_u32(reinterpret_cast<uint8_t*>(bars) + 256) + static_cast<uint32_t>(warp) * 128u;
      for (long long mc = split; mc < p.m_chunks; mc += p.splits, ++s, ++turn) {
        if (turn == pw) turn = 0;
        if (s == p.n_stages) { s = 0; ph ^= 1; }
        if (turn != warp) continue;
        const uint32_t dy_dst = smem_base + s * p.stage_bytes, a_dst = dy_dst + dy_bytes;
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        wg_copy<EHGR_ROW_PLAIN>(p.dy, wl_dy, n0, p.N, dy_dst, gs, mc * kMS, p.M);
        wg_copy<kAMode>(p.a, wl_a, k0, p.K, a_dst, gs, mc * kMS, p.M, scr32);
        cp_async_wait_all();
        if (kAMode == EHGR_ROW_SHIFT) wg_shift_patch(p.a, wl_a, k0, a_dst, gs, mc * kMS, p.M, scr32);
        if (kAMode == EHGR_ROW_AFFINE) wg_affine_inplace(p.a, wl_a, ld_a, a_dst, gs, mc * kMS, p.M);
        if (kAMode == EHGR_ROW_CONV3 && p.a.scale) wg_conv3_affine_inplace(p.a, wl_a, k0, ld_a, a_dst, gs, mc * kMS, p.M);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * s);
      }
    } else {
    using Ld = RowLoader<__nv_bfloat16, 8, true, true>;
      uint32_t it = 0;
      for (long long mc = split; mc < p.m_chunks; mc += p.splits, ++it) {
        if (static_cast<int>(it % static_cast<uint32_t>(pw)) != warp) continue;
        const int s = it % p.n_stages;
        uint8_t* dy_dst = smem + s * p.stage_bytes;
        uint8_t* a_dst = dy_dst + dy_bytes;
        const long long m = mc * kMS + lane;               // lane = row of the stage
        const bool row_ok = m < p.M;
        const int row_off = (lane >> 3) * 128 + (lane & 7) * 16;
        bool waited = false;
        const int total = ng + kg;                          // channel groups of dy, then of a
  #pragma unroll 1
        for (int g0 = 0; g0 < total; g0 += 4) {
          Ld::Raw raw[4];
  #pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int g = g0 + j;
            if (g < total && row_ok) {
              Ld ld;
              const bool is_a = g >= ng;
              ld.c0 = is_a ? k0 + (g - ng) * 8 : n0 + g * 8;
              ld.C = is_a ? p.K : p.N;
              raw[j] = ld.fetch(is_a ? p.a : p.dy, m);
            }
          }
          if (!waited) { mbar_wait(bar_empty + 8 * s, ((it / p.n_stages) & 1) ^ 1); waited = true; }
  #pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int g = g0 + j;
            if (g < total) {
              const bool is_a = g >= ng;
              uint4 packed = make_uint4(0, 0, 0, 0);
              if (row_ok) {
                const RowOp& op = is_a ? p.a : p.dy;
                if (op.mode == EHGR_ROW_PLAIN) {
                  packed = raw[j].a;                        // already bf16: a straight 16-byte copy
                } else {
                  Ld ld;
                  float f[8];
                  ld.init(op, is_a ? k0 + (g - ng) * 8 : n0 + g * 8, is_a ? p.K : p.N);
                  ld.finish(op, raw[j], f);
                  packed = pack8(f);
                }
              }
              uint8_t* dst = is_a ? a_dst + (g - ng) * gs : dy_dst + g * gs;
              *reinterpret_cast<uint4*>(dst + row_off) = packed;
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * s);
      }
    }
  } else if (warp == kWgMmaWarp && lane == 0 && my_chunks > 0) {
    const uint32_t idesc = make_idesc(128, p.BKc, 1, 1);     // both operands MN-major
    uint32_t it = 0, ph = 0;
    int s = 0;
    for (long long mc = split; mc < p.m_chunks; mc += p.splits, ++it, ++s) {
      if (s == p.n_stages) { s = 0; ph ^= 1; }
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after();
      const uint32_t dy_addr = smem_base + s * p.stage_bytes;
      const uint32_t a_addr = dy_addr + dy_bytes;
      const int ksteps = kTma ? p.stage_rows >> 4 : kMS / 16;
      const uint32_t blk = static_cast<uint32_t>(gs);
#pragma unroll 4
      for (int kk = 0; kk < ksteps; ++kk) {
        // 16 rows of m = two 8-row core matrices = 256 bytes; LBO (m groups) = 128, SBO (channel groups) = gs
        // no swizzle: 16 rows of m = two 8-row core matrices = 256 bytes, LBO (m groups) = 128, SBO (channel groups) = gs;
        // kTma: SWIZZLE_128B, a K=16 step = two 1 KB row groups, LBO = 4096 (64-channel blocks), SBO = 1024
        const uint64_t da = kTma ? (make_desc(dy_addr + kk * 2048, blk, 1024) | (2ull << 61)) : make_desc(dy_addr + kk * 256, 128, gs);
        // (operand a with 32-channel pixels: SWIZZLE_64B — 64-byte rows, 8-row groups 512 bytes apart, a K=16 step = 1 KB)
        const uint64_t db = !kTma ? make_desc(a_addr + kk * 256, 128, gs)
                            : p.a_sw64 ? (make_desc(a_addr + kk * 1024, static_cast<uint32_t>(p.stage_rows) * 64u, 512) | (4ull << 61))
                                       : (make_desc(a_addr + kk * 2048, blk, 1024) | (2ull << 61));
        umma_bf16(tmem_base, da, db, idesc, (it | kk) ? 1u : 0u);
      }
      umma_commit(bar_empty + 8 * s);
    }
    umma_commit(bar_done);
  }
  // ---- epilogue (warps 0-3 once their production is done): TMEM -> fp32 atomics into dw[N,K]
  if (warp < 4 && my_chunks > 0) {
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int n = n0 + q * 32 + lane;
    const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int cc = 0; cc * 16 < k_valid; ++cc) {
      float v[16];
      tmem_ld16(t_base + cc * 16, v);
      if (q * 32 + lane < n_valid) {
        float* dst = p.dw + static_cast<size_t>(n) * p.K + k0 + cc * 16;
        // k_valid is a multiple of 8: whole 16-byte groups, one vector reduction each (4x fewer L2 atomics)
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          if (cc * 16 + i < k_valid)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(v[i]), "f"(v[i + 1]),
                         "f"(v[i + 2]), "f"(v[i + 3])
                         : "memory");
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == kWgMmaWarp) {
    __syncwarp();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

}  // namespace tc

bool pw_wgrad_tc_supported(const RowOp& dy, const RowOp& a, long long M, int K, int N, int dtype) {
  if (dtype != EHGR_BF16) return false;
  if (K % 8 || N % 8 || K < 8 || N < 8) return false;
  if (M < 1) return false;
  if (a.mode == EHGR_ROW_CONV3 && dy.mode != EHGR_ROW_PLAIN) return false;   // the register path has no gather
  if (dy.mode == EHGR_ROW_CONV3) return false;
  return true;
}

int pw_wgrad_tc(const RowOp& dy, const RowOp& a, float* dw, long long M, int K, int N, cudaStream_t s) {
  tc::WgradArgs p;
  p.dy = dy; p.a = a; p.dw = dw;
  p.M = M; p.K = K; p.N = N;
  p.BKc = tc::pick_bn(K);
  p.k_tiles = (K + p.BKc - 1) / p.BKc;
  p.n_tiles = (N + 127) / 128;
  p.n_tile = ((N + p.n_tiles - 1) / p.n_tiles + 7) & ~7;
  p.m_chunks = cdiv(M, tc::kMS);
  int tiles = p.n_tiles * p.k_tiles;
  p.splits = static_cast<int>(std::max<long long>(1, std::min<long long>(p.m_chunks, 2 * kNumSMs / tiles)));
  constexpr int kBudget = tc::kWgBudget;
  // TMA + SWIZZLE_128B operand path: both operands read one tensor row by row
  bool use_tma = dy.mode == EHGR_ROW_PLAIN && (a.mode == EHGR_ROW_PLAIN || a.mode == EHGR_ROW_AFFINE) && M < 0x7fffffffLL;
  p.stage_rows = tc::kMS;
  p.box_px = tc::kMS;
  p.cv_hb = p.cv_nb = p.cv_tpf = 1;
  p.a_sw64 = 0;
  // CONV3 im2col by TMA: plain operand at the output resolution with 64-channel pixels; a stage = one box of <= 64 pixels
  bool conv_tma = false;
  if (dy.mode == EHGR_ROW_PLAIN && a.mode == EHGR_ROW_CONV3 && !a.scale && !a.cv_up && (a.cv_cin % 64 == 0 || a.cv_cin == 32) &&
      a.cv_w <= 64) {
    p.a_sw64 = a.cv_cin == 32 ? 1 : 0;
    const int hw = a.cv_h * a.cv_w;
    if (hw <= 64) {
      p.cv_hb = a.cv_h;
      p.cv_nb = 64 / hw;
      p.cv_tpf = 1;
      conv_tma = true;
    } else {
      for (int hb = 64 / a.cv_w; hb >= 1; --hb)
        if (a.cv_h % hb == 0) { p.cv_hb = hb; conv_tma = true; break; }
      p.cv_nb = 1;
      p.cv_tpf = a.cv_h / p.cv_hb;
    }
    if (conv_tma) {
      p.BKc = std::min(256, K);                       // whole 64- (32-) channel blocks: a block never straddles a tap
      p.k_tiles = (K + p.BKc - 1) / p.BKc;
      tiles = p.n_tiles * p.k_tiles;
      p.box_px = p.cv_nb * p.cv_hb * a.cv_w;
      p.stage_rows = (p.box_px + 15) & ~15;
      p.m_chunks = cdiv(M / hw, p.cv_nb) * p.cv_tpf;
      p.splits = static_cast<int>(std::max<long long>(1, std::min<long long>(p.m_chunks, 2 * kNumSMs / tiles)));
      use_tma = true;
    }
  }
  int cols = 32;
  while (cols < p.BKc) cols <<= 1;
  p.tmem_cols = cols;
  p.dy_blocks = (p.n_tile + 63) / 64;
  p.stage_bytes = !use_tma ? (16 + p.BKc / 8) * tc::kWgGroupStride
                  : (conv_tma && p.a_sw64) ? (p.dy_blocks * 128 + (p.BKc + 31) / 32 * 64) * p.stage_rows
                                           : (p.dy_blocks + (p.BKc + 63) / 64) * p.stage_rows * 128;
  p.n_stages = std::max(2, std::min(tc::kWgMaxStages, (kBudget - tc::kWgBarBytes - (use_tma ? 1024 : 0)) / p.stage_bytes));
  const size_t smem = static_cast<size_t>(p.n_stages) * p.stage_bytes + tc::kWgBarBytes + (use_tma ? 1024 : 0);
  CUtensorMap tm_dy, tm_a;
  memset(&tm_dy, 0, sizeof(tm_dy));
  memset(&tm_a, 0, sizeof(tm_a));
  if (conv_tma) {
    if (int st = tma::make_map_2d_sw128(&tm_dy, dy.in1, static_cast<unsigned long long>(N), static_cast<unsigned long long>(M),
                                        static_cast<unsigned>(p.box_px)))
      return st;
    const unsigned long long cin = a.cv_cin, wo = a.cv_w, ho = a.cv_h, frames = M / a.hw;
    const unsigned long long dims[4] = {cin, wo, ho, frames};
    const unsigned long long strides[3] = {cin * 2, wo * cin * 2, ho * wo * cin * 2};
    const unsigned box[4] = {p.a_sw64 ? 32u : 64u, static_cast<unsigned>(wo), static_cast<unsigned>(p.cv_hb), static_cast<unsigned>(p.cv_nb)};
    if (int st = tma::make_map_4d(&tm_a, 2, a.in1, dims, strides, box, p.a_sw64 ? 64 : 128)) return st;
  } else if (use_tma) {
    if (int st = tma::make_map_2d_sw128(&tm_dy, dy.in1, static_cast<unsigned long long>(N), static_cast<unsigned long long>(M), tc::kMS)) return st;
    if (int st = tma::make_map_2d_sw128(&tm_a, a.in1, static_cast<unsigned long long>(K), static_cast<unsigned long long>(M), tc::kMS)) return st;
  }
  const bool async = dy.mode == EHGR_ROW_PLAIN && (a.mode == EHGR_ROW_PLAIN || a.mode == EHGR_ROW_AFFINE ||
                                                   a.mode == EHGR_ROW_SHIFT || a.mode == EHGR_ROW_CONV3);
  auto go = [&](auto mode_tag, auto tma_tag) {
    constexpr int kAMode = decltype(mode_tag)::value;
    constexpr bool kTma = decltype(tma_tag)::value;
    ensure_smem(tc::pw_wgrad_tc_kernel<kAMode, kTma>, kBudget);
    tc::pw_wgrad_tc_kernel<kAMode, kTma><<<static_cast<unsigned>(tiles * p.splits), tc::kWgThreads, smem, s>>>(p, tm_dy, tm_a);
  };
  using T = std::true_type;
  using F = std::false_type;
  if (!async) go(std::integral_constant<int, -1>{}, F{});
  else if (a.mode == EHGR_ROW_PLAIN) { if (use_tma) go(std::integral_constant<int, EHGR_ROW_PLAIN>{}, T{}); else go(std::integral_constant<int, EHGR_ROW_PLAIN>{}, F{}); }
  else if (a.mode == EHGR_ROW_AFFINE) { if (use_tma) go(std::integral_constant<int, EHGR_ROW_AFFINE>{}, T{}); else go(std::integral_constant<int, EHGR_ROW_AFFINE>{}, F{}); }
  else if (a.mode == EHGR_ROW_CONV3) { if (conv_tma) go(std::integral_constant<int, EHGR_ROW_CONV3>{}, T{}); else go(std::integral_constant<int, EHGR_ROW_CONV3>{}, F{}); }
  else go(std::integral_constant<int, EHGR_ROW_SHIFT>{}, F{});
  return launch_status();
}

}  // namespace ehgr
