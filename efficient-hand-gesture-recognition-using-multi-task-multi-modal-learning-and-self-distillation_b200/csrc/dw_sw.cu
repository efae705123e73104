// dw_sw.cu — K7 depthwise 3x3 for bf16 storage: asynchronous tiles + sliding-window taps + packed FP32 math.
//
// What bounded the first tiled kernels (dw_tiled.cu) was not HBM but latency and instruction issue: a CTA
// loaded its tile through registers (global latency exposed between two barriers) and spent ~17
// instructions per 8-channel tap.  Here
//   * a tile (+halo) of a channel chunk is copied RAW with 16-byte cp.async (no registers, the next
//     tile's copies are in flight while the current one is computed: forward double-buffers its tile),
//   * the row operand (BatchNorm+ReLU6 / BatchNorm-backward) is applied IN PLACE in shared memory by the
//     thread that copied the vector (FFMA2, ReLU folded into the bf16x2 conversion),
//   * a thread owns (channel vector, output column) and walks down the rows: every input row is read once
//     (3 vectors) and feeds the three output rows it touches, accumulators rotate through registers
//     (the row loop is fully unrolled, so the rotation is static),
//   * all FP32 arithmetic is on register pairs (FFMA2 / FADD2: two lanes per issue slot on sm_100).
// Tile geometry is a template parameter: no integer division anywhere in the loops.
//
//   CVN = 8  (64 channels / chunk), 16 columns of threads, output tile 14 wide   (maps >= 14 wide)
//   CVN = 16 (128 channels / chunk), 8 columns of threads, output tile  7 wide   (7x7 maps, 14 -> 7)
#include "tma.cuh"
#include "bnfin.cuh"

namespace ehgr {

using tc::cp_async16;
using tc::fence_mbar_init;
using tc::fence_proxy_async;
using tc::mbar_init;
using tc::mbar_wait;
using tc::lds128;
using tc::smem_u32;
using tc::sts128;

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct DwSw {
  int nt, h, w, c, ho, wo;
  int tiles_y, tiles_x, n_chunks;
  long long items;          // nt * tiles_y * tiles_x
};

template <int STRIDE, int CVN, int TW, int TH>
struct DwCfg {
  static constexpr int NCOL = 128 / CVN;                 // thread columns
  static constexpr int IW = (TW - 1) * STRIDE + 3;       // input tile (with halo)
  static constexpr int IH = (TH - 1) * STRIDE + 3;
  static constexpr int CC = CVN * 8;                     // channels per chunk
  static constexpr int TILE_BYTES = IH * IW * CVN * 16;
  static_assert(TW <= NCOL, "one thread column per output column");
};

__device__ __forceinline__ void unpack8(const uint4& r, float2 (&f)[4]) {
  f[0] = bf2_to_f2(r.x); f[1] = bf2_to_f2(r.y); f[2] = bf2_to_f2(r.z); f[3] = bf2_to_f2(r.w);
}
__device__ __forceinline__ uint4 pack8f2(const float2 (&f)[4]) {
  return make_uint4(pack_bf16x2(f[0].x, f[0].y), pack_bf16x2(f[1].x, f[1].y), pack_bf16x2(f[2].x, f[2].y),
                    pack_bf16x2(f[3].x, f[3].y));
}

// Issue the raw copies of one tile: rows [r0, r0+NR) x cols [c0w, c0w+NC) of frame nt of an [*, rh, rw, C]
// tensor, zero outside the image.  Thread (cv, col) copies column col, col+NCOL, ... of every row.
template <int NR, int NC, int CVN>
__device__ __forceinline__ void tile_copy(const __nv_bfloat16* __restrict__ src, uint32_t dst, long long nt, int rh, int rw,
                                          int C, int c0, int r0, int c0w, int cv, int col) {
  constexpr int NCOL = 128 / CVN;
#pragma unroll
  for (int ix0 = 0; ix0 < NC; ix0 += NCOL) {
    const int ix = ix0 + col;
    if (ix < NC) {
      const int ww = c0w + ix;
      const bool col_ok = ww >= 0 && ww < rw;
      const __nv_bfloat16* p = src + ((nt * rh + r0) * rw + (col_ok ? ww : 0)) * C + c0;
      uint32_t d = dst + static_cast<uint32_t>((ix * CVN + cv) * 16);
#pragma unroll 4
      for (int iy = 0; iy < NR; ++iy) {
        const int hh = r0 + iy;
        const bool ok = col_ok && hh >= 0 && hh < rh;
        cp_async16(d, ok ? p : src, ok ? 16u : 0u);
        p += static_cast<long long>(rw) * C;
        d += NC * CVN * 16;
      }
    }
  }
}

// In-place row operand on the vectors this thread copied (positions outside the image stay zero).
template <int NR, int NC, int CVN, typename Ld>
__device__ __forceinline__ void tile_transform(const RowOp& op, const Ld& ld, uint32_t dst, uint32_t dst2, int rh, int rw,
                                               int r0, int c0w, int cv, int col) {
  constexpr int NCOL = 128 / CVN;
#pragma unroll
  for (int ix0 = 0; ix0 < NC; ix0 += NCOL) {
    const int ix = ix0 + col;
    if (ix < NC) {
      const int ww = c0w + ix;
      if (ww >= 0 && ww < rw) {
        uint32_t off = static_cast<uint32_t>((ix * CVN + cv) * 16);
#pragma unroll 4
        for (int iy = 0; iy < NR; ++iy) {
          const int hh = r0 + iy;
          if (hh >= 0 && hh < rh) {
            typename Ld::Raw raw;
            raw.a = lds128(dst + off);
            raw.b[0] = lds128(dst2 + off);     // second tensor of a BNBWD operand (dst2 == dst otherwise)
            sts128(dst + off, ld.finish_packed(op, raw));
          }
          off += NC * CVN * 16;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Sliding-window 3x3 sweep over a staged tile: thread (cv, ox) walks the IH input rows, emits TH output
// rows through `emit(oy, acc)`.  w2[kh*3+kw] are the taps (flip them for the input-gradient form).
// ------------------------------------------------------------------------------------------------
template <int STRIDE, int CVN, int IW, int IH, int TH, typename Emit>
__device__ __forceinline__ void conv_sweep(uint32_t tile, int cv, int ox, const float2 (&w2)[9][4], Emit emit) {
  float2 acc[3][4];
  const uint32_t base = tile + static_cast<uint32_t>((ox * STRIDE * CVN + cv) * 16);
#pragma unroll
  for (int iy = 0; iy < IH; ++iy) {
    float2 x[3][4];
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) unpack8(lds128(base + static_cast<uint32_t>(((iy * IW + kw) * CVN) * 16)), x[kw]);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int t = iy - kh;                       // = oy * STRIDE
      if (t < 0 || (t % STRIDE) != 0 || t / STRIDE >= TH) continue;
      const int oy = t / STRIDE, slot = oy % 3;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          acc[slot][i] = (kh == 0 && kw == 0) ? __fmul2_rn(x[kw][i], w2[kh * 3 + kw][i])
                                               : __ffma2_rn(x[kw][i], w2[kh * 3 + kw][i], acc[slot][i]);
      if (kh == 2) emit(oy, acc[slot]);
    }
  }
}

template <int STRIDE, int CVN, int TW, int TH>
__global__ void __launch_bounds__(128, 3)
dw_fwd_sw_kernel(RowOp a, const __grid_constant__ CUtensorMap tm_a, const float* __restrict__ wgt,
                 __nv_bfloat16* __restrict__ out, double* __restrict__ stats, DwSw g, BnFin fin) {
  using Cfg = DwCfg<STRIDE, CVN, TW, TH>;
  using Ld = RowLoader<__nv_bfloat16, 8, false, false>;
  extern __shared__ __align__(128) uint8_t smem[];
  float* s_stat = reinterpret_cast<float*>(smem + 2 * Cfg::TILE_BYTES);   // [2][CC]
  const uint32_t tile0 = smem_u32(smem);
  const uint32_t bar0 = smem_u32(s_stat + 2 * Cfg::CC);                    // two mbarriers: one per tile buffer
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_mbar_init();
    tma::prefetch_map(&tm_a);
  }
  const int tid = threadIdx.x, cv = tid % CVN, col = tid / CVN;
  const int chunk = blockIdx.x % g.n_chunks;
  const int c0 = chunk * Cfg::CC + cv * 8;
  const bool cv_on = c0 < g.c;
  for (int i = tid; i < 2 * Cfg::CC; i += 128) s_stat[i] = 0.f;

  float2 w2[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // weights rounded to the storage type, like autocast's bf16 conv would
      const float w0 = cv_on ? round_to<__nv_bfloat16>(wgt[(c0 + 2 * i) * 9 + t]) : 0.f;
      const float w1 = cv_on ? round_to<__nv_bfloat16>(wgt[(c0 + 2 * i + 1) * 9 + t]) : 0.f;
      w2[t][i] = make_float2(w0, w1);
    }
  Ld ld;
  if (cv_on && a.mode == EHGR_ROW_AFFINE) {
    RowOp ac = a;
    ac.mode = EHGR_ROW_AFFINE;
    ld.init(ac, c0, g.c);
  }
  float2 tsum[4], tsq[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) tsum[i] = tsq[i] = make_float2(0.f, 0.f);

  const int c_base = chunk * Cfg::CC;
  const long long item_stride = gridDim.x / g.n_chunks;
  const long long first = blockIdx.x / g.n_chunks;
  // One TMA box = the whole input tile (+halo) of this channel chunk: [IH][IW][CC] bf16 lands exactly in the
  // layout the sweeps read; coordinates outside the image and channels beyond C are zero-filled.
  auto issue = [&](uint32_t dst, uint32_t bar, long long nt, int ho0, int wo0) {
    tma::expect_tx(bar, Cfg::TILE_BYTES);
    tma::load_4d(dst, &tm_a, bar, c_base, wo0 * STRIDE - 1, ho0 * STRIDE - 1, static_cast<int>(nt));
  };
  auto origin = [&](long long item, long long& nt, int& ho0, int& wo0) {
    const int tx = static_cast<int>(item % g.tiles_x);
    const long long r = item / g.tiles_x;
    ho0 = static_cast<int>(r % g.tiles_y) * TH;
    wo0 = tx * TW;
    nt = r / g.tiles_y;
  };
  __syncthreads();                                       // barriers initialised
  if (first < g.items && tid == 0) {
    long long nt; int ho0, wo0;
    origin(first, nt, ho0, wo0);
    issue(tile0, bar0, nt, ho0, wo0);
  }
  int buf = 0;
  uint32_t phase[2] = {0, 0};
  for (long long item = first; item < g.items; item += item_stride, buf ^= 1) {
    long long nt; int ho0, wo0;
    origin(item, nt, ho0, wo0);
    const uint32_t tile = tile0 + buf * Cfg::TILE_BYTES;
    if (item + item_stride < g.items && tid == 0) {      // prefetch the next tile into the other buffer
      long long nt2; int ho2, wo2;
      origin(item + item_stride, nt2, ho2, wo2);
      issue(tile0 + (buf ^ 1) * Cfg::TILE_BYTES, bar0 + 8 * (buf ^ 1), nt2, ho2, wo2);
    }
    mbar_wait(bar0 + 8 * buf, phase[buf]);               // this tile has landed (the next one may still be in flight)
    phase[buf] ^= 1;
    if (cv_on && a.mode != EHGR_ROW_PLAIN) {
      RowOp ac = a;
      ac.mode = EHGR_ROW_AFFINE;        // the only other mode the host lets through: a compile-time constant for the loader
      tile_transform<Cfg::IH, Cfg::IW, CVN, Ld>(ac, ld, tile, tile, g.h, g.w, ho0 * STRIDE - 1, wo0 * STRIDE - 1, cv, col);
    }
    __syncthreads();
    if (col < TW) {
      const int wo = wo0 + col;
      const bool st_ok = cv_on && wo < g.wo;
      __nv_bfloat16* orow = out + ((nt * g.ho + ho0) * g.wo + wo) * g.c + c0;
      conv_sweep<STRIDE, CVN, Cfg::IW, Cfg::IH, TH>(tile, cv, col, w2, [&](int oy, const float2 (&acc)[4]) {
        if (st_ok && ho0 + oy < g.ho) {
          *reinterpret_cast<uint4*>(orow + static_cast<long long>(oy) * g.wo * g.c) = pack8f2(acc);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            tsum[i] = __fadd2_rn(tsum[i], acc[i]);
            tsq[i] = __ffma2_rn(acc[i], acc[i], tsq[i]);
          }
        }
      });
    }
    fence_proxy_async();                                // generic-proxy accesses of this buffer precede the next TMA write
    __syncthreads();                                    // the buffer is free for the prefetch after next
  }
  if (stats) {
    if (cv_on && col < TW) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        atomicAdd(&s_stat[cv * 8 + 2 * i], tsum[i].x);
        atomicAdd(&s_stat[cv * 8 + 2 * i + 1], tsum[i].y);
        atomicAdd(&s_stat[Cfg::CC + cv * 8 + 2 * i], tsq[i].x);
        atomicAdd(&s_stat[Cfg::CC + cv * 8 + 2 * i + 1], tsq[i].y);
      }
    }
    __syncthreads();
    for (int i = tid; i < Cfg::CC; i += 128) {
      const int c = chunk * Cfg::CC + i;
      if (c < g.c) {
        atomicAdd(&stats[c], static_cast<double>(s_stat[i]));
        atomicAdd(&stats[g.c + c], static_cast<double>(s_stat[Cfg::CC + i]));
      }
    }
  }
  bn_finalize_if_last(fin, stats, g.c, stats != nullptr && tid < Cfg::CC);
}

// ------------------------------------------------------------------------------------------------
// Fused backward: d(a) and d(w) from ONE staging of rowop(dy) and rowop(a).
//   stride 1: A tile and DY tile both (TH+2) x (TW+2) with origin (-1,-1); d(a) = conv_sweep over the DY tile
//             with flipped taps.
//   stride 2: A tile (2TH+1) x (2TW+1) origin (2ho0-1, 2wo0-1); DY tile (TH+1) x (TW+1) origin (ho0, wo0);
//             thread (cv, ox) produces the input columns 2ox, 2ox+1 of the 2TH input rows the tile owns.
//   d(w)   : generic sweep, acc9[kh*3+kw] += dy[oy][ox] * a[oy*S+kh][ox*S+kw], accumulated in registers over
//            every tile the CTA visits.
// ------------------------------------------------------------------------------------------------
template <int STRIDE, int CVN, int IW, int IH, int TH, int DWID, int DOFF>
__device__ __forceinline__ void wgrad_sweep(uint32_t a_tile, uint32_t d_tile, int cv, int ox, float2 (&acc9)[9][4]) {
  float2 dsl[3][4];
  const uint32_t abase = a_tile + static_cast<uint32_t>((ox * STRIDE * CVN + cv) * 16);
  const uint32_t dbase = d_tile + static_cast<uint32_t>((((DOFF * DWID) + ox + DOFF) * CVN + cv) * 16);
#pragma unroll
  for (int iy = 0; iy < IH; ++iy) {
    if (iy % STRIDE == 0 && iy / STRIDE < TH)
      unpack8(lds128(dbase + static_cast<uint32_t>(((iy / STRIDE) * DWID * CVN) * 16)), dsl[(iy / STRIDE) % 3]);
    float2 xa[3][4];
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) unpack8(lds128(abase + static_cast<uint32_t>(((iy * IW + kw) * CVN) * 16)), xa[kw]);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int t = iy - kh;
      if (t < 0 || (t % STRIDE) != 0 || t / STRIDE >= TH) continue;
      const int slot = (t / STRIDE) % 3;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc9[kh * 3 + kw][i] = __ffma2_rn(dsl[slot][i], xa[kw][i], acc9[kh * 3 + kw][i]);
    }
  }
}

template <int STRIDE, int CVN, int TW, int TH>
__global__ void __launch_bounds__(128, 2)
dw_bwd_sw_kernel(RowOp dy, RowOp a, const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_g,
                 const __grid_constant__ CUtensorMap tm_r, const float* __restrict__ wgt, __nv_bfloat16* __restrict__ da,
                 float* __restrict__ dwgt, DwSw g) {
  using Cfg = DwCfg<STRIDE, CVN, TW, TH>;
  constexpr int DOFF = STRIDE == 1 ? 1 : 0;
  constexpr int DH = TH + 1 + DOFF, DWID = TW + 1 + DOFF;          // dy tile
  constexpr int DT_BYTES = DH * DWID * CVN * 16;
  constexpr int CC = Cfg::CC;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t a_tile = smem_u32(smem), g_tile = a_tile + Cfg::TILE_BYTES, r_tile = g_tile + DT_BYTES;
  float* s_w = reinterpret_cast<float*>(smem + Cfg::TILE_BYTES + 2 * DT_BYTES);   // [9][CC] taps (storage-rounded)
  float* s_acc = s_w + 9 * CC;                                                     // [9][CC]
  const uint32_t bar = smem_u32(s_acc + 9 * CC);                                   // one mbarrier: the tiles of an item
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    tma::prefetch_map(&tm_a);
    tma::prefetch_map(&tm_g);
  }
  const int tid = threadIdx.x, cv = tid % CVN, col = tid / CVN;
  const int chunk = blockIdx.x % g.n_chunks;
  const int c_base = chunk * CC;
  const int c0 = c_base + cv * 8;
  const bool cv_on = c0 < g.c;
  for (int i = tid; i < 9 * CC; i += 128) {
    const int t = i / CC, c = i - t * CC;
    s_w[i] = c_base + c < g.c ? round_to<__nv_bfloat16>(wgt[(c_base + c) * 9 + t]) : 0.f;
    s_acc[i] = 0.f;
  }
  float2 acc9[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc9[t][i] = make_float2(0.f, 0.f);

  const bool two = dy.mode == EHGR_ROW_BNBWD;
  const long long item_stride = gridDim.x / g.n_chunks;
  uint32_t phase = 0;
  for (long long item = blockIdx.x / g.n_chunks; item < g.items; item += item_stride) {
    const int tx = static_cast<int>(item % g.tiles_x);
    const long long r = item / g.tiles_x;
    const int ho0 = static_cast<int>(r % g.tiles_y) * TH, wo0 = tx * TW;
    const long long nt = r / g.tiles_y;
    fence_proxy_async();                   // generic-proxy accesses of the tiles precede the next TMA writes
    __syncthreads();                       // previous item's sweeps are done with the tiles (and the barrier is initialised)
    if (tid == 0) {                        // one TMA box per tensor: tile + halo of this channel chunk, zero-filled outside
      tma::expect_tx(bar, Cfg::TILE_BYTES + (two ? 2 : 1) * DT_BYTES);
      tma::load_4d(a_tile, &tm_a, bar, c_base, wo0 * STRIDE - 1, ho0 * STRIDE - 1, static_cast<int>(nt));
      tma::load_4d(g_tile, &tm_g, bar, c_base, wo0 - DOFF, ho0 - DOFF, static_cast<int>(nt));
      if (two) tma::load_4d(r_tile, &tm_r, bar, c_base, wo0 - DOFF, ho0 - DOFF, static_cast<int>(nt));
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    if (cv_on) {
      if (a.mode != EHGR_ROW_PLAIN) {
        RowOp ac = a;
        ac.mode = EHGR_ROW_AFFINE;      // compile-time constants for the loaders (the host admits only these modes)
        RowLoader<__nv_bfloat16, 8, false, false> ld;
        ld.init(ac, c0, g.c);
        tile_transform<Cfg::IH, Cfg::IW, CVN>(ac, ld, a_tile, a_tile, g.h, g.w, ho0 * STRIDE - 1, wo0 * STRIDE - 1, cv, col);
      }
      if (two) {
        RowOp dc = dy;
        dc.mode = EHGR_ROW_BNBWD;
        RowLoader<__nv_bfloat16, 8, true, false> ld;
        ld.init(dc, c0, g.c);
        tile_transform<DH, DWID, CVN>(dc, ld, g_tile, r_tile, g.ho, g.wo, ho0 - DOFF, wo0 - DOFF, cv, col);
      }
    }
    __syncthreads();
    if (col < TW) {
      // ---- weight gradient
      wgrad_sweep<STRIDE, CVN, Cfg::IW, Cfg::IH, TH, DWID, DOFF>(a_tile, g_tile, cv, col, acc9);
      // ---- input gradient
      if constexpr (STRIDE == 1) {
        float2 wf[9][4];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 lo = *reinterpret_cast<const float4*>(s_w + (8 - t) * CC + cv * 8);
          const float4 hi = *reinterpret_cast<const float4*>(s_w + (8 - t) * CC + cv * 8 + 4);
          wf[t][0] = make_float2(lo.x, lo.y); wf[t][1] = make_float2(lo.z, lo.w);
          wf[t][2] = make_float2(hi.x, hi.y); wf[t][3] = make_float2(hi.z, hi.w);
        }
        const int wi = wo0 + col;
        const bool st_ok = cv_on && wi < g.w;
        __nv_bfloat16* orow = da + ((nt * g.h + ho0) * g.w + wi) * g.c + c0;
        conv_sweep<1, CVN, DWID, DH, TH>(g_tile, cv, col, wf, [&](int oy, const float2 (&acc)[4]) {
          if (st_ok && ho0 + oy < g.h) *reinterpret_cast<uint4*>(orow + static_cast<long long>(oy) * g.w * g.c) = pack8f2(acc);
        });
      } else {
        float2 wt[9][4];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 lo = *reinterpret_cast<const float4*>(s_w + t * CC + cv * 8);
          const float4 hi = *reinterpret_cast<const float4*>(s_w + t * CC + cv * 8 + 4);
          wt[t][0] = make_float2(lo.x, lo.y); wt[t][1] = make_float2(lo.z, lo.w);
          wt[t][2] = make_float2(hi.x, hi.y); wt[t][3] = make_float2(hi.z, hi.w);
        }
        const int wi = (wo0 + col) * 2;
        const bool c0_ok = cv_on && wi < g.w, c1_ok = cv_on && wi + 1 < g.w;
        __nv_bfloat16* orow = da + ((nt * g.h + ho0 * 2) * g.w + wi) * g.c + c0;
        const long long rstride = static_cast<long long>(g.w) * g.c;
        const uint32_t dbase = g_tile + static_cast<uint32_t>((col * CVN + cv) * 16);
        float2 o0[4], o1[4];                 // the odd input row in flight (columns 2ox, 2ox+1)
#pragma unroll
        for (int j = 0; j < DH; ++j) {
          float2 d0[4], d1[4];
          unpack8(lds128(dbase + static_cast<uint32_t>((j * DWID * CVN) * 16)), d0);
          unpack8(lds128(dbase + static_cast<uint32_t>(((j * DWID + 1) * CVN) * 16)), d1);
          if (j >= 1) {                      // kh = 0 taps complete the odd row 2(j-1)+1
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              o0[i] = __ffma2_rn(d0[i], wt[1][i], o0[i]);
              o1[i] = __ffma2_rn(d1[i], wt[0][i], __ffma2_rn(d0[i], wt[2][i], o1[i]));
            }
            const int hi = ho0 * 2 + 2 * (j - 1) + 1;
            if (hi < g.h) {
              if (c0_ok) *reinterpret_cast<uint4*>(orow + (2 * (j - 1) + 1) * rstride) = pack8f2(o0);
              if (c1_ok) *reinterpret_cast<uint4*>(orow + (2 * (j - 1) + 1) * rstride + g.c) = pack8f2(o1);
            }
          }
          if (j < TH) {
            float2 e0[4], e1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              e0[i] = __fmul2_rn(d0[i], wt[4][i]);                                       // even row: kh = 1
              e1[i] = __ffma2_rn(d1[i], wt[3][i], __fmul2_rn(d0[i], wt[5][i]));
              o0[i] = __fmul2_rn(d0[i], wt[7][i]);                                       // odd row: kh = 2 now
              o1[i] = __ffma2_rn(d1[i], wt[6][i], __fmul2_rn(d0[i], wt[8][i]));
            }
            const int hi = ho0 * 2 + 2 * j;
            if (hi < g.h) {
              if (c0_ok) *reinterpret_cast<uint4*>(orow + (2 * j) * rstride) = pack8f2(e0);
              if (c1_ok) *reinterpret_cast<uint4*>(orow + (2 * j) * rstride + g.c) = pack8f2(e1);
            }
          }
        }
      }
    }
  }
  if (cv_on && col < TW) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        atomicAdd(&s_acc[t * CC + cv * 8 + 2 * i], acc9[t][i].x);
        atomicAdd(&s_acc[t * CC + cv * 8 + 2 * i + 1], acc9[t][i].y);
      }
  }
  __syncthreads();
  for (int i = tid; i < 9 * CC; i += 128) {
    const int t = i / CC, c = i - t * CC;
    if (c_base + c < g.c) atomicAdd(&dwgt[(c_base + c) * 9 + t], s_acc[i]);
  }
}

template <int STRIDE, int CVN, int TW, int TH>
static void dw_sw_geom(DwSw& g, int nt, int h, int w, int c) {
  g.nt = nt; g.h = h; g.w = w; g.c = c;
  g.ho = (h - 1) / STRIDE + 1;
  g.wo = (w - 1) / STRIDE + 1;
  g.tiles_y = (g.ho + TH - 1) / TH;
  g.tiles_x = (g.wo + TW - 1) / TW;
  g.n_chunks = (c + CVN * 8 - 1) / (CVN * 8);
  g.items = static_cast<long long>(nt) * g.tiles_y * g.tiles_x;
}

static unsigned dw_sw_grid(const DwSw& g, int per_sm) {
  long long blocks = std::min<long long>(g.items * g.n_chunks, static_cast<long long>(kNumSMs) * per_sm);
  blocks = std::max<long long>(g.n_chunks, blocks / g.n_chunks * g.n_chunks);
  return static_cast<unsigned>(blocks);
}

template <int STRIDE, int CVN, int TW, int TH>
static int dw_fwd_sw_go(const RowOp& a, const float* w, void* out, double* stats, int nt, int h, int wd, int c,
                        cudaStream_t s) {
  using Cfg = DwCfg<STRIDE, CVN, TW, TH>;
  DwSw g;
  dw_sw_geom<STRIDE, CVN, TW, TH>(g, nt, h, wd, c);
  const size_t smem = 2 * Cfg::TILE_BYTES + 2 * Cfg::CC * sizeof(float) + 16;
  CUtensorMap tm_a;
  if (int st = tma::make_nhwc_bf16_map(&tm_a, a.in1, nt, h, wd, c, Cfg::CC, Cfg::IW, Cfg::IH)) return st;
  auto kern = dw_fwd_sw_kernel<STRIDE, CVN, TW, TH>;
  ensure_smem(kern, static_cast<int>(smem));
  const int per_sm = std::max(1, std::min(3, static_cast<int>((220 * 1024) / (smem + 1024))));
  kern<<<dw_sw_grid(g, per_sm), 128, smem, s>>>(a, tm_a, w, static_cast<__nv_bfloat16*>(out), stats, g, take_fin());
  return launch_status();
}

// bf16 forward; returns EHGR_E_UNSUPPORTED when the operand mode is not PLAIN / AFFINE (caller falls back)
int dw_fwd_sw(const RowOp& a, const float* w, void* out, double* stats, int nt, int h, int wd, int c, int stride,
              cudaStream_t s) {
  if (a.mode != EHGR_ROW_PLAIN && a.mode != EHGR_ROW_AFFINE) return EHGR_E_UNSUPPORTED;
  const int wo = (wd - 1) / stride + 1;
  if (stride == 1) {
    if (wo <= 7) return dw_fwd_sw_go<1, 16, 7, 7>(a, w, out, stats, nt, h, wd, c, s);
    return dw_fwd_sw_go<1, 8, 14, 14>(a, w, out, stats, nt, h, wd, c, s);
  }
  if (wo <= 7) return dw_fwd_sw_go<2, 16, 7, 4>(a, w, out, stats, nt, h, wd, c, s);
  return dw_fwd_sw_go<2, 8, 14, 4>(a, w, out, stats, nt, h, wd, c, s);
}

template <int STRIDE, int CVN, int TW, int TH>
static int dw_bwd_sw_go(const RowOp& dy, const RowOp& a, const float* w, void* da, float* dw, int nt, int h, int wd, int c,
                        cudaStream_t s) {
  using Cfg = DwCfg<STRIDE, CVN, TW, TH>;
  constexpr int DOFF = STRIDE == 1 ? 1 : 0;
  constexpr int DT_BYTES = (TH + 1 + DOFF) * (TW + 1 + DOFF) * CVN * 16;
  DwSw g;
  dw_sw_geom<STRIDE, CVN, TW, TH>(g, nt, h, wd, c);
  const size_t smem = Cfg::TILE_BYTES + 2 * DT_BYTES + 18 * Cfg::CC * sizeof(float) + 16;
  const int ho = (h - 1) / STRIDE + 1, wo = (wd - 1) / STRIDE + 1;
  CUtensorMap tm_a, tm_g, tm_r;
  if (int st = tma::make_nhwc_bf16_map(&tm_a, a.in1, nt, h, wd, c, Cfg::CC, Cfg::IW, Cfg::IH)) return st;
  if (int st = tma::make_nhwc_bf16_map(&tm_g, dy.in1, nt, ho, wo, c, Cfg::CC, TW + 1 + DOFF, TH + 1 + DOFF)) return st;
  tm_r = tm_g;
  if (dy.mode == EHGR_ROW_BNBWD)
    if (int st = tma::make_nhwc_bf16_map(&tm_r, dy.in2, nt, ho, wo, c, Cfg::CC, TW + 1 + DOFF, TH + 1 + DOFF)) return st;
  auto kern = dw_bwd_sw_kernel<STRIDE, CVN, TW, TH>;
  ensure_smem(kern, static_cast<int>(smem));
  const int per_sm = std::max(1, std::min(2, static_cast<int>((220 * 1024) / (smem + 1024))));
  kern<<<dw_sw_grid(g, per_sm), 128, smem, s>>>(dy, a, tm_a, tm_g, tm_r, w, static_cast<__nv_bfloat16*>(da), dw, g);
  return launch_status();
}

// bf16 fused backward; EHGR_E_UNSUPPORTED when an operand mode is outside {PLAIN, AFFINE} x {PLAIN, BNBWD}
int dw_bwd_sw(const RowOp& dy, const RowOp& a, const float* w, void* da, float* dw, int nt, int h, int wd, int c,
              int stride, cudaStream_t s) {
  if (a.mode != EHGR_ROW_PLAIN && a.mode != EHGR_ROW_AFFINE) return EHGR_E_UNSUPPORTED;
  if (dy.mode != EHGR_ROW_PLAIN && dy.mode != EHGR_ROW_BNBWD) return EHGR_E_UNSUPPORTED;
  const int wo = (wd - 1) / stride + 1;
  if (stride == 1) {
    if (wo <= 7) return dw_bwd_sw_go<1, 16, 7, 7>(dy, a, w, da, dw, nt, h, wd, c, s);
    return dw_bwd_sw_go<1, 8, 14, 14>(dy, a, w, da, dw, nt, h, wd, c, s);
  }
  if (wo <= 7) return dw_bwd_sw_go<2, 16, 7, 4>(dy, a, w, da, dw, nt, h, wd, c, s);
  return dw_bwd_sw_go<2, 8, 14, 4>(dy, a, w, da, dw, nt, h, wd, c, s);
}

}  // namespace ehgr
