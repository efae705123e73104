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
#include "tc_common.cuh"

namespace ehgr {

using tc::cp_async16;
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
dw_fwd_sw_kernel(RowOp a, const float* __restrict__ wgt, __nv_bfloat16* __restrict__ out, double* __restrict__ stats, DwSw g) {
  using Cfg = DwCfg<STRIDE, CVN, TW, TH>;
  using Ld = RowLoader<__nv_bfloat16, 8, false, false>;
  extern __shared__ __align__(128) uint8_t smem[];
  float* s_stat = reinterpret_cast<float*>(smem + 2 * Cfg::TILE_BYTES);   // [2][CC]
  const uint32_t tile0 = smem_u32(smem);
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
  if (cv_on) ld.init(a, c0, g.c);
  float2 tsum[4], tsq[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) tsum[i] = tsq[i] = make_float2(0.f, 0.f);

  const __nv_bfloat16* in1 = static_cast<const __nv_bfloat16*>(a.in1);
  const long long item_stride = gridDim.x / g.n_chunks;
  const long long first = blockIdx.x / g.n_chunks;
  auto origin = [&](long long item, long long& nt, int& ho0, int& wo0) {
    const int tx = static_cast<int>(item % g.tiles_x);
    const long long r = item / g.tiles_x;
    ho0 = static_cast<int>(r % g.tiles_y) * TH;
    wo0 = tx * TW;
    nt = r / g.tiles_y;
  };
  if (first < g.items && cv_on) {
    long long nt; int ho0, wo0;
    origin(first, nt, ho0, wo0);
    tile_copy<Cfg::IH, Cfg::IW, CVN>(in1, tile0, nt, g.h, g.w, g.c, c0, ho0 * STRIDE - 1, wo0 * STRIDE - 1, cv, col);
  }
  cp_async_commit();
  int buf = 0;
  for (long long item = first; item < g.items; item += item_stride, buf ^= 1) {
    long long nt; int ho0, wo0;
    origin(item, nt, ho0, wo0);
    const uint32_t tile = tile0 + buf * Cfg::TILE_BYTES;
    if (item + item_stride < g.items && cv_on) {        // prefetch the next tile into the other buffer
      long long nt2; int ho2, wo2;
      origin(item + item_stride, nt2, ho2, wo2);
      tile_copy<Cfg::IH, Cfg::IW, CVN>(in1, tile0 + (buf ^ 1) * Cfg::TILE_BYTES, nt2, g.h, g.w, g.c, c0, ho2 * STRIDE - 1,
                                       wo2 * STRIDE - 1, cv, col);
    }
    cp_async_commit();
    cp_async_wait_group<1>();                          // this tile's copies have landed (the next one's may not)
    if (cv_on && a.mode != EHGR_ROW_PLAIN)
      tile_transform<Cfg::IH, Cfg::IW, CVN, Ld>(a, ld, tile, tile, g.h, g.w, ho0 * STRIDE - 1, wo0 * STRIDE - 1, cv, col);
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
    __syncthreads();                                    // the buffer is free for the prefetch after next
  }
  cp_async_wait_group<0>();
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
  const size_t smem = 2 * Cfg::TILE_BYTES + 2 * Cfg::CC * sizeof(float);
  auto kern = dw_fwd_sw_kernel<STRIDE, CVN, TW, TH>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  const int per_sm = std::max(1, std::min(3, static_cast<int>((220 * 1024) / (smem + 1024))));
  kern<<<dw_sw_grid(g, per_sm), 128, smem, s>>>(a, w, static_cast<__nv_bfloat16*>(out), stats, g);
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

}  // namespace ehgr
