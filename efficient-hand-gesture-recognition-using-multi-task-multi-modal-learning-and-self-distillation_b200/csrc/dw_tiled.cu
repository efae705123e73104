// dw_tiled.cu — K7 depthwise 3x3 (pad 1, stride 1|2), shared-memory-tiled kernels.
//
// The row operand (BatchNorm+ReLU6 of the previous layer for the forward input, the BatchNorm-backward
// combination for the gradient) is evaluated ONCE per element while a spatial tile (+halo) of a channel
// chunk is staged into shared memory; the nine taps then read shared memory.  (The untiled kernels in
// dw.cu re-evaluate the operand for every tap and were instruction-bound.)
//
//   forward : out = dw3x3(rowop(a)) + per-channel batch statistics
//   backward: ONE kernel produces both d(a) (input gradient) and d(w) (weight gradient) from a single
//             staging of rowop(dy) and rowop(a): dy/raw/a are read from HBM once instead of twice.
//
// Work item = (frame, tile_y, tile_x) of one channel chunk; persistent blocks stride over the items of
// THEIR chunk, so statistics / weight-gradient partial sums stay in registers or shared memory and are
// flushed once per block.  Thread = (channel vector cv, position lane p); consecutive lanes read
// consecutive 16-byte vectors of one position: conflict-free shared memory, coalesced global memory.
#include "rowop.cuh"

namespace ehgr {

struct DwTile {
  int nt, h, w, c, stride, ho, wo;
  int th, tw;            // output tile
  int tiles_y, tiles_x;
  int cc;                // channels per chunk (multiple of the vector width)
  int n_chunks;
  long long items;       // nt * tiles_y * tiles_x
};

template <typename T>
__device__ __forceinline__ void st_smem_vec(T* dst, const float (&v)[VecOf<T>::N]) { store_vec<T, VecOf<T>::N>(dst, v); }

// Stage rowop(op) for the rectangle [r0, r0+nr) x [c0w, c0w+ncw) of frame `nt` (image rh x rw, zero
// outside) into smem [pos][CV] as T.  Batches of four fetches per thread.
template <typename T, bool kTwo>
__device__ __forceinline__ void stage_tile(const RowOp& op, const RowLoader<T, VecOf<T>::N, kTwo>& ld, T* smem, int CV,
                                           int cv, int p, int P, long long nt, int rh, int rw, int r0, int c0w, int nr,
                                           int ncw) {
  constexpr int V = VecOf<T>::N;
  using Ld = RowLoader<T, V, kTwo>;
  const int np = nr * ncw;
#pragma unroll 1
  for (int i0 = p; i0 < np; i0 += 4 * P) {
    typename Ld::Raw raw[4];
    bool in_img[4];
    int idx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      idx[j] = i0 + j * P;
      in_img[j] = false;
      if (idx[j] < np) {
        const int ir = idx[j] / ncw, ic = idx[j] - ir * ncw;
        const int hh = r0 + ir, ww = c0w + ic;
        in_img[j] = hh >= 0 && hh < rh && ww >= 0 && ww < rw;
        if (in_img[j]) raw[j] = ld.fetch(op, (nt * rh + hh) * rw + ww);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (idx[j] < np) {
        T* dst = smem + (static_cast<size_t>(idx[j]) * CV + cv) * V;
        if constexpr (sizeof(T) == 2) {
          *reinterpret_cast<uint4*>(dst) = in_img[j] ? ld.finish_packed(op, raw[j]) : make_uint4(0, 0, 0, 0);
        } else {
          float v[V];
#pragma unroll
          for (int i = 0; i < V; ++i) v[i] = 0.f;
          if (in_img[j]) ld.finish(op, raw[j], v);
          st_smem_vec<T>(dst, v);
        }
      }
    }
  }
}

template <typename T, bool kTwo>
__global__ void __launch_bounds__(128, 4)
dw_fwd_tiled_kernel(RowOp a, const float* __restrict__ wgt, T* __restrict__ out, double* __restrict__ stats, DwTile g) {
  constexpr int V = VecOf<T>::N;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int chunk = blockIdx.x % g.n_chunks;
  const int c_base = chunk * g.cc;
  const int cc = min(g.cc, g.c - c_base);
  const int CV = cc / V;
  const int P = blockDim.x / CV;
  const int cv = threadIdx.x % CV, p = threadIdx.x / CV;
  const bool active = p < P;
  const int ih = (g.th - 1) * g.stride + 3, iw = (g.tw - 1) * g.stride + 3;
  T* tile = reinterpret_cast<T*>(smem_raw);                                     // [ih*iw][CV][V]
  float* s_stat = reinterpret_cast<float*>(smem_raw + static_cast<size_t>(ih) * iw * g.cc * sizeof(T));  // [2*cc]
  for (int i = threadIdx.x; i < 2 * g.cc; i += blockDim.x) s_stat[i] = 0.f;

  T wreg[9][V];   // weights in the storage type: bf16 mode rounds them like autocast would (fewer registers)
  const int c0 = c_base + cv * V;
  if (active) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int i = 0; i < V; ++i) wreg[t][i] = static_cast<T>(wgt[(c0 + i) * 9 + t]);
  }
  float tsum[V], tsq[V];
#pragma unroll
  for (int i = 0; i < V; ++i) tsum[i] = tsq[i] = 0.f;

  const long long item_stride = gridDim.x / g.n_chunks;
  for (long long item = blockIdx.x / g.n_chunks; item < g.items; item += item_stride) {
    const int tx = static_cast<int>(item % g.tiles_x);
    const long long r = item / g.tiles_x;
    const int ty = static_cast<int>(r % g.tiles_y);
    const long long nt = r / g.tiles_y;
    const int ho0 = ty * g.th, wo0 = tx * g.tw;
    __syncthreads();   // previous item's compute is done with the tile
    if (active) {   // operand coefficients live only while staging (registers are needed by the tap loop)
      RowLoader<T, V, kTwo> ld;
      ld.init(a, c0, g.c);
      stage_tile<T, kTwo>(a, ld, tile, CV, cv, p, P, nt, g.h, g.w, ho0 * g.stride - 1, wo0 * g.stride - 1, ih, iw);
    }
    __syncthreads();
    if (active) {
      const int nout = g.th * g.tw;
#pragma unroll 1
      for (int idx = p; idx < nout; idx += P) {
        const int oh = idx / g.tw, ow = idx - oh * g.tw;
        const int ho = ho0 + oh, wo = wo0 + ow;
        if (ho >= g.ho || wo >= g.wo) continue;
        float acc[V];
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = 0.f;
        const T* base = tile + (static_cast<size_t>(oh * g.stride) * iw + ow * g.stride) * CV * V + cv * V;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          float v[V];
          load_vec<T, V>(base + static_cast<size_t>((t / 3) * iw + (t % 3)) * CV * V, v);
#pragma unroll
          for (int i = 0; i < V; ++i) acc[i] = fmaf(v[i], static_cast<float>(wreg[t][i]), acc[i]);
        }
        store_vec<T, V>(out + ((nt * g.ho + ho) * g.wo + wo) * g.c + c0, acc);
#pragma unroll
        for (int i = 0; i < V; ++i) { tsum[i] += acc[i]; tsq[i] = fmaf(acc[i], acc[i], tsq[i]); }
      }
    }
  }
  if (stats) {
    if (active) {
#pragma unroll
      for (int i = 0; i < V; ++i) { atomicAdd(&s_stat[cv * V + i], tsum[i]); atomicAdd(&s_stat[g.cc + cv * V + i], tsq[i]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cc; i += blockDim.x) {
      atomicAdd(&stats[c_base + i], static_cast<double>(s_stat[i]));
      atomicAdd(&stats[g.c + c_base + i], static_cast<double>(s_stat[g.cc + i]));
    }
  }
}

// Fused backward.  Tile over OUTPUT positions [ho0, ho0+th) x [wo0, wo0+tw):
//   A  tile : rowop(a) over input rows ho0*s-1 .. (ho0+th-1)*s+1 (halo)          -> wgrad
//   DY tile : rowop(dy) over output rows ho0-lo .. ho0+th (lo = 1 for stride 1, 0 for stride 2) -> dgrad + wgrad
//   owns the input positions hi in [ho0*s, (ho0+th)*s): da[hi,wi] = sum_taps dy[(hi+1-kh)/s, (wi+1-kw)/s] w[kh,kw]
template <typename T, bool kTwoDy, int STRIDE>
__global__ void __launch_bounds__(128, 2)
dw_bwd_tiled_kernel(RowOp dy, RowOp a, const float* __restrict__ wgt, T* __restrict__ da, float* __restrict__ dwgt,
                    DwTile g) {
  constexpr int V = VecOf<T>::N;
  constexpr int LO = STRIDE == 1 ? 1 : 0;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int chunk = blockIdx.x % g.n_chunks;
  const int c_base = chunk * g.cc;
  const int cc = min(g.cc, g.c - c_base);
  const int CV = cc / V;
  const int P = blockDim.x / CV;
  const int cv = threadIdx.x % CV, p = threadIdx.x / CV;
  const bool active = p < P;
  const int ih = (g.th - 1) * STRIDE + 3, iw = (g.tw - 1) * STRIDE + 3;
  const int dh = g.th + LO + 1, dwid = g.tw + LO + 1;
  T* a_tile = reinterpret_cast<T*>(smem_raw);
  T* d_tile = a_tile + static_cast<size_t>(ih) * iw * g.cc;
  float* s_w = reinterpret_cast<float*>(d_tile + static_cast<size_t>(dh) * dwid * g.cc);   // [9][cc]
  float* s_acc = s_w + 9 * g.cc;                                                          // [9][cc]
  for (int i = threadIdx.x; i < 9 * cc; i += blockDim.x) {
    const int t = i / cc, c = i - t * cc;
    s_w[t * g.cc + c] = wgt[(c_base + c) * 9 + t];
    s_acc[t * g.cc + c] = 0.f;
  }
  const int c0 = c_base + cv * V;
  float acc9[9][V];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < V; ++i) acc9[t][i] = 0.f;

  const long long item_stride = gridDim.x / g.n_chunks;
  for (long long item = blockIdx.x / g.n_chunks; item < g.items; item += item_stride) {
    const int tx = static_cast<int>(item % g.tiles_x);
    const long long r = item / g.tiles_x;
    const int ty = static_cast<int>(r % g.tiles_y);
    const long long nt = r / g.tiles_y;
    const int ho0 = ty * g.th, wo0 = tx * g.tw;
    __syncthreads();
    if (active) {   // operand coefficients live only while staging
      {
        RowLoader<T, V, false> ld_a;
        ld_a.init(a, c0, g.c);
        stage_tile<T, false>(a, ld_a, a_tile, CV, cv, p, P, nt, g.h, g.w, ho0 * STRIDE - 1, wo0 * STRIDE - 1, ih, iw);
      }
      {
        RowLoader<T, V, kTwoDy> ld_dy;
        ld_dy.init(dy, c0, g.c);
        stage_tile<T, kTwoDy>(dy, ld_dy, d_tile, CV, cv, p, P, nt, g.ho, g.wo, ho0 - LO, wo0 - LO, dh, dwid);
      }
    }
    __syncthreads();
    if (active) {
      // ---- weight gradient: over the tile's output positions
      const int nout = g.th * g.tw;
#pragma unroll 1
      for (int idx = p; idx < nout; idx += P) {
        const int oh = idx / g.tw, ow = idx - oh * g.tw;
        if (ho0 + oh >= g.ho || wo0 + ow >= g.wo) continue;
        float d[V];
        load_vec<T, V>(d_tile + (static_cast<size_t>(oh + LO) * dwid + ow + LO) * CV * V + cv * V, d);
        const T* base = a_tile + (static_cast<size_t>(oh * STRIDE) * iw + ow * STRIDE) * CV * V + cv * V;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          float v[V];
          load_vec<T, V>(base + static_cast<size_t>((t / 3) * iw + (t % 3)) * CV * V, v);
#pragma unroll
          for (int i = 0; i < V; ++i) acc9[t][i] = fmaf(d[i], v[i], acc9[t][i]);
        }
      }
      // ---- input gradient: over the input positions this tile owns
      const int nih = g.th * STRIDE, niw = g.tw * STRIDE;
#pragma unroll 1
      for (int idx = p; idx < nih * niw; idx += P) {
        const int lh = idx / niw, lw = idx - lh * niw;
        const int hi = ho0 * STRIDE + lh, wi = wo0 * STRIDE + lw;
        if (hi >= g.h || wi >= g.w) continue;
        float acc[V];
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = 0.f;
        if (STRIDE == 1) {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            // ho = hi + 1 - kh  ->  tile row (lh + 1 - kh) + LO = lh + 2 - kh
            float v[V], wv[V];
            load_vec<T, V>(d_tile + (static_cast<size_t>(lh + 2 - t / 3) * dwid + (lw + 2 - t % 3)) * CV * V + cv * V, v);
            load_vec<float, V>(s_w + t * g.cc + cv * V, wv);
#pragma unroll
            for (int i = 0; i < V; ++i) acc[i] = fmaf(v[i], wv[i], acc[i]);
          }
        } else {
          const int kh0 = (lh + 1) & 1, kw0 = (lw + 1) & 1;   // hi, wi have the parity of lh, lw
#pragma unroll
          for (int a2 = 0; a2 < 2; ++a2) {
            const int kh = kh0 + 2 * a2;
            if (kh > 2) continue;
#pragma unroll
            for (int b2 = 0; b2 < 2; ++b2) {
              const int kw = kw0 + 2 * b2;
              if (kw > 2) continue;
              const int th = (lh + 1 - kh) >> 1, tw2 = (lw + 1 - kw) >> 1;   // >= 0 by construction
              float v[V], wv[V];
              load_vec<T, V>(d_tile + (static_cast<size_t>(th) * dwid + tw2) * CV * V + cv * V, v);
              load_vec<float, V>(s_w + (kh * 3 + kw) * g.cc + cv * V, wv);
#pragma unroll
              for (int i = 0; i < V; ++i) acc[i] = fmaf(v[i], wv[i], acc[i]);
            }
          }
        }
        store_vec<T, V>(da + ((nt * g.h + hi) * g.w + wi) * g.c + c0, acc);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int i = 0; i < V; ++i) atomicAdd(&s_acc[t * g.cc + cv * V + i], acc9[t][i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * cc; i += blockDim.x) {
    const int t = i / cc, c = i - t * cc;
    atomicAdd(&dwgt[(c_base + c) * 9 + t], s_acc[t * g.cc + c]);
  }
}

static int dw_tile_geom(DwTile& g, int nt, int h, int w, int c, int stride, int dtype) {
  const int es = esize_of(dtype);
  if (es == 0) return EHGR_E_DTYPE;
  const int V = 16 / es;
  if (nt < 0 || h <= 0 || w <= 0 || c <= 0 || (c % V) || (stride != 1 && stride != 2)) return EHGR_E_SHAPE;
  g.nt = nt; g.h = h; g.w = w; g.c = c; g.stride = stride;
  g.ho = (h - 1) / stride + 1;
  g.wo = (w - 1) / stride + 1;
  const int tmax = stride == 1 ? 14 : 7;
  auto pick = [tmax](int n) {   // largest tile <= tmax that wastes little: prefer an exact divisor >= 7
    if (n <= tmax) return n;
    for (int t = tmax; t >= 7; --t) if (n % t == 0) return t;
    return tmax;
  };
  g.th = pick(g.ho);
  g.tw = pick(g.wo);
  g.tiles_y = (g.ho + g.th - 1) / g.th;
  g.tiles_x = (g.wo + g.tw - 1) / g.tw;
  const int target = 64;
  g.n_chunks = (c + target - 1) / target;
  g.cc = ((c + g.n_chunks - 1) / g.n_chunks + V - 1) / V * V;
  g.n_chunks = (c + g.cc - 1) / g.cc;
  g.items = static_cast<long long>(nt) * g.tiles_y * g.tiles_x;
  return EHGR_OK;
}

static unsigned dw_tile_grid(const DwTile& g, int per_sm) {
  long long blocks = std::min<long long>(g.items * g.n_chunks, static_cast<long long>(kNumSMs) * per_sm);
  blocks = std::max<long long>(g.n_chunks, blocks / g.n_chunks * g.n_chunks);
  return static_cast<unsigned>(blocks);
}

template <typename T>
static int dw_fwd_tiled_launch(const RowOp& a, const float* w, void* out, double* stats, const DwTile& g, cudaStream_t s) {
  const int ih = (g.th - 1) * g.stride + 3, iw = (g.tw - 1) * g.stride + 3;
  const size_t smem = static_cast<size_t>(ih) * iw * g.cc * sizeof(T) + 2 * g.cc * sizeof(float);
  const unsigned grid = dw_tile_grid(g, 4);
  if (a.mode == EHGR_ROW_BNBWD) {
    ensure_smem(dw_fwd_tiled_kernel<T, true>, 160 * 1024);
    dw_fwd_tiled_kernel<T, true><<<grid, 128, smem, s>>>(a, w, static_cast<T*>(out), stats, g);
  } else {
    ensure_smem(dw_fwd_tiled_kernel<T, false>, 160 * 1024);
    dw_fwd_tiled_kernel<T, false><<<grid, 128, smem, s>>>(a, w, static_cast<T*>(out), stats, g);
  }
  return launch_status();
}

template <typename T, bool kTwo, int STRIDE>
static void dw_bwd_tiled_go(const RowOp& dy, const RowOp& a, const float* w, void* da, float* dw, const DwTile& g,
                            unsigned grid, size_t smem, cudaStream_t s) {
  ensure_smem(dw_bwd_tiled_kernel<T, kTwo, STRIDE>, 200 * 1024);
  dw_bwd_tiled_kernel<T, kTwo, STRIDE><<<grid, 128, smem, s>>>(dy, a, w, static_cast<T*>(da), dw, g);
}

template <typename T>
static int dw_bwd_tiled_launch(const RowOp& dy, const RowOp& a, const float* w, void* da, float* dw, const DwTile& g,
                               cudaStream_t s) {
  const int lo = g.stride == 1 ? 1 : 0;
  const int ih = (g.th - 1) * g.stride + 3, iw = (g.tw - 1) * g.stride + 3;
  const int dh = g.th + lo + 1, dwid = g.tw + lo + 1;
  const size_t smem = (static_cast<size_t>(ih) * iw + static_cast<size_t>(dh) * dwid) * g.cc * sizeof(T) +
                      18 * g.cc * sizeof(float);
  const unsigned grid = dw_tile_grid(g, 2);
  const bool two = dy.mode == EHGR_ROW_BNBWD;
  if (g.stride == 1) {
    if (two) dw_bwd_tiled_go<T, true, 1>(dy, a, w, da, dw, g, grid, smem, s);
    else dw_bwd_tiled_go<T, false, 1>(dy, a, w, da, dw, g, grid, smem, s);
  } else {
    if (two) dw_bwd_tiled_go<T, true, 2>(dy, a, w, da, dw, g, grid, smem, s);
    else dw_bwd_tiled_go<T, false, 2>(dy, a, w, da, dw, g, grid, smem, s);
  }
  return launch_status();
}

int dw_fwd_sw(const RowOp& a, const float* w, void* out, double* stats, int nt, int h, int wd, int c, int stride,
              cudaStream_t s);

int dw_bwd_sw(const RowOp& dy, const RowOp& a, const float* w, void* da, float* dw, int nt, int h, int wd, int c,
              int stride, cudaStream_t s);

int dw_fwd_tiled(const RowOp& a, const float* w, void* out, double* stats, int nt, int h, int wd, int c, int stride,
                 int dtype, cudaStream_t s) {
  if (dtype == EHGR_BF16 && nt > 0) {   // bf16: asynchronous sliding-window kernels (dw_sw.cu)
    const int st = dw_fwd_sw(a, w, out, stats, nt, h, wd, c, stride, s);
    if (st != EHGR_E_UNSUPPORTED) return st;
  }
  DwTile g;
  if (int st = dw_tile_geom(g, nt, h, wd, c, stride, dtype)) return st;
  if (g.items == 0) return EHGR_OK;
  return dtype == EHGR_F32 ? dw_fwd_tiled_launch<float>(a, w, out, stats, g, s)
                           : dw_fwd_tiled_launch<__nv_bfloat16>(a, w, out, stats, g, s);
}

}  // namespace ehgr

using namespace ehgr;

extern "C" int ehgr_dw_bwd(const ehgr_rowop* dy, const ehgr_rowop* a, const float* w, void* da, float* dw, int nt,
                           int h, int wd, int c, int stride, int dtype, ehgr_stream_t stream) {
  DwTile g;
  if (int st = dw_tile_geom(g, nt, h, wd, c, stride, dtype)) return st;
  if (!w || !da || !dw) return EHGR_E_NULL;
  if (int st = validate_rowop_nogate(dy, esize_of(dtype))) return st;
  if (int st = validate_rowop_nogate(a, esize_of(dtype))) return st;
  if (a->mode == EHGR_ROW_BNBWD) return EHGR_E_UNSUPPORTED;
  if (!aligned_to(da, 16)) return EHGR_E_ALIGN;
  if (g.items == 0) return EHGR_OK;
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_BF16) {   // bf16: asynchronous sliding-window kernel (dw_sw.cu)
    const int st = dw_bwd_sw(*dy, *a, w, da, dw, nt, h, wd, c, stride, s);
    if (st != EHGR_E_UNSUPPORTED) return st;
  }
  return dtype == EHGR_F32 ? dw_bwd_tiled_launch<float>(*dy, *a, w, da, dw, g, s)
                           : dw_bwd_tiled_launch<__nv_bfloat16>(*dy, *a, w, da, dw, g, s);
}
