// action.cu — K2-K6: the ACTION module (models/action.py:61-116) as fused kernels, forward and backward.
//
//   xs  = per-channel 3-tap temporal FIR of x (zero padded)                            (:65-73)
//   g1  = sigmoid(conv3d_3x3x3(mean_c xs))                       [N,T,H,W]   STE      (:77-83)
//   g2  = sigmoid(W_ex relu(conv1d_T(W_sq mean_hw xs)))          [N,T,C]     CE       (:86-96)
//   g3  = sigmoid(W_ex3 mean_hw( dw3x3(x3)[t+1] - x3[t] )),  x3 = BN(W_sq3 xs), g3[T-1] from 0   ME (:99-113)
//   out = net( xs * (3 + g1 + g2 + g3) )                                               (:83,96,113,115)
//
// The reference materialises ~25 full-size temporaries (two permute+contiguous, three x*g+x, ...).
// Here ONE pass over x writes xs and, in the same pass, every reduction the three excitations need
// (channel mean per pixel, spatial sum per channel, the C->C/16 squeeze q and its BatchNorm statistics),
// one small kernel per clip turns them into the gates, and the gating itself is a row operand (GATE,
// rowop.cuh) applied inside the A-load of the wrapped 1x1 convolution's GEMM.  Backward mirrors it:
// one reduction pass, one per-clip kernel, one pass producing d(xs), one FIR-adjoint pass.
// Row reductions use shared-memory staging per 16-byte channel vector and warp shuffles.
#include "rowop.cuh"

namespace ehgr {

using Act = ehgr_action;

__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + __expf(-x)); }

// ------------------------------------------------------------------------------------------------
// forward pass 1: block = (frame, row split); thread = (channel vector cv, row lane py).  The spatial sums
// (pool) are accumulated with global atomics: the caller zeroes `pool`.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
action_xs_generic_kernel(Act a, const T* __restrict__ x, T* __restrict__ xs) {
  constexpr int V = VecOf<T>::N;
  extern __shared__ float smem[];
  const int C = a.c, Cr = a.cr, HW = a.h * a.w, CV = C / V, P = blockDim.x / CV;
  float* s_sq3 = smem;                         // [Cr][C]
  float* s_row = s_sq3 + Cr * C;               // [P][Cr+1]
  float* s_pool = s_row + P * (Cr + 1);        // [C]
  float* s_qs = s_pool + C;                    // [2*Cr]
  for (int i = threadIdx.x; i < Cr * C; i += blockDim.x) s_sq3[i] = a.p3_squeeze[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_pool[i] = 0.f;
  for (int i = threadIdx.x; i < 2 * Cr; i += blockDim.x) s_qs[i] = 0.f;
  const int cv = threadIdx.x % CV, py = threadIdx.x / CV;
  const bool active = py < P;
  const int c0 = cv * V;
  const long long nt = blockIdx.x;
  const int t = static_cast<int>(nt % a.t);
  const bool has_prev = t > 0, has_next = t < a.t - 1;
  float w0[V], w1[V], w2[V], pool_acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    w0[i] = a.shift_w[(c0 + i) * 3 + 0];
    w1[i] = a.shift_w[(c0 + i) * 3 + 1];
    w2[i] = a.shift_w[(c0 + i) * 3 + 2];
    pool_acc[i] = 0.f;
  }
  float qs_sum = 0.f, qs_sq = 0.f;             // used by threads that finalise q (cv < Cr)
  const long long frame_elems = static_cast<long long>(HW) * C;
  const int rows_per_split = ((HW + gridDim.y - 1) / gridDim.y + P - 1) / P * P;
  const int r_begin = blockIdx.y * rows_per_split, r_end = min(HW, r_begin + rows_per_split);
  for (int r0 = r_begin; r0 < r_end; r0 += P) {
    const int row = r0 + py;
    const bool live = active && row < r_end;
    __syncthreads();
    for (int i = threadIdx.x; i < P * (Cr + 1); i += blockDim.x) s_row[i] = 0.f;
    __syncthreads();
    const long long m = nt * HW + row;
    if (live) {
      const T* px = x + m * C + c0;
      float cur[V], prv[V], nxt[V], v[V];
      load_vec<T, V>(px, cur);
#pragma unroll
      for (int i = 0; i < V; ++i) prv[i] = nxt[i] = 0.f;
      if (has_prev) load_vec<T, V>(px - frame_elems, prv);
      if (has_next) load_vec<T, V>(px + frame_elems, nxt);
      float s8 = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        v[i] = fmaf(w0[i], prv[i], fmaf(w1[i], cur[i], w2[i] * nxt[i]));
        s8 += v[i];
        pool_acc[i] += v[i];
      }
      store_vec<T, V>(xs + m * C + c0, v);
      atomicAdd(&s_row[py * (Cr + 1)], s8);
      for (int j = 0; j < Cr; ++j) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) d = fmaf(s_sq3[j * C + c0 + i], v[i], d);
        atomicAdd(&s_row[py * (Cr + 1) + 1 + j], d);
      }
    }
    __syncthreads();
    if (live) {
      if (cv == 0) a.mrow[m] = s_row[py * (Cr + 1)] / static_cast<float>(C);
      for (int j = cv; j < Cr; j += CV) {        // the row's cv threads share the Cr outputs
        const float qv = s_row[py * (Cr + 1) + 1 + j];
        a.q[m * Cr + j] = qv;
        atomicAdd(&s_qs[j], qv);
        atomicAdd(&s_qs[Cr + j], qv * qv);
      }
    }
  }
  (void)qs_sum; (void)qs_sq;
  if (active) {
#pragma unroll
    for (int i = 0; i < V; ++i) atomicAdd(&s_pool[c0 + i], pool_acc[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&a.pool[nt * C + i], s_pool[i]);   // spatial SUM (row splits add up)
  for (int i = threadIdx.x; i < 2 * Cr; i += blockDim.x) atomicAdd(&a.qstats[i], static_cast<double>(s_qs[i]));
}

// ------------------------------------------------------------------------------------------------
// Lane-group streaming kernels (round 2).  The first versions above staged every row reduction through
// shared-memory float atomics with two block barriers per batch of rows (ncu / events: 5 % of the HBM peak on the
// 56x56 site).  Here the channel vectors of a row sit in G consecutive lanes (G = power of two >= C/V, <= 32), a
// warp covers 32/G rows, every per-row reduction is an xor-shuffle inside the group, and the row loop has no
// barrier.  A CTA owns a row range of ONE CLIP and walks the T frames in order: the frame that is read as
// "t+1" is read again as "t" and "t-1" by the same CTA a few microseconds later (L1/L2 hits: x is fetched from
// HBM once).  Used when C/V <= 32 (all MobileNetV2 sites); wider inputs take the generic kernels.
// ------------------------------------------------------------------------------------------------
struct LaneGroup {
  int G, lg, cvl, rg, rows_per_warp;
  __device__ __forceinline__ LaneGroup(int cv_count, int lane) {
    lg = 0;
    while ((1 << lg) < cv_count) ++lg;
    G = 1 << lg;
    cvl = lane & (G - 1);
    rg = lane >> lg;
    rows_per_warp = 32 >> lg;
  }
  // sum over the lanes of a group (every lane of the group gets it)
  __device__ __forceinline__ float group_sum(float v) const {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  // sum over the row groups of the warp, i.e. over lanes with equal cvl
  __device__ __forceinline__ float rows_sum(float v) const {
    for (int o = 16; o >= G; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
};

template <typename T>
__global__ void __launch_bounds__(256)
action_xs_kernel(Act a, const T* __restrict__ x, T* __restrict__ xs, int row_splits) {
  constexpr int V = VecOf<T>::N;
  extern __shared__ float smem[];
  const int C = a.c, Cr = a.cr, HW = a.h * a.w, CV = C / V, Tn = a.t;
  float* s_sq3 = smem;                         // [Cr][C]
  float* s_qs = s_sq3 + Cr * C;                // [2*Cr]
  for (int i = threadIdx.x; i < Cr * C; i += blockDim.x) s_sq3[i] = a.p3_squeeze[i];
  for (int i = threadIdx.x; i < 2 * Cr; i += blockDim.x) s_qs[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const LaneGroup lgp(CV, lane);
  const bool on = lgp.cvl < CV;
  const int c0 = lgp.cvl * V;
  const long long clip = blockIdx.x / row_splits;
  const int split = blockIdx.x % row_splits;
  const int rows_per_split = (HW + row_splits - 1) / row_splits;
  const int r_begin = split * rows_per_split, r_end = min(HW, r_begin + rows_per_split);
  const int rows_per_pass = (blockDim.x >> 5) * lgp.rows_per_warp;
  float w0[V], w1[V], w2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    w0[i] = on ? a.shift_w[(c0 + i) * 3 + 0] : 0.f;
    w1[i] = on ? a.shift_w[(c0 + i) * 3 + 1] : 0.f;
    w2[i] = on ? a.shift_w[(c0 + i) * 3 + 2] : 0.f;
  }
  float qsum = 0.f, qsq = 0.f;                 // statistics of q[:, j] for j = cvl (lanes cvl < Cr)
  const long long frame_elems = static_cast<long long>(HW) * C;
  for (int t = 0; t < Tn; ++t) {
    const long long nt = clip * Tn + t;
    const bool has_prev = t > 0, has_next = t < Tn - 1;
    float pool_acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) pool_acc[i] = 0.f;
    // the loop bound is warp-uniform (the group reductions below are full-warp shuffles)
    constexpr int U = 2;                         // rows per lane and iteration: 3U independent 16-byte loads in flight
                                                 // (U = 4 at half the resident CTAs measured 20 % slower: the kernel
                                                 // lives on warps in flight, not on loads per warp)
    for (int rb = r_begin + warp * lgp.rows_per_warp; rb < r_end; rb += U * rows_per_pass) {
      const int r0 = rb + lgp.rg;
      uint4 rc[U], rp[U], rn[U];
      bool live[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int row = r0 + u * rows_per_pass;
        live[u] = on && row < r_end;
        rc[u] = rp[u] = rn[u] = make_uint4(0, 0, 0, 0);
        if (live[u]) {
          const T* px = x + (nt * HW + row) * C + c0;
          rc[u] = *reinterpret_cast<const uint4*>(px);
          if (has_prev) rp[u] = *reinterpret_cast<const uint4*>(px - frame_elems);
          if (has_next) rn[u] = *reinterpret_cast<const uint4*>(px + frame_elems);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int row = r0 + u * rows_per_pass;
        const bool row_ok = row < r_end;           // uniform over the lanes of a group
        const long long m = nt * HW + row;
        float cur[V], prv[V], nxt[V], v[V];
        load_vec<T, V>(reinterpret_cast<const T*>(&rc[u]), cur);
        load_vec<T, V>(reinterpret_cast<const T*>(&rp[u]), prv);
        load_vec<T, V>(reinterpret_cast<const T*>(&rn[u]), nxt);
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = fmaf(w0[i], prv[i], fmaf(w1[i], cur[i], w2[i] * nxt[i]));
        float s8 = 0.f;
        if (live[u]) {
          store_vec<T, V>(xs + m * C + c0, v);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            // the reductions see the value as the consumers will read it back (storage rounding)
            const float r = round_to<T>(v[i]);
            v[i] = r;
            s8 += r;
            pool_acc[i] += r;
          }
        }
        s8 = lgp.group_sum(s8);
        if (row_ok && lgp.cvl == 0) a.mrow[m] = s8 / static_cast<float>(C);
        for (int j = 0; j < Cr; ++j) {
          float d = 0.f;
          if (live[u]) {
#pragma unroll
            for (int i = 0; i < V; ++i) d = fmaf(s_sq3[j * C + c0 + i], v[i], d);
          }
          d = lgp.group_sum(d);
          if (row_ok && lgp.cvl == j) {            // lane j of the group owns q[:, j]: coalesced 4*Cr-byte rows
            a.q[m * Cr + j] = d;
            qsum += d;
            qsq = fmaf(d, d, qsq);
          }
        }
      }
    }
    // spatial sums of this frame: over the row groups of the warp, then one atomic per channel and warp
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float ps = lgp.rows_sum(pool_acc[i]);
      if (on && lgp.rg == 0) atomicAdd(&a.pool[nt * C + c0 + i], ps);
    }
  }
  qsum = lgp.rows_sum(qsum);
  qsq = lgp.rows_sum(qsq);
  if (lgp.rg == 0 && lgp.cvl < Cr) {
    atomicAdd(&s_qs[lgp.cvl], qsum);
    atomicAdd(&s_qs[Cr + lgp.cvl], qsq);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * Cr; i += blockDim.x) atomicAdd(&a.qstats[i], static_cast<double>(s_qs[i]));
}

// backward pass 1 in the same lane-group form: dG = gy * xs reduced over channels (-> dg1[m]) and over pixels
// (-> dgc[frame][c], accumulated with one atomic per channel and warp: the caller zeroes dgc)
template <typename T>
__global__ void __launch_bounds__(256)
action_bwd_reduce_kernel(Act a, const T* __restrict__ gy, const T* __restrict__ xs, int row_splits) {
  constexpr int V = VecOf<T>::N;
  const int C = a.c, HW = a.h * a.w, CV = C / V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const LaneGroup lgp(CV, lane);
  const bool on = lgp.cvl < CV;
  const int c0 = lgp.cvl * V;
  const long long nt = blockIdx.x / row_splits;
  const int split = blockIdx.x % row_splits;
  const int rows_per_split = (HW + row_splits - 1) / row_splits;
  const int r_begin = split * rows_per_split, r_end = min(HW, r_begin + rows_per_split);
  const int rows_per_pass = (blockDim.x >> 5) * lgp.rows_per_warp;
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  for (int rb = r_begin + warp * lgp.rows_per_warp; rb < r_end; rb += 2 * rows_per_pass) {   // warp-uniform bound
    const int r0 = rb + lgp.rg;
    float g[2][V], v[2][V];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int row = r0 + u * rows_per_pass;
#pragma unroll
      for (int i = 0; i < V; ++i) g[u][i] = v[u][i] = 0.f;
      if (on && row < r_end) {
        const long long off = (nt * HW + row) * C + c0;
        load_vec<T, V>(gy + off, g[u]);
        load_vec<T, V>(xs + off, v[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int row = r0 + u * rows_per_pass;
      float s8 = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) { const float d = g[u][i] * v[u][i]; s8 += d; acc[i] += d; }
      s8 = lgp.group_sum(s8);
      if (row < r_end && lgp.cvl == 0) a.dg1[nt * HW + row] = s8;
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float cs = lgp.rows_sum(acc[i]);
    if (on && lgp.rg == 0) atomicAdd(&a.dgc[nt * C + c0 + i], cs);
  }
}

// ------------------------------------------------------------------------------------------------
// forward pass 2a: block = one FRAME: the spatio-temporal gate g1 = sigmoid(conv3d 3x3x3 of the channel
// mean) with the three frames it needs staged in shared memory, and the motion squeeze
//   pi[t][j] = mean_p( dw3x3(x3[t+1])[p] - x3[t][p] ),  x3 = BN(q),  t < T-1,
// which only needs per-frame sums of q (total, first / last row and column, corners): the sum over output
// positions of a zero-padded 3x3 conv is a weighted sum of nine rectangle sums of its input.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_rect_term(const float* w9, const float (&sm)[5], const float* qf, int Cr, int H, int W,
                                                 float sc, float sh) {
  // sum_p dw3x3(x3)[p] for x3 = q*sc + sh; sm = {total, row0, rowL, col0, colL} of q over the frame
  float acc = 0.f;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      const float ex_row = kh == 0 ? sm[2] : (kh == 2 ? sm[1] : 0.f);
      const float ex_col = kw == 0 ? sm[4] : (kw == 2 ? sm[3] : 0.f);
      float corner = 0.f;
      if (kh != 1 && kw != 1) corner = qf[(static_cast<long long>(kh == 0 ? H - 1 : 0) * W + (kw == 0 ? W - 1 : 0)) * Cr];
      const int nrows = H - (kh != 1 ? 1 : 0), ncols = W - (kw != 1 ? 1 : 0);
      const float cnt = static_cast<float>((nrows > 0 ? nrows : 0) * (ncols > 0 ? ncols : 0));
      acc = fmaf(w9[kh * 3 + kw], fmaf(sc, sm[0] - ex_row - ex_col + corner, sh * cnt), acc);
    }
  return acc;
}

__global__ void __launch_bounds__(256) action_gates_frame_kernel(Act a) {
  extern __shared__ float smem[];
  const int Cr = a.cr, T = a.t, H = a.h, W = a.w, HW = H * W;
  float* mr = smem;            // [3][HW] channel mean of frames t-1, t, t+1 (zero outside the clip)
  float* red = mr + 3 * HW;    // [6]
  const long long f = blockIdx.x;
  const int t = static_cast<int>(f % T);
  const long long m_t = f * HW;
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 3 * HW; i += blockDim.x) {
    const int fr = i / HW, p = i - fr * HW, tt = t + fr - 1;
    mr[i] = (tt >= 0 && tt < T) ? a.mrow[m_t + static_cast<long long>(fr - 1) * HW + p] : 0.f;
  }
  if (tid < 6) red[tid] = 0.f;
  __syncthreads();
  {
    float w27[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) w27[k] = a.p1_w[k];
    for (int p = tid; p < HW; p += blockDim.x) {
      const int h = p / W, w = p - h * W;
      float acc = 0.f;
#pragma unroll
      for (int dt = -1; dt <= 1; ++dt)
#pragma unroll
        for (int dh = -1; dh <= 1; ++dh)
#pragma unroll
          for (int dw = -1; dw <= 1; ++dw) {
            const int hh = h + dh, ww = w + dw;
            if (hh >= 0 && hh < H && ww >= 0 && ww < W)
              acc = fmaf(w27[(dt + 1) * 9 + (dh + 1) * 3 + dw + 1], mr[(1 + dt) * HW + hh * W + ww], acc);
          }
      a.g1[m_t + p] = sigmoidf(acc);
    }
  }
  // ---- motion squeeze
  const float inv_hw = 1.f / static_cast<float>(HW);
  for (int j = 0; j < Cr; ++j) {
    if (t == T - 1) {                         // the last frame has no successor: pi = 0 (models/action.py:103-105 pads)
      if (tid == 0) a.pi[f * Cr + j] = 0.f;
      continue;
    }
    const float* q0 = a.q + m_t * Cr + j;                                     // frame t
    const float* q1 = a.q + (m_t + HW) * Cr + j;                              // frame t+1
    float s0 = 0.f, tot = 0.f, r0 = 0.f, rl = 0.f, c0 = 0.f, cl = 0.f;
    for (int p = tid; p < HW; p += blockDim.x) {
      const int h = p / W, w = p - h * W;
      s0 += q0[static_cast<long long>(p) * Cr];
      const float q = q1[static_cast<long long>(p) * Cr];
      tot += q;
      if (h == 0) r0 += q;
      if (h == H - 1) rl += q;
      if (w == 0) c0 += q;
      if (w == W - 1) cl += q;
    }
    s0 = warp_sum(s0); tot = warp_sum(tot); r0 = warp_sum(r0); rl = warp_sum(rl); c0 = warp_sum(c0); cl = warp_sum(cl);
    if (lane == 0) {
      atomicAdd(&red[0], s0); atomicAdd(&red[1], tot); atomicAdd(&red[2], r0);
      atomicAdd(&red[3], rl); atomicAdd(&red[4], c0); atomicAdd(&red[5], cl);
    }
    __syncthreads();
    if (tid == 0) {
      const float sc = a.bn3_scale[j], sh = a.bn3_shift[j];
      const float sm[5] = {red[1], red[2], red[3], red[4], red[5]};
      const float conv_sum = block_rect_term(a.p3_conv1 + j * 9, sm, q1, Cr, H, W, sc, sh);
      const float self_sum = fmaf(sc, red[0], sh * static_cast<float>(HW));
      a.pi[f * Cr + j] = (conv_sum - self_sum) * inv_hw;
#pragma unroll
      for (int k = 0; k < 6; ++k) red[k] = 0.f;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// forward pass 2b: block = one clip: channel gate g2 and motion gate g3 from the per-frame reductions
// (pass 2a, action_gates_frame_kernel, has already written g1 and pi)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) action_gates_kernel(Act a) {
  extern __shared__ float smem[];
  const int C = a.c, Cr = a.cr, T = a.t, H = a.h, W = a.w, HW = H * W;
  float* pm = smem;                 // [T][C] spatial mean
  float* s = pm + T * C;            // [T][Cr]
  float* r = s + T * Cr;            // [T][Cr] relu(u)
  float* pi = r + T * Cr;           // [T][Cr]
  const long long n = blockIdx.x;
  const long long f0 = n * T, m0 = f0 * HW;
  const float inv_hw = 1.f / static_cast<float>(HW);
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) pm[i] = a.pool[f0 * C + i] * inv_hw;
  for (int i = threadIdx.x; i < T * Cr; i += blockDim.x) pi[i] = a.pi[f0 * Cr + i];   // written by the frame kernel
  __syncthreads();
  // ---- CE: squeeze
  for (int i = threadIdx.x; i < T * Cr; i += blockDim.x) {
    const int t = i / Cr, j = i - t * Cr;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(a.p2_squeeze[j * C + c], pm[t * C + c], acc);
    s[i] = acc;
    a.s[f0 * Cr + i] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * Cr; i += blockDim.x) {
    const int t = i / Cr, j = i - t * Cr;
    float acc = 0.f;
    for (int k = 0; k < 3; ++k) {
      const int tt = t + k - 1;
      if (tt < 0 || tt >= T) continue;
      for (int j2 = 0; j2 < Cr; ++j2) acc = fmaf(a.p2_conv1[(j * Cr + j2) * 3 + k], s[tt * Cr + j2], acc);
    }
    a.u[f0 * Cr + i] = acc;
    r[i] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  // ---- expand + sigmoid
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) {
    const int t = i / C, c = i - t * C;
    float e2 = 0.f, e3 = 0.f;
    for (int j = 0; j < Cr; ++j) {
      e2 = fmaf(a.p2_expand[c * Cr + j], r[t * Cr + j], e2);
      e3 = fmaf(a.p3_expand[c * Cr + j], pi[t * Cr + j], e3);
    }
    a.g2[f0 * C + i] = sigmoidf(e2);
    a.g3[f0 * C + i] = sigmoidf(e3);
  }
}

// ------------------------------------------------------------------------------------------------
// backward pass 1: dG = gy * xs reduced over channels (-> dg1[m]) and over pixels (-> dgc[frame][c])
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
action_bwd_reduce_generic_kernel(Act a, const T* __restrict__ gy, const T* __restrict__ xs) {
  constexpr int V = VecOf<T>::N;
  extern __shared__ float smem[];
  const int C = a.c, HW = a.h * a.w, CV = C / V, P = blockDim.x / CV;
  float* s_row = smem;          // [P]
  float* s_c = smem + P;        // [C]
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_c[i] = 0.f;
  const int cv = threadIdx.x % CV, py = threadIdx.x / CV;
  const bool active = py < P;
  const int c0 = cv * V;
  const long long nt = blockIdx.x;
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  for (int r0 = 0; r0 < HW; r0 += P) {
    const int row = r0 + py;
    const bool live = active && row < HW;
    __syncthreads();
    if (threadIdx.x < P) s_row[threadIdx.x] = 0.f;
    __syncthreads();
    const long long m = nt * HW + row;
    if (live) {
      float g[V], v[V];
      load_vec<T, V>(gy + m * C + c0, g);
      load_vec<T, V>(xs + m * C + c0, v);
      float s8 = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) { const float d = g[i] * v[i]; s8 += d; acc[i] += d; }
      atomicAdd(&s_row[py], s8);
    }
    __syncthreads();
    if (live && cv == 0) a.dg1[m] = s_row[py];
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < V; ++i) atomicAdd(&s_c[c0 + i], acc[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) a.dgc[nt * C + i] = s_c[i];
}

// border-aware sum of the depthwise weights whose transposed tap lands inside the image at (h,w):
//   d/d x3[t][h,w] of sum_p dw3x3(x3[t])[p]  =  sum over (kh,kw) with output (h-kh+1, w-kw+1) inside
__device__ __forceinline__ float border_wsum(const float* w9, int h, int w, int H, int W) {
  float s = 0.f;
  for (int kh = 0; kh < 3; ++kh) {
    const int ho = h - kh + 1;
    if (ho < 0 || ho >= H) continue;
    for (int kw = 0; kw < 3; ++kw) {
      const int wo = w - kw + 1;
      if (wo < 0 || wo >= W) continue;
      s += w9[kh * 3 + kw];
    }
  }
  return s;
}

// ------------------------------------------------------------------------------------------------
// backward pass 2a: block = one clip: the clip-level part of the gate backward (channel and motion
// excitation through their tiny temporal layers): small parameter gradients, dpool, dd.
// The per-frame spatial work is pass 2b (action_bwd_frame_kernel), which reads dd.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) action_bwd_clip_kernel(Act a) {
  extern __shared__ float smem[];
  const int C = a.c, Cr = a.cr, T = a.t, H = a.h, W = a.w, HW = H * W;
  float* pm = smem;                  // [T][C]
  float* de2 = pm + T * C;           // [T][C]
  float* de3 = de2 + T * C;          // [T][C]
  float* sv = de3 + T * C;           // [T][Cr] squeeze output
  float* rv = sv + T * Cr;           // [T][Cr] relu(u)
  float* du = rv + T * Cr;           // [T][Cr]
  float* ds = du + T * Cr;           // [T][Cr]
  float* dd = ds + T * Cr;           // [T][Cr]
  float* acc27 = dd + T * Cr;        // [27]
  float* acc9 = acc27 + 27;          // [Cr][9]
  float* bsum = acc9 + Cr * 9;       // [2*Cr]
  const long long n = blockIdx.x;
  const long long f0 = n * T, m0 = f0 * HW;
  const float inv_hw = 1.f / static_cast<float>(HW);
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) {
    pm[i] = a.pool[f0 * C + i] * inv_hw;
    const float g2 = a.g2[f0 * C + i], g3 = a.g3[f0 * C + i], d = a.dgc[f0 * C + i];
    de2[i] = d * g2 * (1.f - g2);
    de3[i] = d * g3 * (1.f - g3);
  }
  for (int i = threadIdx.x; i < T * Cr; i += blockDim.x) {
    sv[i] = a.s[f0 * Cr + i];
    rv[i] = fmaxf(a.u[f0 * Cr + i], 0.f);
  }
  for (int i = threadIdx.x; i < 27 + Cr * 9 + 2 * Cr; i += blockDim.x) acc27[i] = 0.f;
  __syncthreads();
  // ---- CE backward
  for (int i = threadIdx.x; i < T * Cr; i += blockDim.x) {
    const int t = i / Cr, j = i - t * Cr;
    float dr = 0.f;
    for (int c = 0; c < C; ++c) dr = fmaf(a.p2_expand[c * Cr + j], de2[t * C + c], dr);
    du[i] = a.u[f0 * Cr + i] > 0.f ? dr : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * Cr; i += blockDim.x) {
    const int t = i / Cr, j2 = i - t * Cr;          // ds[t][j2] = sum_{j,k} w[j][j2][k] du[t-k+1][j]
    float acc = 0.f;
    for (int k = 0; k < 3; ++k) {
      const int tt = t - k + 1;
      if (tt < 0 || tt >= T) continue;
      for (int j = 0; j < Cr; ++j) acc = fmaf(a.p2_conv1[(j * Cr + j2) * 3 + k], du[tt * Cr + j], acc);
    }
    ds[i] = acc;
  }
  // ---- ME: d pi, dd
  for (int i = threadIdx.x; i < T * Cr; i += blockDim.x) {
    const int t = i / Cr, j = i - t * Cr;
    float dpi = 0.f;
    if (t < T - 1)
      for (int c = 0; c < C; ++c) dpi = fmaf(a.p3_expand[c * Cr + j], de3[t * C + c], dpi);
    dd[i] = dpi * inv_hw;
    a.dd[f0 * Cr + i] = dd[i];
  }
  __syncthreads();
  // parameter gradients of the tiny layers (one atomic per element per clip)
  for (int i = threadIdx.x; i < C * Cr; i += blockDim.x) {
    const int c = i / Cr, j = i - c * Cr;
    float g2 = 0.f, g3 = 0.f, gs = 0.f;
    for (int t = 0; t < T; ++t) {
      g2 = fmaf(de2[t * C + c], rv[t * Cr + j], g2);
      g3 = fmaf(de3[t * C + c], a.pi[(f0 + t) * Cr + j], g3);
      gs = fmaf(ds[t * Cr + j], pm[t * C + c], gs);
    }
    atomicAdd(&a.d_p2_expand[c * Cr + j], g2);
    atomicAdd(&a.d_p3_expand[c * Cr + j], g3);
    atomicAdd(&a.d_p2_squeeze[j * C + c], gs);
  }
  for (int i = threadIdx.x; i < Cr * Cr * 3; i += blockDim.x) {
    const int k = i % 3, j2 = (i / 3) % Cr, j = i / (3 * Cr);
    float acc = 0.f;
    for (int t = 0; t < T; ++t) {
      const int tt = t + k - 1;
      if (tt < 0 || tt >= T) continue;
      acc = fmaf(du[t * Cr + j], sv[tt * Cr + j2], acc);
    }
    atomicAdd(&a.d_p2_conv1[i], acc);
  }
  // dpool[t][c] = sum_j W_sq[j][c] ds[t][j]   (gradient w.r.t. the spatial MEAN)
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) {
    const int t = i / C, c = i - t * C;
    float acc = 0.f;
    for (int j = 0; j < Cr; ++j) acc = fmaf(a.p2_squeeze[j * C + c], ds[t * Cr + j], acc);
    a.dpool[f0 * C + i] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// backward pass 2b: block = one FRAME (n*T blocks instead of n): everything that runs over pixels.
//   STE : da1 = dg1 * g1 (1 - g1) of frames t-1, t, t+1 and mrow of the same frames staged in shared memory;
//         dm[t] = conv3d^T(da1), dW_p1 += corr(mrow, da1[t])        (27 register accumulators per thread)
//   ME  : dx3[t][p][j] = -dd[t][j] + dd[t-1][j] * border_wsum_j(p);  the BatchNorm-backward sums and the
//         depthwise weight gradient only need seven per-(frame, j) reductions over pixels
//         (sum b, sum q, sum b*q, first/last row and column sums of q): no per-element atomics.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) action_bwd_frame_kernel(Act a) {
  extern __shared__ float smem[];
  const int Cr = a.cr, T = a.t, H = a.h, W = a.w, HW = H * W;
  float* da = smem;             // [3][HW]
  float* mr = da + 3 * HW;      // [3][HW]
  float* red = mr + 3 * HW;     // [27 + 7]
  const long long f = blockIdx.x;
  const int t = static_cast<int>(f % T);
  const long long m_t = f * HW;
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 3 * HW; i += blockDim.x) {
    const int fr = i / HW, p = i - fr * HW, tt = t + fr - 1;
    float d = 0.f, m = 0.f;
    if (tt >= 0 && tt < T) {
      const long long idx = m_t + static_cast<long long>(fr - 1) * HW + p;
      const float g = a.g1[idx];
      d = a.dg1[idx] * g * (1.f - g);
      m = a.mrow[idx];
    }
    da[i] = d;
    mr[i] = m;
  }
  if (tid < 34) red[tid] = 0.f;
  __syncthreads();
  {
    float w27[27], acc27[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) { w27[k] = a.p1_w[k]; acc27[k] = 0.f; }
    for (int p = tid; p < HW; p += blockDim.x) {
      const int h = p / W, w = p - h * W;
      const float ac = da[HW + p];
      float accm = 0.f;
#pragma unroll
      for (int dt = -1; dt <= 1; ++dt)
#pragma unroll
        for (int dh = -1; dh <= 1; ++dh)
#pragma unroll
          for (int dw = -1; dw <= 1; ++dw) {
            const int tap = (dt + 1) * 9 + (dh + 1) * 3 + dw + 1;
            const int hs = h - dh, ws = w - dw;            // transposed conv: the output position that used this input
            if (hs >= 0 && hs < H && ws >= 0 && ws < W) accm = fmaf(w27[tap], da[(1 - dt) * HW + hs * W + ws], accm);
            const int h2 = h + dh, w2 = w + dw;
            if (h2 >= 0 && h2 < H && w2 >= 0 && w2 < W) acc27[tap] = fmaf(ac, mr[(1 + dt) * HW + h2 * W + w2], acc27[tap]);
          }
      a.dm[m_t + p] = accm;
    }
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const float v = warp_sum(acc27[k]);
      if (lane == 0 && v != 0.f) atomicAdd(&red[k], v);
    }
  }
  __syncthreads();
  if (tid < 27) atomicAdd(&a.d_p1_w[tid], red[tid]);
  // ---- motion branch
  float* r7 = red + 27;
  for (int j = 0; j < Cr; ++j) {
    const float* w9 = a.p3_conv1 + j * 9;
    float sb = 0.f, sq = 0.f, sbq = 0.f, r0 = 0.f, rl = 0.f, c0 = 0.f, cl = 0.f;
    for (int p = tid; p < HW; p += blockDim.x) {
      const int h = p / W, w = p - h * W;
      const float q = a.q[(m_t + p) * Cr + j];
      const float b = border_wsum(w9, h, w, H, W);
      sb += b; sq += q; sbq = fmaf(b, q, sbq);
      if (h == 0) r0 += q;
      if (h == H - 1) rl += q;
      if (w == 0) c0 += q;
      if (w == W - 1) cl += q;
    }
    sb = warp_sum(sb); sq = warp_sum(sq); sbq = warp_sum(sbq);
    r0 = warp_sum(r0); rl = warp_sum(rl); c0 = warp_sum(c0); cl = warp_sum(cl);
    if (lane == 0) {
      atomicAdd(&r7[0], sb); atomicAdd(&r7[1], sq); atomicAdd(&r7[2], sbq);
      atomicAdd(&r7[3], r0); atomicAdd(&r7[4], rl); atomicAdd(&r7[5], c0); atomicAdd(&r7[6], cl);
    }
    __syncthreads();
    if (tid == 0) {
      const float dd_t = a.dd[f * Cr + j];
      const float ddp = t > 0 ? a.dd[(f - 1) * Cr + j] : 0.f;
      atomicAdd(&a.bn3_sums[j], static_cast<double>(fmaf(ddp, r7[0], -static_cast<float>(HW) * dd_t)));
      atomicAdd(&a.bn3_sums[Cr + j], static_cast<double>(fmaf(ddp, r7[2], -dd_t * r7[1])));
      if (t > 0 && ddp != 0.f) {
        // dW3[j][kh][kw] += dd[t-1][j] * sum over output positions of x3[t][h+kh-1, w+kw-1], x3 = q*sc + sh:
        // the visited inputs form the image minus one border row / column (inclusion-exclusion on q's sums)
        const float sc = a.bn3_scale[j], sh = a.bn3_shift[j];
        const float* qf = a.q + m_t * Cr + j;
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw) {
            const float ex_row = kh == 0 ? r7[4] : (kh == 2 ? r7[3] : 0.f);
            const float ex_col = kw == 0 ? r7[6] : (kw == 2 ? r7[5] : 0.f);
            float corner = 0.f;
            if (kh != 1 && kw != 1)
              corner = qf[(static_cast<long long>(kh == 0 ? H - 1 : 0) * W + (kw == 0 ? W - 1 : 0)) * Cr];
            const float rsum = r7[1] - ex_row - ex_col + corner;
            const int nrows = H - (kh != 1 ? 1 : 0), ncols = W - (kw != 1 ? 1 : 0);
            const float cnt = static_cast<float>((nrows > 0 ? nrows : 0) * (ncols > 0 ? ncols : 0));
            atomicAdd(&a.d_p3_conv1[j * 9 + kh * 3 + kw], ddp * fmaf(sc, rsum, sh * cnt));
          }
      }
#pragma unroll
      for (int k = 0; k < 7; ++k) r7[k] = 0.f;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// backward pass 3: dxs = gy*G + dm/C + dpool/HW + W_sq3^T dq,  dq = BN3-backward(dx3);  dW_sq3 += dq xs^T
// ------------------------------------------------------------------------------------------------
template <typename T, int CRMAX>
__global__ void __launch_bounds__(256)
action_bwd_dxs_kernel(Act a, const T* __restrict__ gy, const T* __restrict__ xs, T* __restrict__ dxs) {
  constexpr int V = VecOf<T>::N;
  extern __shared__ float smem[];
  const int C = a.c, Cr = a.cr, H = a.h, W = a.w, HW = H * W, CV = C / V, P = blockDim.x / CV;
  float* s_sq3 = smem;                 // [Cr][C]
  float* s_dsq = s_sq3 + Cr * C;       // [Cr][C]
  float* s_dq = s_dsq + Cr * C;        // [P][Cr]
  for (int i = threadIdx.x; i < Cr * C; i += blockDim.x) { s_sq3[i] = a.p3_squeeze[i]; s_dsq[i] = 0.f; }
  const int cv = threadIdx.x % CV, py = threadIdx.x / CV;
  const bool active = py < P;
  const int c0 = cv * V;
  const long long nt = blockIdx.x;
  const int t = static_cast<int>(nt % a.t);
  const float inv_c = 1.f / static_cast<float>(C), inv_hw = 1.f / static_cast<float>(HW);
  float g23[V], dpl[V], dacc[CRMAX][V];
  if (active) {
    float g2[V], g3[V];
    load_vec<float, V>(a.g2 + nt * C + c0, g2);
    load_vec<float, V>(a.g3 + nt * C + c0, g3);
    load_vec<float, V>(a.dpool + nt * C + c0, dpl);
#pragma unroll
    for (int i = 0; i < V; ++i) { g23[i] = 3.f + g2[i] + g3[i]; dpl[i] *= inv_hw; }
  }
#pragma unroll
  for (int j = 0; j < CRMAX; ++j)
#pragma unroll
    for (int i = 0; i < V; ++i) dacc[j][i] = 0.f;
  for (int r0 = 0; r0 < HW; r0 += P) {
    const int row = r0 + py;
    const bool live = active && row < HW;
    const long long m = nt * HW + row;
    __syncthreads();
    // dq for the P rows of this iteration (threads cv < Cr of each row)
    if (live) {
      const int h = row / W, w = row - h * W;
      for (int j = cv; j < Cr; j += CV) {
        const float ddc = a.dd[nt * Cr + j];
        const float ddp = t > 0 ? a.dd[(nt - 1) * Cr + j] : 0.f;
        const float dx3 = -ddc + ddp * border_wsum(a.p3_conv1 + j * 9, h, w, H, W);
        s_dq[py * Cr + j] = fmaf(a.bn3_ca[j], dx3, fmaf(a.bn3_cb[j], a.q[m * Cr + j], a.bn3_cc[j]));
      }
    }
    __syncthreads();
    if (live) {
      float g[V], v[V], o[V];
      load_vec<T, V>(gy + m * C + c0, g);
      load_vec<T, V>(xs + m * C + c0, v);
      const float g1 = a.g1[m], dmv = a.dm[m] * inv_c;
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = fmaf(g[i], g23[i] + g1, dmv + dpl[i]);
#pragma unroll
      for (int j = 0; j < CRMAX; ++j) {
        if (j < Cr) {
          const float dq = s_dq[py * Cr + j];
#pragma unroll
          for (int i = 0; i < V; ++i) {
            o[i] = fmaf(dq, s_sq3[j * C + c0 + i], o[i]);
            dacc[j][i] = fmaf(dq, v[i], dacc[j][i]);
          }
        }
      }
      store_vec<T, V>(dxs + m * C + c0, o);
    }
  }
  if (active) {
#pragma unroll
    for (int j = 0; j < CRMAX; ++j)
      if (j < Cr)
#pragma unroll
        for (int i = 0; i < V; ++i) atomicAdd(&s_dsq[j * C + c0 + i], dacc[j][i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cr * C; i += blockDim.x) atomicAdd(&a.d_p3_squeeze[i], s_dsq[i]);
}

// ------------------------------------------------------------------------------------------------
// backward pass 4: adjoint of the temporal FIR (+ residual gradient) and its weight gradient
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
action_fir_bwd_kernel(Act a, const T* __restrict__ dxs, const T* __restrict__ x, const T* __restrict__ addend,
                      T* __restrict__ dx) {
  constexpr int V = VecOf<T>::N;
  extern __shared__ float smem[];   // [C][3]
  const int C = a.c, HW = a.h * a.w, CV = C / V, P = blockDim.x / CV;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) smem[i] = 0.f;
  __syncthreads();
  const int cv = threadIdx.x % CV, py = threadIdx.x / CV;
  const bool active = py < P;
  const int c0 = cv * V;
  const long long nt = blockIdx.x;
  const int t = static_cast<int>(nt % a.t);
  const bool has_prev = t > 0, has_next = t < a.t - 1;
  float w0[V], w1[V], w2[V], a0[V], a1[V], a2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    w0[i] = a.shift_w[(c0 + i) * 3 + 0];
    w1[i] = a.shift_w[(c0 + i) * 3 + 1];
    w2[i] = a.shift_w[(c0 + i) * 3 + 2];
    a0[i] = a1[i] = a2[i] = 0.f;
  }
  const long long fe = static_cast<long long>(HW) * C;
  if (active) {
    for (int row = py; row < HW; row += P) {
      const long long off = (nt * HW + row) * C + c0;
      float dc[V], dp[V], dn[V], xc[V], xp[V], xn[V], o[V];
      load_vec<T, V>(dxs + off, dc);
      load_vec<T, V>(x + off, xc);
#pragma unroll
      for (int i = 0; i < V; ++i) dp[i] = dn[i] = xp[i] = xn[i] = 0.f;
      if (has_prev) { load_vec<T, V>(dxs + off - fe, dp); load_vec<T, V>(x + off - fe, xp); }
      if (has_next) { load_vec<T, V>(dxs + off + fe, dn); load_vec<T, V>(x + off + fe, xn); }
      // xs[t] = w0 x[t-1] + w1 x[t] + w2 x[t+1]  =>  dx[t] = w0 dxs[t+1] + w1 dxs[t] + w2 dxs[t-1]
#pragma unroll
      for (int i = 0; i < V; ++i) {
        o[i] = fmaf(w0[i], dn[i], fmaf(w1[i], dc[i], w2[i] * dp[i]));
        a0[i] = fmaf(dc[i], xp[i], a0[i]);
        a1[i] = fmaf(dc[i], xc[i], a1[i]);
        a2[i] = fmaf(dc[i], xn[i], a2[i]);
      }
      if (addend) {
        float ad[V];
        load_vec<T, V>(addend + off, ad);
#pragma unroll
        for (int i = 0; i < V; ++i) o[i] += ad[i];
      }
      store_vec<T, V>(dx + off, o);
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      atomicAdd(&smem[(c0 + i) * 3 + 0], a0[i]);
      atomicAdd(&smem[(c0 + i) * 3 + 1], a1[i]);
      atomicAdd(&smem[(c0 + i) * 3 + 2], a2[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) atomicAdd(&a.d_shift_w[i], smem[i]);
}

static int act_check(const Act* a, int dtype) {
  const int es = esize_of(dtype);
  if (es == 0) return EHGR_E_DTYPE;
  if (!a) return EHGR_E_NULL;
  const int V = 16 / es;
  if (a->n <= 0 || a->t <= 0 || a->h <= 0 || a->w <= 0 || a->c <= 0 || a->cr <= 0 || (a->c % V)) return EHGR_E_SHAPE;
  if (a->c / V > 256 || a->cr > 16) return EHGR_E_UNSUPPORTED;
  return EHGR_OK;
}

}  // namespace ehgr

using namespace ehgr;

extern "C" int ehgr_action_xs(const ehgr_action* a, const void* x, void* xs, int dtype, ehgr_stream_t stream) {
  if (int st = act_check(a, dtype)) return st;
  if (!x || !xs || !a->shift_w || !a->p3_squeeze || !a->mrow || !a->pool || !a->q || !a->qstats) return EHGR_E_NULL;
  const int V = 16 / esize_of(dtype), CV = a->c / V;
  const unsigned frames = static_cast<unsigned>(a->n) * a->t;
  const int hw = a->h * a->w;
  cudaStream_t s = as_stream(stream);
  cudaMemsetAsync(a->pool, 0, static_cast<size_t>(frames) * a->c * sizeof(float), s);      // pool is accumulated by atomics
  if (CV <= 32) {
    // lane-group kernel: a CTA = (clip, row range), ~4 CTAs per SM in total, at least one pass of rows per CTA
    int lg = 0;
    while ((1 << lg) < CV) ++lg;
    const int rows_per_pass = 8 * (32 >> lg);
    int splits = (4 * kNumSMs) / std::max(1, a->n);        // rounded DOWN: all CTAs resident in one wave (4 per SM)
    splits = std::max(1, std::min(splits, (hw + 2 * rows_per_pass - 1) / (2 * rows_per_pass)));
    const size_t smem = (static_cast<size_t>(a->cr) * a->c + 2 * a->cr) * sizeof(float);
    const unsigned grid = static_cast<unsigned>(a->n) * splits;
    if (dtype == EHGR_F32)
      action_xs_kernel<float><<<grid, 256, smem, s>>>(*a, static_cast<const float*>(x), static_cast<float*>(xs), splits);
    else
      action_xs_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(*a, static_cast<const __nv_bfloat16*>(x),
                                                              static_cast<__nv_bfloat16*>(xs), splits);
    return launch_status();
  }
  const int P = 256 / CV;
  const size_t smem = (static_cast<size_t>(a->cr) * a->c + static_cast<size_t>(P) * (a->cr + 1) + a->c + 2 * a->cr) * sizeof(float);
  int splits = static_cast<int>((4 * kNumSMs + frames - 1) / std::max(1u, frames));      // ~4 CTAs per SM in total
  splits = std::max(1, std::min(std::min(splits, 16), (hw + P - 1) / P));
  const dim3 grid(frames, splits);
  if (dtype == EHGR_F32) action_xs_generic_kernel<float><<<grid, 256, smem, s>>>(*a, static_cast<const float*>(x), static_cast<float*>(xs));
  else action_xs_generic_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(*a, static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(xs));
  return launch_status();
}

extern "C" int ehgr_action_gates(const ehgr_action* a, ehgr_stream_t stream) {
  if (int st = act_check(a, EHGR_F32)) return st;
  if (!a->mrow || !a->pool || !a->q || !a->bn3_scale || !a->bn3_shift || !a->g1 || !a->g2 || !a->g3 || !a->s || !a->u || !a->pi)
    return EHGR_E_NULL;
  const size_t smem = (static_cast<size_t>(a->t) * a->c + 3 * static_cast<size_t>(a->t) * a->cr) * sizeof(float);
  const size_t smem_f = (3 * static_cast<size_t>(a->h) * a->w + 8) * sizeof(float);
  if (smem_f > 200 * 1024) return EHGR_E_UNSUPPORTED;
  ensure_smem(action_gates_frame_kernel, smem_f);
  action_gates_frame_kernel<<<static_cast<unsigned>(a->n) * a->t, 256, smem_f, as_stream(stream)>>>(*a);
  if (int st = launch_status()) return st;
  action_gates_kernel<<<a->n, 256, smem, as_stream(stream)>>>(*a);
  return launch_status();
}

extern "C" int ehgr_action_bwd_reduce(const ehgr_action* a, const void* gy, const void* xs, int dtype,
                                      ehgr_stream_t stream) {
  if (int st = act_check(a, dtype)) return st;
  if (!gy || !xs || !a->dg1 || !a->dgc) return EHGR_E_NULL;
  const int V = 16 / esize_of(dtype), CV = a->c / V;
  const unsigned frames = static_cast<unsigned>(a->n) * a->t;
  cudaStream_t s = as_stream(stream);
  if (CV <= 32) {
    int lg = 0;
    while ((1 << lg) < CV) ++lg;
    const int rows_per_pass = 8 * (32 >> lg), hw = a->h * a->w;
    int splits = static_cast<int>((4 * kNumSMs + frames - 1) / std::max(1u, frames));
    splits = std::max(1, std::min(splits, (hw + 2 * rows_per_pass - 1) / (2 * rows_per_pass)));
    cudaMemsetAsync(a->dgc, 0, static_cast<size_t>(frames) * a->c * sizeof(float), s);     // accumulated by atomics
    if (dtype == EHGR_F32)
      action_bwd_reduce_kernel<float><<<frames * splits, 256, 0, s>>>(*a, static_cast<const float*>(gy),
                                                                      static_cast<const float*>(xs), splits);
    else
      action_bwd_reduce_kernel<__nv_bfloat16><<<frames * splits, 256, 0, s>>>(*a, static_cast<const __nv_bfloat16*>(gy),
                                                                              static_cast<const __nv_bfloat16*>(xs), splits);
    return launch_status();
  }
  const int P = 256 / CV;
  const size_t smem = (static_cast<size_t>(P) + a->c) * sizeof(float);
  if (dtype == EHGR_F32)
    action_bwd_reduce_generic_kernel<float><<<frames, 256, smem, s>>>(*a, static_cast<const float*>(gy), static_cast<const float*>(xs));
  else
    action_bwd_reduce_generic_kernel<__nv_bfloat16><<<frames, 256, smem, s>>>(*a, static_cast<const __nv_bfloat16*>(gy),
                                                                            static_cast<const __nv_bfloat16*>(xs));
  return launch_status();
}

extern "C" int ehgr_action_bwd_small(const ehgr_action* a, ehgr_stream_t stream) {
  if (int st = act_check(a, EHGR_F32)) return st;
  if (!a->dg1 || !a->dgc || !a->dm || !a->dpool || !a->dd || !a->bn3_sums || !a->d_p1_w || !a->d_p2_squeeze ||
      !a->d_p2_conv1 || !a->d_p2_expand || !a->d_p3_conv1 || !a->d_p3_expand)
    return EHGR_E_NULL;
  const size_t smem = (3 * static_cast<size_t>(a->t) * a->c + 5 * static_cast<size_t>(a->t) * a->cr + 27 + 11 * a->cr) * sizeof(float);
  action_bwd_clip_kernel<<<a->n, 256, smem, as_stream(stream)>>>(*a);
  if (int st = launch_status()) return st;
  const size_t smem_f = (6 * static_cast<size_t>(a->h) * a->w + 40) * sizeof(float);
  if (smem_f > 200 * 1024) return EHGR_E_UNSUPPORTED;
  ensure_smem(action_bwd_frame_kernel, smem_f);
  action_bwd_frame_kernel<<<static_cast<unsigned>(a->n) * a->t, 256, smem_f, as_stream(stream)>>>(*a);
  return launch_status();
}

extern "C" int ehgr_action_bwd_dxs(const ehgr_action* a, const void* gy, const void* xs, void* dxs, int dtype,
                                   ehgr_stream_t stream) {
  if (int st = act_check(a, dtype)) return st;
  if (!gy || !xs || !dxs || !a->bn3_ca || !a->bn3_cb || !a->bn3_cc || !a->d_p3_squeeze) return EHGR_E_NULL;
  const int V = 16 / esize_of(dtype), CV = a->c / V, P = 256 / CV;
  const size_t smem = (2 * static_cast<size_t>(a->cr) * a->c + static_cast<size_t>(P) * a->cr) * sizeof(float);
  const unsigned grid = static_cast<unsigned>(a->n) * a->t;
  cudaStream_t s = as_stream(stream);
#define EHGR_DXS(TT, CRM) action_bwd_dxs_kernel<TT, CRM><<<grid, 256, smem, s>>>(*a, static_cast<const TT*>(gy), static_cast<const TT*>(xs), static_cast<TT*>(dxs))
  if (dtype == EHGR_F32) {
    if (a->cr <= 4) EHGR_DXS(float, 4); else if (a->cr <= 10) EHGR_DXS(float, 10); else EHGR_DXS(float, 16);
  } else {
    if (a->cr <= 4) EHGR_DXS(__nv_bfloat16, 4); else if (a->cr <= 10) EHGR_DXS(__nv_bfloat16, 10); else EHGR_DXS(__nv_bfloat16, 16);
  }
#undef EHGR_DXS
  return launch_status();
}

extern "C" int ehgr_action_fir_bwd(const ehgr_action* a, const void* dxs, const void* x, const void* addend, void* dx,
                                   int dtype, ehgr_stream_t stream) {
  if (int st = act_check(a, dtype)) return st;
  if (!dxs || !x || !dx || !a->d_shift_w) return EHGR_E_NULL;
  const size_t smem = 3 * static_cast<size_t>(a->c) * sizeof(float);
  const unsigned grid = static_cast<unsigned>(a->n) * a->t;
  cudaStream_t s = as_stream(stream);
  if (dtype == EHGR_F32)
    action_fir_bwd_kernel<float><<<grid, 256, smem, s>>>(*a, static_cast<const float*>(dxs), static_cast<const float*>(x),
                                                         static_cast<const float*>(addend), static_cast<float*>(dx));
  else
    action_fir_bwd_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(*a, static_cast<const __nv_bfloat16*>(dxs),
                                                                 static_cast<const __nv_bfloat16*>(x),
                                                                 static_cast<const __nv_bfloat16*>(addend),
                                                                 static_cast<__nv_bfloat16*>(dx));
  return launch_status();
}
