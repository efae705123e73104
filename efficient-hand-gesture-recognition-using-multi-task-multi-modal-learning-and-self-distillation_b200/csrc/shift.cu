// shift.cu — K1: temporal shift forward / backward (reference: models/temporal_shift.py:27-46,
// models/action.py:135-154; the in-place variant models/temporal_shift.py:49-76 computes the same
// thing).  A pure gather: every output vector is read from frame t+1, t-1 or t of the same clip (or
// is zero at the clip boundary), so it is bit-exact in every dtype.  One read + one write per
// element; HBM-bound.
//
// NCHW: a frame is C*HW contiguous elements and the two shifted folds are the first 2*fold*HW of
//       them, so the kernel works on opaque VB-byte units (VB = widest of 16/8/4/2 bytes that
//       divides the frame size, both fold boundaries and the base addresses) and never straddles a
//       fold boundary.
// NHWC: channels are innermost; a 16-byte vector of channels may straddle a fold boundary
//       (fold = 3, 12, 20 ...), in which case the (at most three) source vectors are loaded and
//       selected per element.
#include "common.cuh"

namespace ehgr {

template <int VB, int UNROLL>
__global__ void __launch_bounds__(256)
shift_nchw_kernel(const typename Vec<VB>::type* __restrict__ x, typename Vec<VB>::type* __restrict__ out,
                  int T, uint32_t frame_units, uint32_t b1u, uint32_t b2u, uint32_t chunks_per_frame,
                  int dir) {
  using V = typename Vec<VB>::type;
  const uint32_t frame = blockIdx.x / chunks_per_frame;
  const uint32_t chunk = blockIdx.x - frame * chunks_per_frame;
  const int t = static_cast<int>(frame % static_cast<uint32_t>(T));
  // class 0 (c < fold) reads frame t+dir, class 1 (fold <= c < 2 fold) reads frame t-dir.
  const bool ok_fwd = (dir > 0) ? (t < T - 1) : (t > 0);
  const bool ok_bwd = (dir > 0) ? (t > 0) : (t < T - 1);
  const size_t base = static_cast<size_t>(frame) * frame_units;
  const long long step = static_cast<long long>(dir) * static_cast<long long>(frame_units);
  const uint32_t u0 = chunk * (256u * UNROLL) + threadIdx.x;

  V v[UNROLL];
#pragma unroll
  for (int j = 0; j < UNROLL; ++j) {
    const uint32_t u = u0 + j * 256u;
    v[j] = zero_of(V{});
    if (u < frame_units) {
      const int cls = (u < b1u) ? 0 : (u < b2u) ? 1 : 2;
      const bool ok = (cls == 2) || (cls == 0 ? ok_fwd : ok_bwd);
      const long long off = (cls == 0) ? step : (cls == 1) ? -step : 0;
      if (ok) v[j] = ld_stream(x + (static_cast<long long>(base + u) + off));
    }
  }
#pragma unroll
  for (int j = 0; j < UNROLL; ++j) {
    const uint32_t u = u0 + j * 256u;
    if (u < frame_units) st_stream(out + base + u, v[j]);
  }
}

// NHWC: E = element as raw bits (uint16_t / uint32_t), VE elements per vector (VE*sizeof(E) <= 16).
template <typename E, int VE>
__global__ void __launch_bounds__(256)
shift_nhwc_kernel(const E* __restrict__ x, E* __restrict__ out, int T, int C, uint32_t vec_per_pixel,
                  uint32_t vec_per_frame, uint32_t chunks_per_frame, int fold, int dir) {
  using V = typename Vec<VE * sizeof(E)>::type;
  const uint32_t frame = blockIdx.x / chunks_per_frame;
  const uint32_t chunk = blockIdx.x - frame * chunks_per_frame;
  const uint32_t u = chunk * 256u + threadIdx.x;
  if (u >= vec_per_frame) return;
  const int t = static_cast<int>(frame % static_cast<uint32_t>(T));
  const bool ok_fwd = (dir > 0) ? (t < T - 1) : (t > 0);
  const bool ok_bwd = (dir > 0) ? (t > 0) : (t < T - 1);
  const uint32_t cv = u % vec_per_pixel;
  const int c0 = static_cast<int>(cv) * VE;
  const size_t idx = static_cast<size_t>(frame) * vec_per_frame + u;  // in vectors
  const long long step = static_cast<long long>(dir) * static_cast<long long>(vec_per_frame);
  const V* xv = reinterpret_cast<const V*>(x);
  V* ov = reinterpret_cast<V*>(out);

  auto cls_of = [fold](int c) { return c < fold ? 0 : (c < 2 * fold ? 1 : 2); };
  const int cls_lo = cls_of(c0), cls_hi = cls_of(c0 + VE - 1);
  V r = zero_of(V{});
  if (cls_lo == cls_hi) {
    const bool ok = (cls_lo == 2) || (cls_lo == 0 ? ok_fwd : ok_bwd);
    const long long off = (cls_lo == 0) ? step : (cls_lo == 1) ? -step : 0;
    if (ok) r = ld_stream(xv + (static_cast<long long>(idx) + off));
  } else {
    // straddles a fold boundary: select per element from up to three source frames
    V src[3];
    src[0] = ok_fwd ? ld_stream(xv + (static_cast<long long>(idx) + step)) : zero_of(V{});
    src[1] = ok_bwd ? ld_stream(xv + (static_cast<long long>(idx) - step)) : zero_of(V{});
    src[2] = ld_stream(xv + idx);
    const E* s0 = reinterpret_cast<const E*>(&src[0]);
    const E* s1 = reinterpret_cast<const E*>(&src[1]);
    const E* s2 = reinterpret_cast<const E*>(&src[2]);
    E* d = reinterpret_cast<E*>(&r);
#pragma unroll
    for (int j = 0; j < VE; ++j) {
      const int k = cls_of(c0 + j);
      d[j] = (k == 0) ? s0[j] : (k == 1) ? s1[j] : s2[j];
    }
  }
  st_stream(ov + idx, r);
}

static int gcd_pow2_bytes(unsigned long long v, int cap) {
  int b = cap;
  while (b > 1 && (v % static_cast<unsigned long long>(b)) != 0) b >>= 1;
  return b;
}

template <int VB>
static int launch_nchw(const void* x, void* out, long long frames, int T, long long frame_bytes,
                       long long b1, long long b2, int dir, cudaStream_t s) {
  constexpr int UNROLL = 4;
  const uint32_t fu = static_cast<uint32_t>(frame_bytes / VB);
  const uint32_t cpf = static_cast<uint32_t>(cdiv(fu, 256 * UNROLL));
  const long long blocks = frames * cpf;
  if (blocks > 0x7fffffffLL) return EHGR_E_SHAPE;
  using V = typename Vec<VB>::type;
  shift_nchw_kernel<VB, UNROLL><<<static_cast<unsigned>(blocks), 256, 0, s>>>(
      static_cast<const V*>(x), static_cast<V*>(out), T, fu, static_cast<uint32_t>(b1 / VB),
      static_cast<uint32_t>(b2 / VB), cpf, dir);
  return launch_status();
}

template <typename E, int VE>
static int launch_nhwc(const void* x, void* out, long long frames, int T, int C, int hw, int fold, int dir,
                       cudaStream_t s) {
  const uint32_t vpp = static_cast<uint32_t>(C / VE);
  const long long vpf_ll = static_cast<long long>(vpp) * hw;
  if (vpf_ll > 0x7fffffffLL) return EHGR_E_SHAPE;
  const uint32_t vpf = static_cast<uint32_t>(vpf_ll);
  const uint32_t cpf = static_cast<uint32_t>(cdiv(vpf, 256));
  const long long blocks = frames * cpf;
  if (blocks > 0x7fffffffLL) return EHGR_E_SHAPE;
  shift_nhwc_kernel<E, VE><<<static_cast<unsigned>(blocks), 256, 0, s>>>(
      static_cast<const E*>(x), static_cast<E*>(out), T, C, vpp, vpf, cpf, fold, dir);
  return launch_status();
}

static int shift_dispatch(const void* x, void* out, int n_batch, int n_segment, int c, int hw, int fold,
                          int dtype, int layout, int dir, ehgr_stream_t stream) {
  if (!x || !out) return EHGR_E_NULL;
  const int es = esize_of(dtype);
  if (es == 0 || (layout != EHGR_NCHW && layout != EHGR_NHWC)) return EHGR_E_DTYPE;
  if (n_batch < 0 || n_segment <= 0 || c <= 0 || hw <= 0 || fold < 0 || 2LL * fold > c) return EHGR_E_SHAPE;
  if (!aligned_to(x, es) || !aligned_to(out, es)) return EHGR_E_ALIGN;
  if (x == out) return EHGR_E_UNSUPPORTED;
  const long long frames = static_cast<long long>(n_batch) * n_segment;
  if (frames == 0) return EHGR_OK;
  cudaStream_t s = as_stream(stream);
  const long long frame_elems = static_cast<long long>(c) * hw;
  if (frame_elems * es > 0x7fffffffLL * 2) return EHGR_E_SHAPE;

  if (layout == EHGR_NCHW) {
    const long long frame_bytes = frame_elems * es;
    const long long b1 = static_cast<long long>(fold) * hw * es, b2 = 2 * b1;
    unsigned long long g = static_cast<unsigned long long>(frame_bytes) |
                           reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out);
    if (fold > 0) g |= static_cast<unsigned long long>(b1);
    // lowest set bit of the OR = largest power of two dividing all of them
    const int vb = gcd_pow2_bytes(g & (~g + 1ULL), 16);
    switch (vb) {
      case 16: return launch_nchw<16>(x, out, frames, n_segment, frame_bytes, b1, b2, dir, s);
      case 8: return launch_nchw<8>(x, out, frames, n_segment, frame_bytes, b1, b2, dir, s);
      case 4: return launch_nchw<4>(x, out, frames, n_segment, frame_bytes, b1, b2, dir, s);
      default: return launch_nchw<2>(x, out, frames, n_segment, frame_bytes, b1, b2, dir, s);
    }
  }
  // NHWC
  unsigned long long g = static_cast<unsigned long long>(c) * es | reinterpret_cast<uintptr_t>(x) |
                         reinterpret_cast<uintptr_t>(out);
  const int vb = gcd_pow2_bytes(g & (~g + 1ULL), 16);
  if (es == 2) {
    switch (vb) {
      case 16: return launch_nhwc<uint16_t, 8>(x, out, frames, n_segment, c, hw, fold, dir, s);
      case 8: return launch_nhwc<uint16_t, 4>(x, out, frames, n_segment, c, hw, fold, dir, s);
      case 4: return launch_nhwc<uint16_t, 2>(x, out, frames, n_segment, c, hw, fold, dir, s);
      default: return launch_nhwc<uint16_t, 1>(x, out, frames, n_segment, c, hw, fold, dir, s);
    }
  }
  switch (vb) {
    case 16: return launch_nhwc<uint32_t, 4>(x, out, frames, n_segment, c, hw, fold, dir, s);
    case 8: return launch_nhwc<uint32_t, 2>(x, out, frames, n_segment, c, hw, fold, dir, s);
    default: return launch_nhwc<uint32_t, 1>(x, out, frames, n_segment, c, hw, fold, dir, s);
  }
}

}  // namespace ehgr

extern "C" int ehgr_temporal_shift_fwd(const void* x, void* out, int n_batch, int n_segment, int c, int hw,
                                       int fold, int dtype, int layout, ehgr_stream_t stream) {
  return ehgr::shift_dispatch(x, out, n_batch, n_segment, c, hw, fold, dtype, layout, +1, stream);
}

extern "C" int ehgr_temporal_shift_bwd(const void* grad_out, void* grad_in, int n_batch, int n_segment,
                                       int c, int hw, int fold, int dtype, int layout,
                                       ehgr_stream_t stream) {
  return ehgr::shift_dispatch(grad_out, grad_in, n_batch, n_segment, c, hw, fold, dtype, layout, -1, stream);
}
