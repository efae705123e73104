// rowop.cuh — the "row operand": how a kernel reads an activation row [M, C] (NHWC, M = NT*H*W).
//
// Activations are stored RAW (the convolution output before BatchNorm); the BatchNorm affine, the
// ReLU6, the temporal shift and — in backward — the BatchNorm-backward combination are applied by
// the CONSUMER kernel while it loads the row.  That removes every stand-alone BN / ReLU6 / shift pass
// of the reference (archs/mobilenet_v2.py:40-59 = conv, BN, ReLU6 as three kernels each).
//
//   PLAIN  : v = in1
//   AFFINE : v = in1*scale[c] + shift[c]; relu6 ? clamp(v,0,6)                  (lazy BN(+ReLU6))
//   SHIFT  : v = in1 of frame t+1 (c < fold), t-1 (fold <= c < 2 fold), t (rest), zero at clip ends
//            (TemporalShift.shift, models/temporal_shift.py:40-44)
//   BNBWD  : v = ca[c]*mask*in1 + cb[c]*in2 + cc[c],  mask = relu6 ? (0 < in2*scale+shift < 6) : 1
//            in1 = gradient w.r.t. the post-activation, in2 = raw forward output of that layer:
//            this is d(loss)/d(raw) of conv->BN(->ReLU6) with batch statistics (cb, cc != 0) or frozen
//            statistics (cb = cc = 0).
//   CONV3  : the im2col row of a dense 3x3 convolution (pad 1, stride 1; optionally over the nearest-x2 upsampled
//            tensor) with the producer's lazy BatchNorm+ReLU applied to the gathered pixels — GEMM family only
//            (include/ehgr_b200.h; models/models_MTMM.py:129-155).
#pragma once
#include "common.cuh"

namespace ehgr {

using RowOp = ehgr_rowop;

template <typename T>
struct VecOf;  // 16-byte vector of T
template <>
struct VecOf<float> { static constexpr int N = 4; };
template <>
struct VecOf<__nv_bfloat16> { static constexpr int N = 8; };

// ---- raw vector loads / stores of NV elements (NV*sizeof(T) in {8,16}) as floats ---------------
template <typename T, int NV>
__device__ __forceinline__ void load_vec(const T* __restrict__ p, float (&v)[NV]);

template <>
__device__ __forceinline__ void load_vec<float, 4>(const float* __restrict__ p, float (&v)[4]) {
  const float4 r = *reinterpret_cast<const float4*>(p);
  v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
template <>
__device__ __forceinline__ void load_vec<float, 8>(const float* __restrict__ p, float (&v)[8]) {
  const float4 r0 = *reinterpret_cast<const float4*>(p);
  const float4 r1 = *reinterpret_cast<const float4*>(p + 4);
  v[0] = r0.x; v[1] = r0.y; v[2] = r0.z; v[3] = r0.w;
  v[4] = r1.x; v[5] = r1.y; v[6] = r1.z; v[7] = r1.w;
}
template <>
__device__ __forceinline__ void load_vec<__nv_bfloat16, 4>(const __nv_bfloat16* __restrict__ p, float (&v)[4]) {
  const uint2 r = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&r.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&r.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
template <>
__device__ __forceinline__ void load_vec<__nv_bfloat16, 8>(const __nv_bfloat16* __restrict__ p, float (&v)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    v[2 * i] = __low2float(a);
    v[2 * i + 1] = __high2float(a);
  }
}

template <typename T, int NV>
__device__ __forceinline__ void store_vec(T* __restrict__ p, const float (&v)[NV]);

template <>
__device__ __forceinline__ void store_vec<float, 4>(float* __restrict__ p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void store_vec<float, 8>(float* __restrict__ p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// ---- packed helpers: a bf16x2 word <-> a float2 (Blackwell FFMA2 / FADD2 operate on register pairs) ----
__device__ __forceinline__ float2 bf2_to_f2(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
// round-to-nearest pack with the negative half clamped to zero (ReLU folded into the conversion)
__device__ __forceinline__ uint32_t f2_to_bf2_relu(float2 v) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(v.y), "f"(v.x));
  return d;
}
__device__ __forceinline__ uint32_t bf2_min6(uint32_t w) {
  uint32_t d;
  asm("min.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(w), "r"(0x40C040C0u));   // 6.0 is exact in bf16
  return d;
}
// The `relu6` field of a row operand is an activation code: 0 none, 1 ReLU6 (MobileNetV2), 2 plain ReLU (the SepConv
// exit heads and the depth decoder, models/models_SD.py:81-101, models/models_MTMM.py:129-155).  Everything that
// clamps or masks uses the upper bound relu_hi(): 6 or +infinity.
__device__ __forceinline__ float relu_hi(int act) { return act == 2 ? INFINITY : 6.f; }
__device__ __forceinline__ uint32_t bf2_min_hi(uint32_t w, int act) { return act == 2 ? w : bf2_min6(w); }

template <>
__device__ __forceinline__ void store_vec<__nv_bfloat16, 4>(__nv_bfloat16* __restrict__ p, const float (&v)[4]) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
}
template <>
__device__ __forceinline__ void store_vec<__nv_bfloat16, 8>(__nv_bfloat16* __restrict__ p, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                            pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// value as the consumer's storage type would hold it (bf16 storage rounds, fp32 does not)
template <typename T>
__device__ __forceinline__ float round_to(float v);
template <>
__device__ __forceinline__ float round_to<float>(float v) { return v; }
template <>
__device__ __forceinline__ float round_to<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// ---- CONV3 geometry ---------------------------------------------------------------------------------
struct Conv3Tap { int dy, dx, c; };
// column k of the implicit GEMM -> (dy, dx) in {-1,0,1}^2 and the channel offset inside the pixel
__device__ __forceinline__ Conv3Tap conv3_tap(const RowOp& op, int k) {
  const int tap = k / op.cv_cin;
  const int ty = tap / 3;
  return Conv3Tap{ty - 1, tap - ty * 3 - 1, k - tap * op.cv_cin};
}
// element offset of pixel (hs, ws) of frame fr — coordinates on the OUTPUT grid — inside the stored tensor
__device__ __forceinline__ long long conv3_src(const RowOp& op, int fr, int hs, int ws) {
  const int Hs = op.cv_h >> op.cv_up, Ws = op.cv_w >> op.cv_up;
  return ((static_cast<long long>(fr) * Hs + (hs >> op.cv_up)) * Ws + (ws >> op.cv_up)) * op.cv_cin;
}

// ---- the row operand ------------------------------------------------------------------------------
// Loads NV consecutive channels [c0, c0+NV) of row m (C channels per row) as fp32 with the operand's
// transformation applied.  The caller guarantees c0 % NV == 0, C % NV == 0 and 0 <= m < M.
// kGate = false compiles the GATE mode out (kernels that never take a gated operand).
template <typename T, int NV, bool kGate = true>
__device__ __forceinline__ void load_row(const RowOp& op, long long m, int c0, int C, float (&v)[NV]) {
  const T* in1 = static_cast<const T*>(op.in1);
  const long long off = m * C + c0;
  if (op.mode == EHGR_ROW_PLAIN) {
    load_vec<T, NV>(in1 + off, v);
  } else if (op.mode == EHGR_ROW_AFFINE) {
    load_vec<T, NV>(in1 + off, v);
    float s[NV], b[NV];
    load_vec<float, NV>(op.scale + c0, s);
    load_vec<float, NV>(op.shift + c0, b);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float z = fmaf(v[i], s[i], b[i]);
      if (op.relu6) z = fminf(fmaxf(z, 0.f), relu_hi(op.relu6));
      v[i] = z;
    }
  } else if (op.mode == EHGR_ROW_SHIFT) {
    const long long frame = m / op.hw;
    const int t = static_cast<int>(frame % op.n_segment);
    const int dir = op.shift_dir < 0 ? -1 : 1;
    const long long step = static_cast<long long>(dir) * op.hw * C;
    const int fold = op.fold;
    // class 0 reads frame t+dir, class 1 reads frame t-dir
    const bool has_next = dir > 0 ? (t < op.n_segment - 1) : (t > 0);
    const bool has_prev = dir > 0 ? (t > 0) : (t < op.n_segment - 1);
    const int lo = c0, hi = c0 + NV - 1;
    auto cls_of = [fold](int c) { return c < fold ? 0 : (c < 2 * fold ? 1 : 2); };
    const int cl = cls_of(lo), ch = cls_of(hi);
    if (cl == ch) {
      const bool ok = cl == 2 || (cl == 0 ? has_next : has_prev);
      if (ok) {
        load_vec<T, NV>(in1 + off + (cl == 0 ? step : cl == 1 ? -step : 0), v);
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = 0.f;
      }
    } else {
      float a[NV], b[NV], c[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) a[i] = b[i] = 0.f;
      if (has_next) load_vec<T, NV>(in1 + off + step, a);
      if (has_prev) load_vec<T, NV>(in1 + off - step, b);
      load_vec<T, NV>(in1 + off, c);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int k = cls_of(c0 + i);
        v[i] = k == 0 ? a[i] : (k == 1 ? b[i] : c[i]);
      }
    }
  } else if (kGate && op.mode == EHGR_ROW_CONV3) {
    // column c0 of the implicit GEMM = (tap, channel); row m = (frame, ho, wo) of the output grid
    const Conv3Tap tp = conv3_tap(op, c0);
    const int fr = static_cast<int>(m / op.hw);
    const int rem = static_cast<int>(m - static_cast<long long>(fr) * op.hw);
    const int ho = rem / op.cv_w, wo = rem - ho * op.cv_w;
    const int hs = ho + tp.dy, ws = wo + tp.dx;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = 0.f;
    if (static_cast<unsigned>(hs) < static_cast<unsigned>(op.cv_h) && static_cast<unsigned>(ws) < static_cast<unsigned>(op.cv_w)) {
      load_vec<T, NV>(in1 + conv3_src(op, fr, hs, ws) + tp.c, v);
      if (op.scale) {
        float s[NV], b[NV];
        load_vec<float, NV>(op.scale + tp.c, s);
        load_vec<float, NV>(op.shift + tp.c, b);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          float z = fmaf(v[i], s[i], b[i]);
          if (op.relu6) z = fminf(fmaxf(z, 0.f), relu_hi(op.relu6));
          v[i] = z;
        }
      }
    }
  } else if (kGate && op.mode == EHGR_ROW_GATE) {
    load_vec<T, NV>(in1 + off, v);
    const long long f = m / op.hw;
    const float g1 = static_cast<const float*>(op.in2)[m];
    float g2[NV], g3[NV];
    load_vec<float, NV>(op.scale + f * C + c0, g2);
    load_vec<float, NV>(op.shift + f * C + c0, g3);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] *= (3.f + g1) + (g2[i] + g3[i]);
  } else {  // EHGR_ROW_BNBWD
    float g[NV], r[NV], ca[NV], cb[NV], cc[NV];
    load_vec<T, NV>(in1 + off, g);
    load_vec<T, NV>(static_cast<const T*>(op.in2) + off, r);
    load_vec<float, NV>(op.ca + c0, ca);
    load_vec<float, NV>(op.cb + c0, cb);
    load_vec<float, NV>(op.cc + c0, cc);
    if (op.relu6) {
      float s[NV], b[NV];
      load_vec<float, NV>(op.scale + c0, s);
      load_vec<float, NV>(op.shift + c0, b);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float z = fmaf(r[i], s[i], b[i]);
        if (!(z > 0.f && z < relu_hi(op.relu6))) g[i] = 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = fmaf(ca[i], g[i], fmaf(cb[i], r[i], cc[i]));
  }
}

// SHIFT vector that straddles a fold boundary, assembled and re-packed to its 16 raw bytes.  Kept out
// of line: it is the rare path and would otherwise be inlined into every unrolled fetch.
template <typename T, int NV>
__device__ __noinline__ uint4 shift_straddle_raw(const RowOp& op, long long m, int c0, int C) {
  float v[NV];
  load_row<T, NV, false>(op, m, c0, C, v);
  uint4 out;
  store_vec<T, NV>(reinterpret_cast<T*>(&out), v);
  return out;
}

// Same operand, for threads that read MANY rows of the SAME channel vector (depthwise taps, reductions,
// GEMM operand staging): the per-channel coefficients are fetched once into registers by init() and
// load() only touches the activation tensors.
// kTwo = false compiles the two-tensor BNBWD mode out (fewer registers) for kernels whose operand is
// known to be PLAIN / AFFINE / SHIFT; the host checks the mode before choosing such an instantiation.
// kGate = true compiles the GATE mode in (pointwise GEMM operands and row_apply only).
template <typename T, int NV, bool kTwo = true, bool kGate = false>
struct RowLoader {
  float s[NV], b[NV], ca[kTwo ? NV : 1], cb[kTwo ? NV : 1], cc[kTwo ? NV : 1];
  int c0, C;

  __device__ __forceinline__ void init(const RowOp& op, int c0_, int C_) {
    c0 = c0_;
    C = C_;
    if (op.mode == EHGR_ROW_AFFINE || (op.mode == EHGR_ROW_BNBWD && op.relu6)) {
      load_vec<float, NV>(op.scale + c0, s);
      load_vec<float, NV>(op.shift + c0, b);
    }
    if constexpr (kTwo) {
      if (op.mode == EHGR_ROW_BNBWD) {
        load_vec<float, NV>(op.ca + c0, ca);
        load_vec<float, NV>(op.cb + c0, cb);
        load_vec<float, NV>(op.cc + c0, cc);
      }
    }
  }

  __device__ __forceinline__ void load(const RowOp& op, long long m, float (&v)[NV]) const {
    const T* in1 = static_cast<const T*>(op.in1);
    const long long off = m * C + c0;
    if (op.mode == EHGR_ROW_PLAIN) {
      load_vec<T, NV>(in1 + off, v);
    } else if (op.mode == EHGR_ROW_AFFINE) {
      load_vec<T, NV>(in1 + off, v);
      if (op.relu6) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = fminf(fmaxf(fmaf(v[i], s[i], b[i]), 0.f), relu_hi(op.relu6));
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = fmaf(v[i], s[i], b[i]);
      }
    } else if (kTwo && op.mode == EHGR_ROW_BNBWD) {
      float g[NV], r[NV];
      load_vec<T, NV>(in1 + off, g);
      load_vec<T, NV>(static_cast<const T*>(op.in2) + off, r);
      if (op.relu6) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float z = fmaf(r[i], s[i], b[i]);
          if (!(z > 0.f && z < relu_hi(op.relu6))) g[i] = 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = fmaf(ca[kTwo ? i : 0], g[i], fmaf(cb[kTwo ? i : 0], r[i], cc[kTwo ? i : 0]));
    } else {
      load_row<T, NV, kGate>(op, m, c0, C, v);  // SHIFT / GATE: no per-channel coefficients held in registers
    }
  }

  // Split form for memory-level parallelism: fetch() only ISSUES the 16-byte loads of a row (no
  // instruction depends on the data), finish() does the arithmetic.  Kernels fetch a batch of rows
  // (all nine depthwise taps, several GEMM operand vectors) before finishing the first one, so a thread
  // has many loads in flight instead of one round trip per row.  Requires NV == VecOf<T>::N.
  struct Raw { uint4 a; uint4 b[1]; };   // b: second tensor (BNBWD) or the row index (GATE); dead otherwise

  __device__ __forceinline__ Raw fetch(const RowOp& op, long long m) const {
    static_assert(NV == VecOf<T>::N, "fetch/finish work on full 16-byte vectors");
    Raw r;
    r.a = make_uint4(0, 0, 0, 0);
    r.b[0] = r.a;
    const long long off = m * C + c0;
    const T* in1 = static_cast<const T*>(op.in1);
    if (op.mode == EHGR_ROW_SHIFT) {
      const long long frame = m / op.hw;
      const int t = static_cast<int>(frame % op.n_segment);
      const int dir = op.shift_dir < 0 ? -1 : 1;
      const long long step = static_cast<long long>(dir) * op.hw * C;
      const int fold = op.fold;
      auto cls_of = [fold](int c) { return c < fold ? 0 : (c < 2 * fold ? 1 : 2); };
      const int cl = cls_of(c0), ch = cls_of(c0 + NV - 1);
      if (cl == ch) {
        const bool has_next = dir > 0 ? (t < op.n_segment - 1) : (t > 0);
        const bool has_prev = dir > 0 ? (t > 0) : (t < op.n_segment - 1);
        const bool ok = cl == 2 || (cl == 0 ? has_next : has_prev);
        if (ok) r.a = *reinterpret_cast<const uint4*>(in1 + off + (cl == 0 ? step : cl == 1 ? -step : 0));
      } else {  // straddles a fold boundary (C = 24, 32, 96, 160: one vector per row): assemble now
        r.a = shift_straddle_raw<T, NV>(op, m, c0, C);
      }
    } else {
      r.a = *reinterpret_cast<const uint4*>(in1 + off);
      if constexpr (kGate) {
        if (op.mode == EHGR_ROW_GATE) {   // finish() needs the row index for the per-row / per-frame gates
          r.b[0].x = static_cast<uint32_t>(m & 0xffffffffLL);
          r.b[0].y = static_cast<uint32_t>(m >> 32);
        }
      }
      if constexpr (kTwo) {
        if (op.mode == EHGR_ROW_BNBWD) r.b[0] = *reinterpret_cast<const uint4*>(static_cast<const T*>(op.in2) + off);
      }
    }
    return r;
  }

  // finish() for bf16 storage with the result re-packed to its 16 bytes (shared-memory staging): packed
  // FFMA2 arithmetic, ReLU folded into the conversion, min(.,6) on the packed halves.  Bit-identical to
  // pack(finish()): rounding is monotonic and 6 is representable.
  __device__ __forceinline__ uint4 finish_packed(const RowOp& op, const Raw& r) const {
    static_assert(sizeof(T) == 2 && NV == 8, "finish_packed: bf16 storage only");
    if (op.mode == EHGR_ROW_PLAIN || op.mode == EHGR_ROW_SHIFT) return r.a;
    const uint32_t w[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
    uint32_t o[4];
    if (op.mode == EHGR_ROW_AFFINE) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 z = __ffma2_rn(bf2_to_f2(w[i]), make_float2(s[2 * i], s[2 * i + 1]), make_float2(b[2 * i], b[2 * i + 1]));
        o[i] = op.relu6 ? bf2_min_hi(f2_to_bf2_relu(z), op.relu6) : pack_bf16x2(z.x, z.y);
      }
      return make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (kTwo && op.mode == EHGR_ROW_BNBWD) {
      const uint32_t rw[4] = {r.b[0].x, r.b[0].y, r.b[0].z, r.b[0].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 g = bf2_to_f2(w[i]);
        const float2 raw = bf2_to_f2(rw[i]);
        if (op.relu6) {
          const float2 z = __ffma2_rn(raw, make_float2(s[2 * i], s[2 * i + 1]), make_float2(b[2 * i], b[2 * i + 1]));
          const float hi = relu_hi(op.relu6);
          if (!(z.x > 0.f && z.x < hi)) g.x = 0.f;
          if (!(z.y > 0.f && z.y < hi)) g.y = 0.f;
        }
        const int j = kTwo ? 2 * i : 0;
        const float2 t = __ffma2_rn(make_float2(cb[j], cb[kTwo ? j + 1 : 0]), raw, make_float2(cc[j], cc[kTwo ? j + 1 : 0]));
        const float2 v = __ffma2_rn(make_float2(ca[j], ca[kTwo ? j + 1 : 0]), g, t);
        o[i] = pack_bf16x2(v.x, v.y);
      }
      return make_uint4(o[0], o[1], o[2], o[3]);
    }
    float v[NV];
    finish(op, r, v);
    uint4 out;
    store_vec<T, NV>(reinterpret_cast<T*>(&out), v);
    return out;
  }

  __device__ __forceinline__ void finish(const RowOp& op, const Raw& r, float (&v)[NV]) const {
    load_vec<T, NV>(reinterpret_cast<const T*>(&r.a), v);
    if (op.mode == EHGR_ROW_AFFINE) {
      if (op.relu6) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = fminf(fmaxf(fmaf(v[i], s[i], b[i]), 0.f), relu_hi(op.relu6));
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = fmaf(v[i], s[i], b[i]);
      }
    } else if (kGate && op.mode == EHGR_ROW_GATE) {
      const long long m = static_cast<long long>(r.b[0].x) | (static_cast<long long>(r.b[0].y) << 32);
      const long long f = m / op.hw;
      const float g1 = static_cast<const float*>(op.in2)[m];
      float g2[NV], g3[NV];
      load_vec<float, NV>(op.scale + f * C + c0, g2);
      load_vec<float, NV>(op.shift + f * C + c0, g3);
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] *= (3.f + g1) + (g2[i] + g3[i]);
    } else if (kTwo && op.mode == EHGR_ROW_BNBWD) {
      float raw[NV];
      load_vec<T, NV>(reinterpret_cast<const T*>(&r.b[0]), raw);
      if (op.relu6) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float z = fmaf(raw[i], s[i], b[i]);
          if (!(z > 0.f && z < relu_hi(op.relu6))) v[i] = 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = fmaf(ca[kTwo ? i : 0], v[i], fmaf(cb[kTwo ? i : 0], raw[i], cc[kTwo ? i : 0]));
    }
  }
};

inline int validate_rowop(const RowOp* op, int es, bool allow_conv3 = false) {
  if (!op || !op->in1) return EHGR_E_NULL;
  if (!aligned_to(op->in1, 16)) return EHGR_E_ALIGN;
  switch (op->mode) {
    case EHGR_ROW_CONV3:
      if (!allow_conv3) return EHGR_E_UNSUPPORTED;            // GEMM family only
      if (op->cv_h <= 0 || op->cv_w <= 0 || op->cv_cin <= 0 || (op->cv_cin % 8) || op->hw != op->cv_h * op->cv_w) return EHGR_E_SHAPE;
      if (op->cv_up != 0 && (op->cv_up != 1 || (op->cv_h & 1) || (op->cv_w & 1))) return EHGR_E_SHAPE;
      if (op->scale && !op->shift) return EHGR_E_NULL;
      return EHGR_OK;
    case EHGR_ROW_PLAIN: return EHGR_OK;
    case EHGR_ROW_AFFINE: return (op->scale && op->shift) ? EHGR_OK : EHGR_E_NULL;
    case EHGR_ROW_SHIFT:
      return (op->n_segment > 0 && op->hw > 0 && op->fold >= 0) ? EHGR_OK : EHGR_E_SHAPE;
    case EHGR_ROW_GATE:
      return (op->in2 && op->scale && op->shift && op->hw > 0) ? EHGR_OK : EHGR_E_NULL;
    case EHGR_ROW_BNBWD:
      if (!op->in2 || !op->ca || !op->cb || !op->cc) return EHGR_E_NULL;
      if (op->relu6 && (!op->scale || !op->shift)) return EHGR_E_NULL;
      return aligned_to(op->in2, 16) ? EHGR_OK : EHGR_E_ALIGN;
    default: return EHGR_E_DTYPE;
  }
  (void)es;
}

// for kernels compiled without the GATE mode (depthwise family)
inline int validate_rowop_nogate(const RowOp* op, int es) {
  if (op && op->mode == EHGR_ROW_GATE) return EHGR_E_UNSUPPORTED;
  return validate_rowop(op, es);
}

}  // namespace ehgr
