// pw_tc.cu — K8 on the 5th-generation tensor cores: pointwise-conv GEMM (forward and dgrad form) and
// the weight-gradient GEMM as hand-written tcgen05 kernels for sm_100a, bf16 operands, fp32
// accumulation in TMEM.
//
//   out[M,N] = rowop(A)[M,K] * B[K,N] (+ addend)          (ehgr_pw_gemm, bf16 storage)
//
// Persistent, warp-specialised CTA (one per SM, 288 threads):
//   warps 0-3  PRODUCERS : read the A row operand (BatchNorm+ReLU6 / temporal shift / BN-backward applied
//                          on the fly, rowop.cuh) and the fp32 weights, convert to bf16 and store them
//                          into the shared-memory ring in the UMMA canonical no-swizzle layout
//                          (8x16-byte core matrices).  A register path instead of TMA because the
//                          operand is TRANSFORMED on load — that is what removes the separate BN /
//                          ReLU6 / shift passes — and because K is as small as 16 (32-byte rows).
//   warp  4    MMA       : one elected thread issues tcgen05.mma (M=128, N=BN<=256, K=16 per
//                          instruction) into one of two TMEM accumulator buffers, tcgen05.commit
//                          releases ring slots / publishes the accumulator through mbarriers.
//   warps 5-8  EPILOGUE  : tcgen05.ld the accumulator (each warp its 32-lane quarter), fold the
//                          BatchNorm batch statistics (shuffle transpose-reduction, kept in registers
//                          across all tiles of the CTA, one flush of double atomics at the end), add the
//                          optional addend, convert to bf16, store 32-byte row segments.
// The two TMEM buffers let the epilogue of tile i overlap the loads and MMAs of tile i+1.
// w_is_kn selects the B operand's major-ness: forward reads the conv weight [N,K] as a K-major B,
// dgrad reads the same array [K,N] as an MN-major B — no transposed weight copy exists.
//
//   dw[N,K] += rowop(dy)^T[N,M] * rowop(a)[M,K]            (ehgr_pw_wgrad)
// uses the same staging with BOTH operands MN-major (the reduction runs over rows): see pw_wgrad_tc.
#include "rowop.cuh"

namespace ehgr {
namespace tc {

constexpr int BM = 128;            // rows per tile (UMMA M)
constexpr int BK = 64;             // reduction elements per ring stage
constexpr int kStages = 4;
constexpr int kProducerGroups = 2;  // x 4 warps each; groups take ring stages round-robin
constexpr int kProducerWarps = 4 * kProducerGroups;
constexpr int kThreads = (kProducerWarps + 1 + 4) * 32;   // producers + 1 MMA warp + 4 epilogue warps
constexpr int kABytes = BM * BK * 2;                 // 16 KB
constexpr uint32_t kSpinLimit = 1u << 28;            // bounded waits: trap instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor): start address, leading
// byte offset (stride between core matrices along K), stride byte offset (stride between core
// matrices along M/N), all in 16-byte units; version 1 at bit 46.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, major-ness, N>>3, M>>4.
__device__ __forceinline__ uint32_t make_idesc(int m, int n, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                                  // c_format = F32
  d |= 1u << 7;                                  // a_format = BF16
  d |= 1u << 10;                                 // b_format = BF16
  d |= static_cast<uint32_t>(a_mn_major) << 15;
  d |= static_cast<uint32_t>(b_mn_major) << 16;
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(m >> 4) << 24;
  return d;
}

__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

struct GemmArgs {
  RowOp a;
  const float* w;
  int w_is_kn;
  __nv_bfloat16* out;
  const __nv_bfloat16* addend;
  double* stats;
  long long M;
  int K, N;        // reduction length, output columns
  int BN;          // output columns per tile (multiple of 16, <= 256)
  int n_chunks;    // ceil(N / BN)
  int m_tiles;
  int tmem_cols;   // power of two >= 2*BN
};

// transpose-reduce 16 columns held by the 32 lanes of a warp: returns, in every lane, the sum over
// the 32 lanes of column ((lane >> 1) & 15).
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
  float a[8];
  const bool u4 = lane & 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float keep = u4 ? v[8 + i] : v[i], send = u4 ? v[i] : v[8 + i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float b[4];
  const bool u3 = lane & 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = u3 ? a[4 + i] : a[i], send = u3 ? a[i] : a[4 + i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float c[2];
  const bool u2 = lane & 4;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = u2 ? b[2 + i] : b[i], send = u2 ? b[i] : b[2 + i];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const bool u1 = lane & 2;
  const float keep = u1 ? c[1] : c[0], send = u1 ? c[0] : c[1];
  float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}

__global__ void __launch_bounds__(kThreads, 1) pw_gemm_tc_kernel(GemmArgs p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int b_bytes = p.BN * BK * 2;
  const int stage_bytes = kABytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
  // bars: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kStages);
  const uint32_t bar_tfull = smem_u32(bars + 2 * kStages), bar_tempty = smem_u32(bars + 2 * kStages + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  const uint32_t smem_base = smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 128);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, 128);
    }
    fence_mbar_init();
  }
  if (warp == kProducerWarps) tmem_alloc(smem_u32(tmem_slot), static_cast<uint32_t>(p.tmem_cols));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int chunk = blockIdx.x % p.n_chunks;
  const int n0 = chunk * p.BN;
  const int total_tiles = p.m_tiles * p.n_chunks;
  const int k_stages = (p.K + BK - 1) / BK;
  const int Kp = (p.K + 15) & ~15;

  if (warp < kProducerWarps) {
    // ===================== PRODUCERS =====================
    // kProducerGroups groups of 128 threads take ring stages round-robin: each group has its own
    // registers, so that many stages' global loads are in flight per SM.  Within a stage a thread
    // first FETCHES all its vectors (loads only), then converts and stores them.
    const int tid = threadIdx.x & 127;
    const int group = warp >> 2;
    const int r = tid & 7, t8 = tid >> 3;          // row inside an 8-row core matrix, 16 vector slots
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const long long m0 = static_cast<long long>(tile / p.n_chunks) * BM;
      for (int ks = 0; ks < k_stages; ++ks, ++it) {
        if (static_cast<int>(it % kProducerGroups) != group) continue;
        const int s = it % kStages;
        const uint32_t round = it / kStages;
        uint8_t* a_dst = smem + s * stage_bytes;
        uint8_t* b_dst = a_dst + kABytes;
        const int k_base = ks * BK;
        const int kvalid = min(BK, Kp - k_base);     // multiple of 16
        const int kv = kvalid >> 3;                    // 16-byte vectors per row in this stage: 2, 4, 6 or 8
        bool waited = false;
        // ---- A: [128 rows][kvalid] -> K-major core matrices: (row/8)*1024 + k8*128 + (row%8)*16
        if ((16 % kv) == 0) {
          // thread owns channel vector k8 and kv row groups (rg0, rg0 + 16/kv, ...)
          const int k8 = t8 % kv, rg0 = t8 / kv, rg_step = 16 / kv;
          const int k = k_base + k8 * 8;
          RowLoader<__nv_bfloat16, 8> ld;
          if (k < p.K) ld.init(p.a, k, p.K);
#pragma unroll 1
          for (int j0 = 0; j0 < kv; j0 += 4) {
            RowLoader<__nv_bfloat16, 8>::Raw raw[4];
            bool live[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const long long m = m0 + (rg0 + (j0 + j) * rg_step) * 8 + r;
              live[j] = (j0 + j) < kv && m < p.M && k < p.K;
              if (live[j]) raw[j] = ld.fetch(p.a, m);
            }
            if (!waited) { mbar_wait(bar_empty + 8 * s, (round & 1) ^ 1); waited = true; }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if ((j0 + j) < kv) {
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = 0.f;
                if (live[j]) ld.finish(p.a, raw[j], f);
                *reinterpret_cast<uint4*>(a_dst + (rg0 + (j0 + j) * rg_step) * 1024 + k8 * 128 + r * 16) = pack8(f);
              }
            }
          }
        } else {
          mbar_wait(bar_empty + 8 * s, (round & 1) ^ 1);
          waited = true;
          for (int v = tid; v < BM * kv; v += 128) {
            const int k8 = (v >> 3) % kv, rg = (v >> 3) / kv;
            const int k = k_base + k8 * 8;
            const long long m = m0 + rg * 8 + r;
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = 0.f;
            if (m < p.M && k < p.K) load_row<__nv_bfloat16, 8>(p.a, m, k, p.K, f);
            *reinterpret_cast<uint4*>(a_dst + rg * 1024 + k8 * 128 + r * 16) = pack8(f);
          }
        }
        // ---- B (fp32 weights, L2 resident): batches of 4 vectors = 8 x 16-byte loads in flight
        const int nvec = (p.BN >> 3) * kvalid;       // (BN/8 groups) x (kv vectors) x 8
#pragma unroll 1
        for (int v0 = tid; v0 < nvec; v0 += 4 * 128) {
          float4 lo[4], hi[4];
          int dst[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int v = v0 + j * 128;
            lo[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            hi[j] = lo[j];
            dst[j] = -1;
            if (v < nvec) {
              const int x = v & 7, kg = (v >> 3) % kv, ng = (v >> 3) / kv;
              dst[j] = ng * 1024 + kg * 128 + x * 16;
              const float* src = nullptr;
              if (!p.w_is_kn) {   // B[k][n] = w[n*K + k]: K-major, x = n % 8, vector = 8 consecutive k
                const int n = n0 + ng * 8 + x, k = k_base + kg * 8;
                if (n < p.N && k < p.K) src = p.w + static_cast<size_t>(n) * p.K + k;
              } else {            // B[k][n] = w[k*N + n]: MN-major, x = k % 8, vector = 8 consecutive n
                const int k = k_base + kg * 8 + x, n = n0 + ng * 8;
                if (k < p.K && n < p.N) src = p.w + static_cast<size_t>(k) * p.N + n;
              }
              if (src) {
                lo[j] = __ldg(reinterpret_cast<const float4*>(src));
                hi[j] = __ldg(reinterpret_cast<const float4*>(src) + 1);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (dst[j] >= 0) {
              const float f[8] = {lo[j].x, lo[j].y, lo[j].z, lo[j].w, hi[j].x, hi[j].y, hi[j].z, hi[j].w};
              *reinterpret_cast<uint4*>(b_dst + dst[j]) = pack8(f);
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(bar_full + 8 * s);
      }
    }
  } else if (warp == kProducerWarps) {
    // ===================== MMA ISSUER =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, p.BN, 0, p.w_is_kn ? 1 : 0);
      uint32_t it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
        const uint32_t buf = tl & 1;
        mbar_wait(bar_tempty + 8 * buf, ((tl >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * static_cast<uint32_t>(p.BN);
        for (int ks = 0; ks < k_stages; ++ks, ++it) {
          const int s = it % kStages;
          mbar_wait(bar_full + 8 * s, (it / kStages) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_base + s * stage_bytes;
          const uint32_t b_addr = a_addr + kABytes;
          const int ksteps = min(BK, Kp - ks * BK) >> 4;
          for (int kk = 0; kk < ksteps; ++kk) {
            // one K=16 step = two 8-element core matrices along K = 256 bytes in both layouts
            const uint64_t da = make_desc(a_addr + kk * 256, 128, 1024);
            const uint64_t db = make_desc(b_addr + kk * 256, 128, 1024);
            umma_bf16(d_tmem, da, db, idesc, (ks | kk) ? 1u : 0u);
          }
          umma_commit(bar_empty + 8 * s);          // ring slot free once these MMAs have read it
        }
        umma_commit(bar_tfull + 8 * buf);          // accumulator complete
      }
    }
  } else {
    // ===================== EPILOGUE =====================
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int n_cc = p.BN >> 4;                     // 16-column chunks
    float ssum[16], ssq[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) ssum[i] = ssq[i] = 0.f;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
      const uint32_t buf = tl & 1;
      const long long m = static_cast<long long>(tile / p.n_chunks) * BM + q * 32 + lane;
      mbar_wait(bar_tfull + 8 * buf, (tl >> 1) & 1);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * static_cast<uint32_t>(p.BN);
#pragma unroll
      for (int cc = 0; cc < 16; ++cc) {
        if (cc < n_cc) {
          float v[16];
          tmem_ld16(t_base + cc * 16, v);
          if (p.stats) {
            float sq[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) sq[i] = v[i] * v[i];
            ssum[cc] += warp_colsum16(v, lane);
            ssq[cc] += warp_colsum16(sq, lane);
          }
          const int n = n0 + cc * 16;
          if (m < p.M && n < p.N) {
            const long long off = m * p.N + n;
            const bool second = n + 8 < p.N;       // N % 8 == 0: a chunk is 1 or 2 valid 8-column halves
            if (p.addend) {
              float ad[8];
              load_vec<__nv_bfloat16, 8>(p.addend + off, ad);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += ad[i];
              if (second) {
                load_vec<__nv_bfloat16, 8>(p.addend + off + 8, ad);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[8 + i] += ad[i];
              }
            }
            float lo[8], hi[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { lo[i] = v[i]; hi[i] = v[8 + i]; }
            *reinterpret_cast<uint4*>(p.out + off) = pack8(lo);
            if (second) *reinterpret_cast<uint4*>(p.out + off + 8) = pack8(hi);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * buf);
    }
    if (p.stats && !(lane & 1)) {
      const int col = (lane >> 1) & 15;
#pragma unroll
      for (int cc = 0; cc < 16; ++cc) {
        const int n = n0 + cc * 16 + col;
        if (cc < n_cc && n < p.N) {
          atomicAdd(&p.stats[n], static_cast<double>(ssum[cc]));
          atomicAdd(&p.stats[p.N + n], static_cast<double>(ssq[cc]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == kProducerWarps) {
    __syncwarp();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

static int pick_bn(int N) {
  // output columns per tile: a multiple of 16 (UMMA N for M=128), <= 256, as few chunks as possible
  const int Np = (N + 15) & ~15;
  const int chunks = (Np + 255) / 256;
  int bn = (Np / chunks + 15) & ~15;
  while (bn * chunks < Np) bn += 16;
  return bn;
}

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

bool pw_gemm_tc_supported(const RowOp& a, int w_is_kn, long long M, int K, int N, int dtype) {
  (void)a; (void)w_is_kn;
  if (dtype != EHGR_BF16) return false;
  if (K % 8 || N % 8 || K < 8 || N < 8) return false;
  if (M < 1 || M / tc::BM > 0x3fffffff) return false;
  return true;
}

int pw_gemm_tc(const RowOp& a, const float* w, int w_is_kn, void* out, const void* addend, double* stats,
               long long M, int K, int N, cudaStream_t s) {
  tc::GemmArgs p;
  p.a = a; p.w = w; p.w_is_kn = w_is_kn;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.addend = static_cast<const __nv_bfloat16*>(addend);
  p.stats = stats;
  p.M = M; p.K = K; p.N = N;
  p.BN = tc::pick_bn(N);
  p.n_chunks = (N + p.BN - 1) / p.BN;
  p.m_tiles = static_cast<int>(cdiv(M, tc::BM));
  int cols = 32;
  while (cols < 2 * p.BN) cols <<= 1;
  p.tmem_cols = cols;
  const long long tiles = static_cast<long long>(p.m_tiles) * p.n_chunks;
  // persistent grid: a multiple of n_chunks (a CTA keeps one column chunk -> statistics stay in registers)
  long long grid = std::min<long long>(tiles, kNumSMs);
  grid = std::max<long long>(p.n_chunks, grid / p.n_chunks * p.n_chunks);
  const size_t smem = static_cast<size_t>(tc::kStages) * (tc::kABytes + p.BN * tc::BK * 2) + 128;
  cudaFuncSetAttribute(tc::pw_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  tc::pw_gemm_tc_kernel<<<static_cast<unsigned>(grid), tc::kThreads, smem, s>>>(p);
  return launch_status();
}

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
