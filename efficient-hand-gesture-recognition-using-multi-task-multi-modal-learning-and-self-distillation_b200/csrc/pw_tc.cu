// pw_tc.cu — K8 on the 5th-generation tensor cores: the pointwise-conv GEMM (forward and dgrad form)
// as a hand-written tcgen05 kernel for sm_100a, bf16 operands, fp32 accumulation in TMEM.
//
//   out[M,N] = rowop(A)[M,K] * B[K,N] (+ addend)          (ehgr_pw_gemm, bf16 storage)
//
// These GEMMs are HBM-bound (K, N <= 1280, M up to 3.2 M rows; 20-140 FLOP/B against a ridge of ~216):
// the design goal is bytes in flight, not MMA issue rate.
//
// Persistent, warp-specialised CTA, one per SM, 13 warps:
//   warps 0-7  PRODUCERS : each warp owns whole ring stages (stage i -> warp i % 8), so eight stages'
//                          global loads are in flight per SM.  A warp FETCHES a batch of eight 16-byte
//                          vectors per lane (loads only), then applies the row operand (BatchNorm+ReLU6
//                          / temporal shift / BN-backward, rowop.cuh), converts to bf16 and stores into
//                          the UMMA canonical no-swizzle layout (8-row x 16-byte core matrices).  A
//                          register path instead of TMA because the operand is TRANSFORMED on load —
//                          that is what removes the reference's separate BN / ReLU6 / shift passes —
//                          and K is as small as 16 (32-byte rows).
//   B operand            : the fp32 master weights, converted to bf16.  When BN*K*2 bytes fit they are
//                          staged ONCE per CTA and stay resident; otherwise (K >= 320: the small-M
//                          layers) a 64-wide slice travels with every A stage.
//   warp  8    MMA       : one elected thread issues tcgen05.mma (M=128, N=BN<=256, K=16 per
//                          instruction) into one of two TMEM accumulator buffers; tcgen05.commit
//                          releases ring slots / publishes the accumulator through mbarriers.
//   16 EPILOGUE warps    : four per TMEM lane quarter, each a contiguous quarter of the tile's columns (<= 4 chunks of
//                          16).  The epilogue is a chain of dependent instructions per warp (ncu: the tile rate
//                          of the output-heavy expand layers followed the per-warp instruction count, not the
//                          bytes), so it is spread over many warps and kept short: two tcgen05.ld in flight,
//                          optional addend (prefetched into the staging rows with cp.async while the warp waits
//                          for the accumulator), bf16 rows staged in shared memory (odd 16-byte pitch: no bank
//                          conflicts), the TMEM buffer released as soon as the last chunk has been read, then
//                          COALESCED 16-byte global stores: a lane owns one 16-byte piece column of the row
//                          segment and walks down the rows — and accumulates the BatchNorm batch statistics of
//                          its eight channels from the very words it stores (packed FADD2 / FFMA2, registers kept
//                          across all tiles of the CTA, one reduction at the end).  The statistics therefore
//                          describe the STORED bf16 tensor, which is the tensor the consumer normalises.
// The two TMEM buffers let the epilogue of tile i overlap the loads and MMAs of tile i+1.
// w_is_kn selects the B operand's major-ness: forward reads the conv weight [N,K] as a K-major B,
// dgrad reads the same array [K,N] as an MN-major B — no transposed weight copy exists.
#include <cstring>
#include <type_traits>

#include "tma.cuh"
#include "bnfin.cuh"

namespace ehgr {
namespace tc {

constexpr int BM = 128;            // rows per tile (UMMA M)
constexpr int BK = 64;             // reduction elements per ring stage
constexpr int kProducerWarps = 7;   // 12 warps = 384 threads: up to 168 registers per thread
constexpr int kMmaWarp = kProducerWarps;
// Epilogue warps: a template parameter (kEpi).  16 (four per TMEM lane quarter; 768 threads, <= 80 registers) for the
// output-heavy shapes (N > K: expand forward, dgrad of the project layers), whose tile rate follows the epilogue's
// per-warp instruction chain; 8 (two per quarter; 512 threads, <= 128 registers) for the input-heavy shapes
// (K >= N), whose bound is the producers' load + in-place transform and which want the registers and issue slots.
constexpr int threads_of(int epi_warps) { return (kProducerWarps + 1 + epi_warps) * 32; }
constexpr int kMaxStages = 12;
constexpr int kBatch = 4;                      // vectors a producer lane fetches before it converts any
constexpr int kMaxBufs = 8;                    // TMEM accumulator buffers (n_bufs * BN <= 512 columns)
constexpr int kBarBytes = 512;                 // mbarriers + TMEM slot
constexpr int kTailBytes = kBarBytes;

struct GemmArgs {
  RowOp a;
  const float* w;
  const __nv_bfloat16* w16;   // optional bf16 mirror of w (same layout): B staged with cp.async
  int w_is_kn;
  __nv_bfloat16* out;
  const __nv_bfloat16* addend;
  double* stats;
  long long M;
  int K, N;        // reduction length, output columns
  int BN;          // output columns per tile (multiple of 16, <= 256)
  int n_chunks;    // ceil(N / BN)
  int m_tiles;
  int tmem_cols;   // power of two >= n_bufs*BN
  int n_bufs;      // accumulator buffers in TMEM: the tile hand-shake latency is spread over n_bufs tiles
  int b_resident;  // 1: whole [BN x Kp] B staged once; 0: a [BN x 64] slice per stage
  int bm_rows;     // rows of a tile that exist (128; CONV3 by TMA: the pixels of one box, 98 or 112) — tile t starts at row t*bm_rows
  int cv_hbox, cv_nbox, cv_tiles_per_frame;   // CONV3 by TMA: a tile = cv_nbox frames x cv_hbox rows x the full width
  int cv_sw64;     // CONV3 by TMA with 32-channel pixels: a stage = two taps, each a box of 64-byte rows (SWIZZLE_64B), 8 KB apart
  int b_tma;       // streamed slice arrives as TMA box(es) in the SWIZZLE_128B layout (else 16-byte cp.async, no swizzle)
  int b_bytes;     // bytes of the B part of a stage
  int n_stages;    // ring depth
  int a_bytes;     // bytes of the A part of a stage = 128 * min(Kp,64) * 2
  int stage_bytes; // a_bytes (+ BN*128 when B is streamed)
  int epi_pitch;   // bytes between the 32 staging rows of an epilogue warp (odd multiple of 16)
  BnFin fin;       // BatchNorm finalisation by the last CTA (fin.counter == NULL: none)
};

// B tile -> shared memory in core-matrix layout: group stride `gs` bytes, k-group stride 128.
//   forward (w_is_kn = 0): B[k][n] = w[n*K + k]; K-major: x = n % 8, a vector = 8 consecutive k
//   dgrad   (w_is_kn = 1): B[k][n] = w[k*N + n]; MN-major: x = k % 8, a vector = 8 consecutive n
__device__ __forceinline__ void stage_b(const GemmArgs& p, uint8_t* dst, int gs, int n0, int k_base, int kvalid,
                                        int tid, int nthreads) {
  const int kv = kvalid >> 3;
  const int nvec = (p.BN >> 3) * kvalid;
  if (p.w16) {   // bf16 mirror: every 16-byte chunk is one asynchronous copy (the caller waits for them)
    // Lanes run along the CONTIGUOUS direction of the weight array (k-chunks of one row n for the K-major
    // form, n-chunks of one row k for the MN-major form): coalesced 128-byte reads, no per-chunk division.
    const uint32_t dst32 = smem_u32(dst);
    // Source / destination / live counts advance by constants: a producer warp streams a slice with every stage, so
    // the per-copy instruction count is what bounds the stage rate of the weight-streaming (small-M, large-K) layers.
    if (!p.w_is_kn) {
      const int step = nthreads >> 3;                   // 4 (one warp) or a multiple of 8 (whole CTA)
      const int nl0 = tid >> 3;
      for (int kg = tid & 7; kg < kv; kg += 8) {        // kv > 8 only for the resident form (whole K at once)
        const int k = k_base + kg * 8;
        const bool k_ok = k < p.K;
        const __nv_bfloat16* src = p.w16 + static_cast<size_t>(n0 + nl0) * p.K + k;
        const size_t sstep = static_cast<size_t>(step) * p.K;
        int left = k_ok ? p.N - n0 - nl0 : 0;           // > 0: row n exists
        if (step >= 8) {
          uint32_t d = dst32 + (nl0 >> 3) * gs + kg * 128 + (nl0 & 7) * 16;
          const uint32_t dstep = static_cast<uint32_t>((step >> 3) * gs);
#pragma unroll 4
          for (int nl = nl0; nl < p.BN; nl += step, d += dstep, src += sstep, left -= step)
            cp_async16(d, left > 0 ? src : p.w16, left > 0 ? 16u : 0u);
        } else {                                        // step == 4: rows nl0 and nl0 + 4 of every 8-row group
          uint32_t d = dst32 + kg * 128 + nl0 * 16;
#pragma unroll 4
          for (int g = 0; g < (p.BN >> 3); ++g, d += gs, src += 2 * sstep, left -= 8) {
            cp_async16(d, left > 0 ? src : p.w16, left > 0 ? 16u : 0u);
            cp_async16(d + 64, left > 4 ? src + sstep : p.w16, left > 4 ? 16u : 0u);
          }
        }
      }
    } else {
      const int ng = tid & 31, step = nthreads >> 5;
      const int n = n0 + ng * 8;
      if (ng < (p.BN >> 3)) {
        const int kl0 = tid >> 5;
        // (kl >> 3) * 128 + (kl & 7) * 16 == kl * 16: the k rows of one n group are contiguous in the stage
        uint32_t d = dst32 + ng * gs + kl0 * 16;
        const __nv_bfloat16* src = p.w16 + static_cast<size_t>(k_base + kl0) * p.N + n;
        const size_t sstep = static_cast<size_t>(step) * p.N;
        int left = n < p.N ? p.K - k_base - kl0 : 0;
#pragma unroll 4
        for (int kl = kl0; kl < kvalid; kl += step, d += 16 * step, src += sstep, left -= step)
          cp_async16(d, left > 0 ? src : p.w16, left > 0 ? 16u : 0u);
      }
    }
    return;
  }
#pragma unroll 1
  for (int v0 = tid; v0 < nvec; v0 += 4 * nthreads) {
    float4 lo[4], hi[4];
    int off[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int v = v0 + j * nthreads;
      lo[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      hi[j] = lo[j];
      off[j] = -1;
      if (v < nvec) {
        const int x = v & 7, kg = (v >> 3) % kv, ng = (v >> 3) / kv;
        off[j] = ng * gs + kg * 128 + x * 16;
        const float* src = nullptr;
        if (!p.w_is_kn) {
          const int n = n0 + ng * 8 + x, k = k_base + kg * 8;
          if (n < p.N && k < p.K) src = p.w + static_cast<size_t>(n) * p.K + k;
        } else {
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
      if (off[j] >= 0) {
        const float f[8] = {lo[j].x, lo[j].y, lo[j].z, lo[j].w, hi[j].x, hi[j].y, hi[j].z, hi[j].w};
        *reinterpret_cast<uint4*>(dst + off[j]) = pack8(f);
      }
    }
  }
}

// kMode != BNBWD (kAsync): operand modes PLAIN / AFFINE / SHIFT / GATE.  A producer warp fills its ring stage with
//                  16-byte cp.async copies straight into the core-matrix layout (a whole 16 KB stage in
//                  flight per warp, no registers held), waits, and — for AFFINE / GATE — transforms the
//                  stage IN PLACE (each lane re-reads exactly the chunks it copied).  SHIFT is a pure gather:
//                  the copy's source is the neighbouring frame or a zero fill.
// kAsync = false: register path (batched fetch -> rowop -> store) for the two-tensor BNBWD operand.
// kTma (operand modes PLAIN / AFFINE / GATE): the A stage — 128 rows x 64 channels, 128-byte rows — arrives as ONE
//                  cp.async.bulk.tensor box written in the SWIZZLE_128B layout (zero fill past M and past K); AFFINE /
//                  GATE then transform the 16-byte chunks in place (chunk c of row r sits at r*128 + ((c ^ (r & 7)) << 4)).
//                  The producers issue one instruction per stage instead of 32 copies per lane, which is what bounded
//                  the stage rate of the register / cp.async forms (ncu: ~20 dependent instructions per 16-byte copy).
template <int kMode, int kEpiWarps, bool kTma>
__global__ void __launch_bounds__(threads_of(kEpiWarps), 1)
pw_gemm_tc_kernel(const __grid_constant__ GemmArgs p, const __grid_constant__ CUtensorMap tm_a,
                  const __grid_constant__ CUtensorMap tm_b) {
  constexpr int kThreads = threads_of(kEpiWarps);
  constexpr int kParts = kEpiWarps / 4;           // epilogue warps per TMEM lane quarter
  constexpr bool kAsync = kMode != EHGR_ROW_BNBWD;     // the operand mode is a compile-time constant: one kernel per mode
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // SWIZZLE_128B stages must start on 1024-byte boundaries: the dynamic window follows the static variables, so the
  // base is rounded up here (the launch asks for 1 KB more) and the resident weights are padded to 1 KB
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int Kp = (p.K + 15) & ~15;
  const int b_res_bytes = p.b_resident ? (kTma ? (p.BN * Kp * 2 + 1023) & ~1023 : p.BN * Kp * 2) : 0;
  uint8_t* ring = smem + b_res_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + p.n_stages * p.stage_bytes);
  // bars: full[kMaxStages], empty[kMaxStages], tmem_full[2], tmem_empty[2]
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kMaxStages);
  const uint32_t bar_tfull = smem_u32(bars + 2 * kMaxStages), bar_tempty = smem_u32(bars + 2 * kMaxStages + kMaxBufs);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 2 * kMaxBufs);
  [[maybe_unused]] const uint32_t bar_landed = smem_u32(bars + 2 * kMaxStages + 2 * kMaxBufs + 1);   // kTma: box has landed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x % p.n_chunks;
  const int n0 = chunk * p.BN;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);          // one arrival per producer warp (lane 0, after __syncwarp)
      mbar_init(bar_empty + 8 * s, 1);
      if (kTma) mbar_init(bar_landed + 8 * s, 1);
    }
    if (kTma) tma::prefetch_map(&tm_a);
    if (p.b_tma) tma::prefetch_map(&tm_b);
    for (int b = 0; b < p.n_bufs; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, kEpiWarps);   // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), static_cast<uint32_t>(p.tmem_cols));
  if (p.b_resident) {
    stage_b(p, smem, Kp * 16, n0, 0, Kp, threadIdx.x, kThreads);
    cp_async_wait_all();
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_chunks;
  const int k_stages = (p.K + BK - 1) / BK;
  const int a_sbo = p.a_bytes >> 4;                 // BYTES between 8-row groups of an A stage = min(Kp,64)*16

  if (warp < kProducerWarps) {
    // ===================== PRODUCERS (one warp per ring stage) =====================
    // Lane -> (row r of an 8-row group, channel-vector slot).  16-byte accesses are served per QUARTER warp (8 lanes),
    // and the copies bypass L1: a 32-byte sector is fetched once only if both of its halves are asked for by the same
    // quarter.  Lane bit 0 therefore selects the half (slot bit 0), bits 1-3 the row and bit 4 the second slot bit: a
    // quarter reads four whole sectors (4 rows x 32 bytes).  (With r = lane & 7 every sector travelled twice from L2 —
    // ncu: 29.6 sectors per LDGSTS instead of 16.)  The price is a 2-way bank conflict on the shared-memory side of the
    // copy (slots k and k+1 are 128 bytes apart).
    const int r = (lane >> 1) & 7, slot = (lane & 1) | ((lane >> 3) & 2);
    const int pw = p.n_stages < kProducerWarps ? p.n_stages : kProducerWarps;   // active producer warps
    // ring position kept with counters (no integer division in the per-stage path): s = it % n_stages,
    // ph = (it / n_stages) & 1, turn = it % pw
    int s = 0, turn = 0;
    uint32_t ph = 0;
    const int tile_step = gridDim.x / p.n_chunks;            // grid is a multiple of n_chunks
    int m_tile = blockIdx.x / p.n_chunks;
    // CONV3: per-warp row table (1 KB each) behind the epilogue staging rows
    [[maybe_unused]] const uint32_t conv_tab32 =
        smem_u32(reinterpret_cast<uint8_t*>(bars) + kBarBytes) + static_cast<uint32_t>(kEpiWarps * 32 * p.epi_pitch);
    [[maybe_unused]] int tab_tile = -1;
    // Streamed weight slice of a stage.  TMA form: lane 0 adds the slice's bytes to the stage's FULL barrier and issues
    // one box (K-major: [BN rows x 64 k]) or one box per 64-column block (MN-major: [64 k x 64 n], 8 KB apart); the
    // warp's arrival on that barrier comes later in program order.  Otherwise 16-byte cp.async copies (stage_b).
    auto stream_b = [&](uint8_t* b_dst, int s_, int k_base_, int kvalid_) {
      if (p.b_tma) {
        if (lane == 0) {
          const uint32_t bar = bar_full + 8 * s_, d32 = smem_u32(b_dst);
          tma::expect_tx_only(bar, static_cast<uint32_t>(p.b_bytes));
          if (!p.w_is_kn) {
            tma::load_2d(d32, &tm_b, bar, k_base_, n0);
          } else {
            for (int b = 0; b * 64 < p.BN; ++b) tma::load_2d(d32 + b * 8192, &tm_b, bar, n0 + b * 64, k_base_);
          }
        }
      } else {
        stage_b(p, b_dst, 1024, n0, k_base_, kvalid_, lane, 32);
      }
    };
    // SHIFT with an odd fold: one scratch word per tile row and producer warp (same place as the CONV3 tables)
    [[maybe_unused]] const uint32_t shift_scr32 = conv_tab32 + static_cast<uint32_t>(warp) * 512u;
    for (int tile = blockIdx.x; tile < total_tiles && warp < pw; tile += gridDim.x, m_tile += tile_step) {
      const long long m0 = static_cast<long long>(m_tile) * p.bm_rows;
      for (int ks = 0; ks < k_stages; ++ks, ++turn, ++s) {
        if (turn == pw) turn = 0;
        if (s == p.n_stages) { s = 0; ph ^= 1; }
        // Stage i belongs to warp i % pw.  The empty-slot wait only tracks phase PARITY, so a producer
        // must never get two ring rounds ahead of the MMA warp: either pw == n_stages (a slot is only
        // ever filled by one warp, whose stages are sequential) or pw < n_stages (a warp's previous
        // stage was at most pw < n_stages stages back, and it waited for that slot's previous round).
        if (turn != warp) continue;
        const uint32_t parity = ph ^ 1;
        uint8_t* a_dst = ring + s * p.stage_bytes;
        const int k_base = ks * BK;
        const int kvalid = min(BK, Kp - k_base);     // multiple of 16
        const int kv = kvalid >> 3;                    // 2, 4, 6 or 8 channel vectors per row
        const int kvp = kv < 4 ? kv : 4;               // channel vectors handled per pass by the 4 lane slots
        const int f = 4 / kvp;                         // spare slots interleave row groups (kv == 2)
        if constexpr (kTma) {
          mbar_wait(bar_empty + 8 * s, parity);
          const uint32_t a_dst32 = smem_u32(a_dst);
          constexpr uint32_t kBoxBytes = BM * BK * 2;
          if constexpr (kMode == EHGR_ROW_CONV3) {
            // im2col by TMA: the stage's 64 channels of ONE tap for the tile's pixels are a 4-D box of the NHWC tensor
            // (channels, full width, cv_hbox rows, cv_nbox frames) shifted by the tap; the zero padding of the
            // convolution is the hardware's out-of-bounds fill.  Box rows land in pixel order = the tile's row order.
            if (lane == 0) {
              const int cin = p.a.cv_cin;
              const int frame_tile = m_tile / p.cv_tiles_per_frame;
              const int h0 = (m_tile - frame_tile * p.cv_tiles_per_frame) * p.cv_hbox;
              if (!p.cv_sw64) {
                const int tap = k_base / cin, ty = tap / 3;
                tma::expect_tx_only(bar_full + 8 * s, static_cast<uint32_t>(p.bm_rows) * 128u);
                tma::load_4d(a_dst32, &tm_a, bar_full + 8 * s, k_base - tap * cin, tap - ty * 3 - 1, h0 + ty - 1,
                             frame_tile * p.cv_nbox);
              } else {                                 // cin == 32: the stage's 64 columns are two whole taps
                const int n_box = kvalid >> 5;
                tma::expect_tx_only(bar_full + 8 * s, static_cast<uint32_t>(n_box * p.bm_rows) * 64u);
                for (int b = 0; b < n_box; ++b) {
                  const int tap = (k_base >> 5) + b, ty = tap / 3;
                  tma::load_4d(a_dst32 + b * 8192, &tm_a, bar_full + 8 * s, 0, tap - ty * 3 - 1, h0 + ty - 1, frame_tile * p.cv_nbox);
                }
              }
            }
          } else if constexpr (kMode == EHGR_ROW_PLAIN) {
            // the box completes its bytes on the FULL barrier itself; the warp's arrival follows the weight slice
            if (lane == 0) {
              tma::expect_tx_only(bar_full + 8 * s, kBoxBytes);
              tma::load_2d(a_dst32, &tm_a, bar_full + 8 * s, k_base, static_cast<int>(m0));
            }
          } else {
            if (lane == 0) {
              tma::expect_tx(bar_landed + 8 * s, kBoxBytes);
              tma::load_2d(a_dst32, &tm_a, bar_landed + 8 * s, k_base, static_cast<int>(m0));
            }
          }
          if (!p.b_resident && p.w16) stream_b(a_dst + p.a_bytes, s, k_base, kvalid);
          if constexpr (kMode != EHGR_ROW_PLAIN && kMode != EHGR_ROW_CONV3) {
            mbar_wait(bar_landed + 8 * s, ph);
            // in-place row operand; lanes of a quarter warp take the 8 rows of a group at one chunk: the XOR swizzle
            // spreads them over all banks
            const int tr = lane & 7, ts = lane >> 3;
            using Ld = RowLoader<__nv_bfloat16, 8, false, kMode == EHGR_ROW_GATE>;
            RowOp ac = p.a;
            ac.mode = kMode;
#pragma unroll 1
            for (int pass = 0; pass * 4 < kv; ++pass) {
              const int k8 = pass * 4 + ts;
              const int k = k_base + k8 * 8;
              if (!(k8 < kv && k < p.K)) continue;
              Ld ld;
              ld.init(ac, k, p.K);
              const uint32_t base = a_dst32 + tr * 128 + ((k8 ^ tr) << 4);
#pragma unroll 4
              for (int rg = 0; rg < 16; ++rg) {
                const long long m = m0 + rg * 8 + tr;
                if (m < p.M) {
                  const uint32_t dst = base + rg * 1024;
                  typename Ld::Raw raw;
                  raw.a = lds128(dst);
                  if (kMode == EHGR_ROW_GATE) {
                    raw.b[0].x = static_cast<uint32_t>(m & 0xffffffffLL);
                    raw.b[0].y = static_cast<uint32_t>(m >> 32);
                  }
                  sts128(dst, ld.finish_packed(ac, raw));
                }
              }
            }
          }
          if (!p.b_resident && p.w16) cp_async_wait_all();
        } else if constexpr (kAsync) {
          mbar_wait(bar_empty + 8 * s, parity);
          constexpr int mode = kMode;
          const __nv_bfloat16* in1 = static_cast<const __nv_bfloat16*>(p.a.in1);
          const uint32_t a_dst32 = smem_u32(a_dst);
          if constexpr (mode == EHGR_ROW_CONV3) {
            // Implicit-GEMM gather of a dense 3x3 convolution: column k = (tap, channel), row = output pixel.  The rows of
            // a tile are decomposed ONCE per tile into a per-warp table (element offset of the pixel in the stored tensor,
            // ho | wo << 16; rows past M get ho = 0x7fff and are never live); a stage then costs a table read, two
            // bounds checks and an add per 16-byte copy.  K is a multiple of 32, so a lane always owns the 16 rows
            // r, r+8, ... of its channel vector (f == 1).
            const int Wo = p.a.cv_w, Ho = p.a.cv_h, up = p.a.cv_up, cin = p.a.cv_cin;
            const int Ws = Wo >> up;
            const uint32_t tab32 = conv_tab32 + static_cast<uint32_t>(warp) * 1024u;
            if (tab_tile != tile) {
              const int Hs = Ho >> up, hw = p.a.hw;
#pragma unroll 1
              for (int i = lane; i < BM; i += 32) {
                const long long m = m0 + i;
                uint32_t off = 0, hwp = 0x7fffu;
                if (m < p.M) {
                  const int mi = static_cast<int>(m);
                  const int fr = mi / hw, rem = mi - fr * hw;
                  const int ho = rem / Wo, wo = rem - ho * Wo;
                  off = static_cast<uint32_t>(((fr * Hs + (ho >> up)) * Ws + (wo >> up)) * cin);
                  hwp = static_cast<uint32_t>(ho) | (static_cast<uint32_t>(wo) << 16);
                }
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(tab32 + i * 8), "r"(off), "r"(hwp) : "memory");
              }
              __syncwarp();
              tab_tile = tile;
            }
            const bool affine = p.a.scale != nullptr;
#pragma unroll 1
            for (int round = 0; round < (affine ? 2 : 1); ++round) {
              // round 0: the copies; round 1 (lazy BatchNorm + activation of the producer layer): the live pixels in place
              RowOp ac = p.a;
              ac.mode = EHGR_ROW_AFFINE;
#pragma unroll 1
              for (int pass = 0; pass * 4 < kv; ++pass) {
                const int k8 = pass * 4 + (slot % kvp);
                const int k = k_base + k8 * 8;
                if (k8 >= kv) continue;
                const bool kin = k < p.K;
                const int rg0 = slot / kvp;
                const int tap = kin ? k / cin : 0, ty = tap / 3;
                const int dy = ty - 1, dx = tap - ty * 3 - 1, c = k - tap * cin;
                const int tapoff = (dy * Ws + dx) * cin + c;       // up == 0: the tap is a constant element offset
                RowLoader<__nv_bfloat16, 8, false, false> ld;
                if (round) {
                  if (!kin) continue;
                  ld.init(ac, c, cin);
                }
                uint32_t dst = a_dst32 + rg0 * a_sbo + k8 * 128 + r * 16;
                uint32_t tp = tab32 + (rg0 * 8 + r) * 8;
                const uint32_t dstep = static_cast<uint32_t>(f * a_sbo), tstep = static_cast<uint32_t>(f * 64);
#pragma unroll 4
                for (int rg = rg0; rg < 16; rg += f, dst += dstep, tp += tstep) {
                  uint32_t off, hwp;
                  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(off), "=r"(hwp) : "r"(tp));
                  const int ho = static_cast<int>(hwp & 0xffffu), wo = static_cast<int>(hwp >> 16);
                  const bool live = kin && static_cast<unsigned>(ho + dy) < static_cast<unsigned>(Ho) &&
                                    static_cast<unsigned>(wo + dx) < static_cast<unsigned>(Wo);
                  if (round) {
                    if (live) {
                      RowLoader<__nv_bfloat16, 8, false, false>::Raw raw;
                      raw.a = lds128(dst);
                      sts128(dst, ld.finish_packed(ac, raw));
                    }
                    continue;
                  }
                  // up == 1: (ho + dy) >> 1 = (ho >> 1) + ((dy + (ho & 1)) >> 1), likewise for the column
                  const int o = up ? (((dy + (ho & 1)) >> 1) * Ws + ((dx + (wo & 1)) >> 1)) * cin + c : tapoff;
                  const __nv_bfloat16* src = in1 + (static_cast<int>(off) + o);
                  cp_async16(dst, live ? src : in1, live ? 16u : 0u);
                }
              }
              if (round == 0) {
                if (!p.b_resident && p.w16) stream_b(a_dst + p.a_bytes, s, k_base, kvalid);
                cp_async_wait_all();
              }
            }
          } else {
          // SHIFT: frame / segment index of the tile's first row (rows advance by < 128 inside a stage)
          int t0 = 0, rem0 = 0;
          if (mode == EHGR_ROW_SHIFT) {
            const long long f0 = m0 / p.a.hw;
            rem0 = static_cast<int>(m0 - f0 * p.a.hw);
            t0 = static_cast<int>(f0 % p.a.n_segment);
          }
#pragma unroll 1
          for (int pass = 0; pass * 4 < kv; ++pass) {
            const int k8 = pass * 4 + (slot % kvp);
            const int k = k_base + k8 * 8;
            const bool kin = k8 < kv && k < p.K;
            int cls = 2;               // SHIFT: 0 reads frame t+dir, 1 reads t-dir, 2 reads t, 3 straddles a fold boundary
            if (mode == EHGR_ROW_SHIFT && kin) {
              const int fold = p.a.fold;
              const int cl = k < fold ? 0 : (k < 2 * fold ? 1 : 2);
              const int ch = k + 7 < fold ? 0 : (k + 7 < 2 * fold ? 1 : 2);
              cls = cl == ch ? cl : 3;
            }
            if (mode != EHGR_ROW_SHIFT || cls == 2) {
              // plain copy: destination / source / live-row count advance by constants (no per-chunk index math)
              if (k8 < kv) {
                const int rg0 = slot / kvp;
                const long long mrow = m0 + rg0 * 8 + r;
                uint32_t dst = a_dst32 + rg0 * a_sbo + k8 * 128 + r * 16;
                const __nv_bfloat16* src = in1 + mrow * p.K + k;
                const long long left64 = kin ? p.M - mrow : 0;
                int left = left64 > 4096 ? 4096 : static_cast<int>(left64);    // > 0: row exists
                const uint32_t dst_step = static_cast<uint32_t>(f * a_sbo);
                const long long src_step = 8LL * f * p.K;
#pragma unroll 4
                for (int rg = rg0; rg < 16; rg += f) {
                  const bool live = left > 0;
                  cp_async16(dst, live ? src : in1, live ? 16u : 0u);
                  dst += dst_step;
                  src += src_step;
                  left -= 8 * f;
                }
              }
              continue;
            }
            const int dir = p.a.shift_dir < 0 ? -1 : 1;
            const long long step = static_cast<long long>(dir) * p.a.hw * p.K;
#pragma unroll 4
            for (int j = 0; j * f < 16; ++j) {
              const int rg = j * f + slot / kvp;
              if (!(k8 < kv && rg < 16)) continue;
              const int d = rg * 8 + r;
              const long long m = m0 + d;
              bool live = kin && m < p.M;
              const uint32_t dst = a_dst32 + rg * a_sbo + k8 * 128 + r * 16;
              const __nv_bfloat16* src = in1 + m * p.K + k;
              if (mode == EHGR_ROW_SHIFT && cls != 2 && live) {
                int rem = rem0 + d, t = t0;
                while (rem >= p.a.hw) { rem -= p.a.hw; t = t + 1 == p.a.n_segment ? 0 : t + 1; }
                if (cls == 3) {
                  // The vector straddles a fold boundary: four asynchronous 4-byte copies, one per channel pair, each
                  // from the frame its class reads (zero fill at the clip ends).  A pair is split only by an ODD fold
                  // (channels fold-1 | fold): its upper channel travels to a per-warp scratch word and is patched in
                  // after the wait — no synchronous global load anywhere on this path.
                  const int fold = p.a.fold;
                  const bool ok0 = t + dir >= 0 && t + dir < p.a.n_segment, ok1 = t - dir >= 0 && t - dir < p.a.n_segment;
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const int c = k + 2 * q;
                    const int c_lo = c < fold ? 0 : (c < 2 * fold ? 1 : 2), c_hi = c + 1 < fold ? 0 : (c + 1 < 2 * fold ? 1 : 2);
                    const __nv_bfloat16* sq = src + 2 * q;
                    const bool lv = c_lo == 2 || (c_lo == 0 ? ok0 : ok1);
                    cp_async4(dst + 4 * q, lv ? sq + (c_lo == 0 ? step : c_lo == 1 ? -step : 0) : in1, lv ? 4u : 0u);
                    if (c_hi != c_lo) {
                      const bool lh = c_hi == 2 || (c_hi == 0 ? ok0 : ok1);
                      cp_async4(shift_scr32 + static_cast<uint32_t>(d) * 4u, lh ? sq + (c_hi == 0 ? step : c_hi == 1 ? -step : 0) : in1,
                                lh ? 4u : 0u);
                    }
                  }
                  continue;
                }
                const int tt = cls == 0 ? t + dir : t - dir;       // frame the data comes from
                live = tt >= 0 && tt < p.a.n_segment;
                src += cls == 0 ? step : -step;
              }
              cp_async16(dst, live ? src : in1, live ? 16u : 0u);
            }
          }
          if (!p.b_resident && p.w16) stream_b(a_dst + p.a_bytes, s, k_base, kvalid);
          cp_async_wait_all();
          if (mode == EHGR_ROW_SHIFT && (p.a.fold & 1)) {
            // odd fold: the pair (fold-1, fold) was copied from the frame of its lower channel; patch the upper one
            const int fold = p.a.fold;
            const int k8 = (fold - 1 - k_base) >> 3;                   // vector holding channels fold-1 and fold (one pair)
            if (fold - 1 >= k_base && k8 < kv && (slot % kvp) == (k8 & 3)) {
              const uint32_t e = static_cast<uint32_t>(fold - (k_base + k8 * 8)) * 2u;   // byte offset of channel `fold`
#pragma unroll 4
              for (int j = 0; j * f < 16; ++j) {
                const int rg = j * f + slot / kvp;
                const int d = rg * 8 + r;
                if (rg < 16 && m0 + d < p.M) {
                  uint16_t v;
                  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(shift_scr32 + static_cast<uint32_t>(d) * 4u + 2u));
                  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a_dst32 + rg * a_sbo + k8 * 128 + r * 16 + e), "h"(v) : "memory");
                }
              }
            }
          }
          if (mode == EHGR_ROW_AFFINE) {            // the hot one: constant mode, no GATE code in the loop
            RowOp ac = p.a;
            ac.mode = EHGR_ROW_AFFINE;
#pragma unroll 1
            for (int pass = 0; pass * 4 < kv; ++pass) {
              const int k8 = pass * 4 + (slot % kvp);
              const int k = k_base + k8 * 8;
              if (!(k8 < kv && k < p.K)) continue;
              RowLoader<__nv_bfloat16, 8, false, false> ld;
              ld.init(ac, k, p.K);
#pragma unroll 4
              for (int j = 0; j * f < 16; ++j) {
                const int rg = j * f + slot / kvp;
                const long long m = m0 + rg * 8 + r;
                if (rg < 16 && m < p.M) {
                  const uint32_t dst = a_dst32 + rg * a_sbo + k8 * 128 + r * 16;
                  RowLoader<__nv_bfloat16, 8, false, false>::Raw raw;
                  raw.a = lds128(dst);
                  sts128(dst, ld.finish_packed(ac, raw));
                }
              }
            }
          } else if (mode == EHGR_ROW_GATE) {
#pragma unroll 1
            for (int pass = 0; pass * 4 < kv; ++pass) {
              const int k8 = pass * 4 + (slot % kvp);
              const int k = k_base + k8 * 8;
              if (!(k8 < kv && k < p.K)) continue;
              RowLoader<__nv_bfloat16, 8, false, true> ld;
              ld.init(p.a, k, p.K);
#pragma unroll 4
              for (int j = 0; j * f < 16; ++j) {
                const int rg = j * f + slot / kvp;
                const long long m = m0 + rg * 8 + r;
                if (rg < 16 && m < p.M) {
                  const uint32_t dst = a_dst32 + rg * a_sbo + k8 * 128 + r * 16;
                  RowLoader<__nv_bfloat16, 8, false, true>::Raw raw;
                  raw.a = lds128(dst);
                  raw.b[0].x = static_cast<uint32_t>(m & 0xffffffffLL);
                  raw.b[0].y = static_cast<uint32_t>(m >> 32);
                  sts128(dst, ld.finish_packed(p.a, raw));
                }
              }
            }
          }
          }   // mode != CONV3
        } else {
        bool waited = false;
#pragma unroll 1
        for (int pass = 0; pass * 4 < kv; ++pass) {
          const int k8 = pass * 4 + (slot % kvp);
          const int k = k_base + k8 * 8;
          const bool kin = k8 < kv && k < p.K;
          RowLoader<__nv_bfloat16, 8, true, true> ld;
          if (kin) ld.init(p.a, k, p.K);
#pragma unroll 1
          for (int j0 = 0; j0 * f < 16; j0 += kBatch) {
            RowLoader<__nv_bfloat16, 8, true, true>::Raw raw[kBatch];
            bool live[kBatch];
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
              const int rg = (j0 + j) * f + slot / kvp;
              const long long m = m0 + rg * 8 + r;
              live[j] = kin && rg < 16 && m < p.M;
              if (live[j]) raw[j] = ld.fetch(p.a, m);
            }
            if (!waited) { mbar_wait(bar_empty + 8 * s, parity); waited = true; }
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
              const int rg = (j0 + j) * f + slot / kvp;
              if (k8 < kv && rg < 16) {
                uint4 packed = make_uint4(0, 0, 0, 0);
                if (live[j]) {
                  if (kMode == EHGR_ROW_PLAIN) {
                    packed = raw[j].a;                      // already bf16: a straight 16-byte copy
                  } else {
                    float v[8];
                    ld.finish(p.a, raw[j], v);
                    packed = pack8(v);
                  }
                }
                *reinterpret_cast<uint4*>(a_dst + rg * a_sbo + k8 * 128 + r * 16) = packed;
              }
            }
          }
        }
        }
        if (!p.b_resident && !(kAsync && p.w16)) {       // register-path operand (BNBWD) or fp32 weights without a mirror
          stream_b(a_dst + p.a_bytes, s, k_base, kvalid);
          cp_async_wait_all();
        }
        fence_proxy_async();
        __syncwarp();                                // every lane's writes are fenced before the single arrival
        if (lane == 0) mbar_arrive(bar_full + 8 * s);
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA ISSUER =====================
    if (lane == 0) {
      // The single issuing thread is the serial resource of the whole pipeline (its per-stage instruction
      // count bounds the stage rate), so everything that does not change is hoisted: descriptor high words,
      // the ring's start-address field per stage (incremental), K-steps of the last stage.
      const uint32_t idesc = make_idesc(BM, p.BN, 0, p.w_is_kn ? 1 : 0);
      // A: no-swizzle core matrices (SBO = bytes between 8-row groups), or — kTma — SWIZZLE_128B rows (SBO = 1024 bytes,
      // layout type 2 in descriptor bits 61-63; a K=16 step advances the start address by 32 bytes inside the atom)
      // (CONV3 with 32-channel pixels: SWIZZLE_64B — layout type 4, SBO 512 bytes; K steps 0,1 in the first tap's box, 2,3 in the second)
      const uint32_t hi_a = kTma ? (p.cv_sw64 ? (32u | (1u << 14) | (4u << 29)) : (64u | (1u << 14) | (2u << 29)))
                                 : (((static_cast<uint32_t>(a_sbo) >> 4) & 0x3FFF) | (1u << 14));      // SBO | version
      constexpr uint32_t a_kstep = kTma ? 2u : 16u;
      // B: resident / cp.async slices = no-swizzle core matrices (SBO = Kp*16 or 1024 bytes, K step 256 bytes); TMA slices =
      // SWIZZLE_128B, K-major (SBO 1024, K step 32 bytes) or MN-major (LBO 8192 between 64-column blocks, SBO 1024, K step 2 KB)
      const uint32_t hi_b = p.b_tma ? (64u | (1u << 14) | (2u << 29))
                                    : (((p.b_resident ? static_cast<uint32_t>(Kp) : 64u) & 0x3FFF) | (1u << 14));
      const uint32_t b_kstep = p.b_tma ? (p.w_is_kn ? 128u : 2u) : 16u;
      const uint32_t lbo = (128u >> 4) << 16;
      const uint32_t lbo_b = (p.b_tma && p.w_is_kn) ? (8192u >> 4) << 16 : lbo;
      const uint32_t ring_lo = ((smem_u32(ring) & 0x3FFFF) >> 4) | lbo, bres_lo = ((smem_u32(smem) & 0x3FFFF) >> 4) | lbo;
      const uint32_t stage16 = static_cast<uint32_t>(p.stage_bytes) >> 4, abytes16 = static_cast<uint32_t>(p.a_bytes) >> 4;
      const int last_ksteps = (Kp - (k_stages - 1) * BK) >> 4;
      auto desc = [](uint32_t lo, uint32_t hi) { return (static_cast<uint64_t>(hi) << 32) | lo; };
      uint32_t ph = 0, a_lo = ring_lo, buf = 0, bph = 0;
      int s = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++buf) {
        if (buf == static_cast<uint32_t>(p.n_bufs)) { buf = 0; bph ^= 1; }
        mbar_wait(bar_tempty + 8 * buf, bph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * static_cast<uint32_t>(p.BN);
        for (int ks = 0; ks < k_stages; ++ks, ++s, a_lo += stage16) {
          if (s == p.n_stages) { s = 0; ph ^= 1; a_lo = ring_lo; }
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const int ksteps = ks == k_stages - 1 ? last_ksteps : BK / 16;
          // one K=16 step = two 8-element core matrices along K = 256 bytes = 16 address units
          const uint32_t b_lo = p.b_resident ? bres_lo + static_cast<uint32_t>(ks) * 64u : (((a_lo & 0xFFFFu) + abytes16) | lbo_b);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk)
            if (kk < ksteps)
              umma_bf16(d_tmem, desc(a_lo + ((kTma && p.cv_sw64) ? (kk >> 1) * 512u + (kk & 1) * 2u : kk * a_kstep), hi_a), desc(b_lo + kk * b_kstep, hi_b), idesc, (ks | kk) ? 1u : 0u);
          umma_commit(bar_empty + 8 * s);          // ring slot free once these MMAs have read it
        }
        umma_commit(bar_tfull + 8 * buf);          // accumulator complete
      }
    }
  } else {
    // ===================== EPILOGUE =====================
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int ew = warp - (kMmaWarp + 1);           // epilogue warp index 0..15
    const int part = ew >> 2;                       // which of the quarter's kParts warps
    const int n_cc = p.BN >> 4;                     // 16-column chunks of a tile
    const int cpp = (n_cc + kParts - 1) / kParts;   // chunks per warp
    const int c_lo = min(part * cpp, n_cc), c_hi = min(c_lo + cpp, n_cc);   // this warp's chunks: a contiguous column range
    const int nch = c_hi - c_lo;
    uint8_t* tail = reinterpret_cast<uint8_t*>(bars) + kBarBytes;
    const uint32_t pitch = static_cast<uint32_t>(p.epi_pitch);
    const uint32_t stage32 = smem_u32(tail) + static_cast<uint32_t>(ew) * 32u * pitch;
    const uint32_t my_row32 = stage32 + static_cast<uint32_t>(lane) * pitch;
    // store phase: lane = (row group rgp, 16-byte piece pc) of the row segment; ppr valid pieces per row
    const int n_sub = n0 + c_lo * 16;
    int nv = p.N - n_sub;                            // valid columns of this warp (multiple of 8)
    nv = nv < 0 ? 0 : (nv > nch * 16 ? nch * 16 : nv);
    const int ppr = nv >> 3;
    int lg = 0;
    while ((1 << lg) < ppr) ++lg;                    // pieces per row padded to a power of two (1, 2, 4, 8)
    const int pc = lane & ((1 << lg) - 1), rgp = lane >> lg, rstep = 32 >> lg;
    const bool pc_on = pc < ppr;
    float2 st_sum[4], st_sq[4];                      // statistics of this lane's eight channels, all tiles
#pragma unroll
    for (int i = 0; i < 4; ++i) st_sum[i] = st_sq[i] = make_float2(0.f, 0.f);
    uint32_t buf = 0, bph = 0;
    const int tile_step = gridDim.x / p.n_chunks;
    int m_tile = blockIdx.x / p.n_chunks;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++buf, m_tile += tile_step) {
      if (buf == static_cast<uint32_t>(p.n_bufs)) { buf = 0; bph ^= 1; }
      const long long m_base = static_cast<long long>(m_tile) * p.bm_rows + q * 32;
      // rows this tile owns end at the next tile's first row (bm_rows < 128: a CONV3 box) or at M
      const long long m_end = min(p.M, static_cast<long long>(m_tile + 1) * p.bm_rows);
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * static_cast<uint32_t>(p.BN);
      __syncwarp();                                  // the previous tile's staging reads are complete
      if (p.addend && pc_on) {                       // coalesced prefetch of the addend rows into the staging rows
        const __nv_bfloat16* src = p.addend + (m_base + rgp) * p.N + n_sub + pc * 8;
        uint32_t dst = stage32 + static_cast<uint32_t>(rgp) * pitch + static_cast<uint32_t>(pc) * 16u;
        long long left = m_end - m_base - rgp;
        for (int r = rgp; r < 32; r += rstep) {
          const bool live = left > 0;
          cp_async16(dst, live ? src : p.addend, live ? 16u : 0u);
          dst += static_cast<uint32_t>(rstep) * pitch;
          src += static_cast<long long>(rstep) * p.N;
          left -= rstep;
        }
      }
      mbar_wait_warp(bar_tfull + 8 * buf, bph);      // lane 0 polls: 16 waiting warps must not flood the barrier
      tc_fence_after();
      if (p.addend) {
        cp_async_wait_all();
        __syncwarp();
      }
      // ---- TMEM -> registers -> (+ addend) -> bf16 staging row of this lane
#pragma unroll 1
      for (int j = 0; j < nch; j += 2) {
        uint32_t r0[16], r1[16];
        const bool two = j + 1 < nch;
        tmem_ld16_issue(t_base + (c_lo + j) * 16, r0);
        if (two) tmem_ld16_issue(t_base + (c_lo + j + 1) * 16, r1);
        tmem_ld_wait(r0);
        if (two) tmem_ld_wait(r1);
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          if (h2 == 1 && !two) break;
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(h2 ? r1[i] : r0[i]);
          const uint32_t dst = my_row32 + static_cast<uint32_t>(j + h2) * 32u;
          if (p.addend) {
            float ad[8];
            uint4 a4 = lds128(dst);
            load_vec<__nv_bfloat16, 8>(reinterpret_cast<const __nv_bfloat16*>(&a4), ad);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += ad[i];
            a4 = lds128(dst + 16);
            load_vec<__nv_bfloat16, 8>(reinterpret_cast<const __nv_bfloat16*>(&a4), ad);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[8 + i] += ad[i];
          }
          float lo[8], hi[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { lo[i] = v[i]; hi[i] = v[8 + i]; }
          sts128(dst, pack8(lo));
          sts128(dst + 16, pack8(hi));
        }
      }
      tc_fence_before();                             // last chunk read: hand the accumulator buffer back to the MMA warp
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
      // ---- staging rows -> global, statistics from the stored words
      if (pc_on) {
        __nv_bfloat16* dst = p.out + (m_base + rgp) * p.N + n_sub + pc * 8;
        uint32_t src = stage32 + static_cast<uint32_t>(rgp) * pitch + static_cast<uint32_t>(pc) * 16u;
        long long left = m_end - m_base - rgp;
        for (int r = rgp; r < 32; r += rstep) {
          if (left > 0) {
            const uint4 w = lds128(src);
            stg128(dst, w);
            if (p.stats) {                           // rows past M are never stored and never counted
              const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f = bf2_to_f2(ww[i]);
                st_sum[i] = __fadd2_rn(st_sum[i], f);
                st_sq[i] = __ffma2_rn(f, f, st_sq[i]);
              }
            }
          }
          src += static_cast<uint32_t>(rstep) * pitch;
          dst += static_cast<long long>(rstep) * p.N;
          left -= rstep;
        }
      }
    }
    if (p.stats) {
      // reduce over the row-group lanes that share a piece column (fixed order), then one pair of atomics per channel
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        for (int o = 16; o >= (1 << lg) && o > 0; o >>= 1) {
          st_sum[i].x += __shfl_xor_sync(0xffffffffu, st_sum[i].x, o);
          st_sum[i].y += __shfl_xor_sync(0xffffffffu, st_sum[i].y, o);
          st_sq[i].x += __shfl_xor_sync(0xffffffffu, st_sq[i].x, o);
          st_sq[i].y += __shfl_xor_sync(0xffffffffu, st_sq[i].y, o);
        }
      }
      if (pc_on && rgp == 0) {
        const int n = n_sub + pc * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          atomicAdd(&p.stats[n + 2 * i], static_cast<double>(st_sum[i].x));
          atomicAdd(&p.stats[n + 2 * i + 1], static_cast<double>(st_sum[i].y));
          atomicAdd(&p.stats[p.N + n + 2 * i], static_cast<double>(st_sq[i].x));
          atomicAdd(&p.stats[p.N + n + 2 * i + 1], static_cast<double>(st_sq[i].y));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == kMmaWarp) {
    __syncwarp();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
  // every epilogue warp's statistics atomics precede the barrier above; only those warps fence
  bn_finalize_if_last(p.fin, p.stats, p.N, warp > kMmaWarp && p.stats != nullptr);
}

}  // namespace tc

bool pw_gemm_tc_supported(const RowOp& a, int w_is_kn, long long M, int K, int N, int dtype) {
  (void)a; (void)w_is_kn;
  if (dtype != EHGR_BF16) return false;
  if (K % 8 || N % 8 || K < 8 || N < 8) return false;
  if (M < 1 || M / tc::BM > 0x3fffffff) return false;
  return true;
}

int pw_gemm_tc(const RowOp& a, const float* w, const void* w16, int w_is_kn, void* out, const void* addend,
               double* stats, long long M, int K, int N, cudaStream_t s) {
  tc::GemmArgs p;
  p.a = a; p.w = w; p.w_is_kn = w_is_kn;
  p.w16 = static_cast<const __nv_bfloat16*>(w16);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.addend = static_cast<const __nv_bfloat16*>(addend);
  p.stats = stats;
  p.fin = take_fin();
  p.M = M; p.K = K; p.N = N;
  p.m_tiles = static_cast<int>(cdiv(M, tc::BM));
  p.bm_rows = tc::BM;
  p.cv_hbox = p.cv_nbox = p.cv_tiles_per_frame = 1;
  p.cv_sw64 = 0;
  // CONV3 im2col by TMA: plain operand at the output resolution, whole 64-channel stages, and a tile geometry of whole
  // image rows (cv_hbox divides the height) or whole frames that fits the 128 rows of a UMMA tile
  bool conv_tma = false;
  if (a.mode == EHGR_ROW_CONV3 && !a.scale && !a.cv_up && (a.cv_cin % 64 == 0 || a.cv_cin == 32) && a.cv_w <= 128 && w16) {
    p.cv_sw64 = a.cv_cin == 32 ? 1 : 0;
    const int hw = a.cv_h * a.cv_w;
    if (hw <= tc::BM) {
      p.cv_hbox = a.cv_h;
      p.cv_nbox = tc::BM / hw;
      p.cv_tiles_per_frame = 1;
      conv_tma = true;
    } else {
      for (int hb = tc::BM / a.cv_w; hb >= 1; --hb)
        if (a.cv_h % hb == 0) { p.cv_hbox = hb; break; }
      p.cv_nbox = 1;
      p.cv_tiles_per_frame = a.cv_h / p.cv_hbox;
      conv_tma = p.cv_hbox * a.cv_w >= 64;            // at least half a tile of real rows
    }
    if (conv_tma) {
      p.bm_rows = p.cv_nbox * p.cv_hbox * a.cv_w;
      const long long frames = M / hw;
      p.m_tiles = static_cast<int>(cdiv(frames, p.cv_nbox) * p.cv_tiles_per_frame);
    }
  }
  p.n_bufs = 2;   // more buffers bought nothing (the hand-shake is not the bound) and a 512-column allocation
                  // makes the next kernel's CTAs wait for TMEM
  const int Kp = (K + 15) & ~15;
  constexpr int kBudget = 200 * 1024;
  // TMA + SWIZZLE_128B operand path: the modes that read ONE tensor row by row (the in1 rows are the GEMM rows)
  const bool use_tma = (a.mode == EHGR_ROW_PLAIN || a.mode == EHGR_ROW_AFFINE || a.mode == EHGR_ROW_GATE) && M < 0x7fffffffLL;
  p.a_bytes = (use_tma || conv_tma) ? tc::BM * tc::BK * 2 : tc::BM * std::min(Kp, tc::BK) * 2;   // a box is always 128 x 128 bytes
  const int pad = 1023;                             // resident weights padded to the 1 KB stage alignment
  // streamed weight slices by TMA (needs the bf16 mirror and whole 64-wide K stages on the operand side)
  const bool b_tma_ok = w16 != nullptr && Kp >= tc::BK;
  auto slice_bytes = [&](int bn) { return (b_tma_ok && w_is_kn) ? (bn + 63) / 64 * 8192 : bn * tc::BK * 2; };
  // Output columns per tile: as few column chunks as possible (A is re-read once per chunk), but the epilogue
  // staging (the whole bf16 output tile) and the weights share the 200 KB with the operand ring: take the first
  // chunk count that leaves at least four ring stages, else the one with the deepest ring.  Input-heavy shapes
  // (K >= N: the project / dgrad-of-expand layers) never take extra chunks: re-reading A costs more than a shallow ring.
  const int Np = (N + 15) & ~15;
  const int epi_warps = (K >= N || a.mode == EHGR_ROW_CONV3) ? 8 : 16, parts = epi_warps / 4;
  int best_chunks = 0, best_stages = -1, bar_bytes = 0, b_res = 0;
  const int conv_tab = conv_tma ? 0 : a.mode == EHGR_ROW_CONV3 ? tc::kProducerWarps * 1024          // per-warp row tables
                       : (a.mode == EHGR_ROW_SHIFT && (a.fold & 1)) ? tc::kProducerWarps * 512 : 0;   // odd-fold scratch words
  const int min_chunks = (Np + 255) / 256;
  // (CONV3: K = 9*cin is large and the operand comes out of L2 — a deeper ring is worth narrower tiles)
  const bool extra_chunks = epi_warps != 8 || a.mode == EHGR_ROW_CONV3;
  for (int chunks = min_chunks; chunks <= min_chunks + (extra_chunks ? 3 : 0); ++chunks) {
    int bn = (Np / chunks + 15) & ~15;
    while (bn * chunks < Np) bn += 16;
    if (bn > 256 || bn < 16) continue;
    const int pitch = ((bn >> 4) + parts - 1) / parts * 32 + 16;
    const int bar = tc::kTailBytes + epi_warps * 32 * pitch + conv_tab;
    const int bres = (bn * Kp * 2 + pad) & ~pad;
    const bool resident = bres + 6 * p.a_bytes + bar <= kBudget;
    const int stage = p.a_bytes + (resident ? 0 : slice_bytes(bn));
    const int stages = std::min(tc::kMaxStages, (kBudget - bar - (resident ? bres : 0)) / stage);
    if (stages > best_stages) { best_stages = stages; best_chunks = chunks; }
    if (stages >= 4) { best_chunks = chunks; break; }
  }
  p.n_chunks = best_chunks;
  p.BN = (Np / p.n_chunks + 15) & ~15;
  while (p.BN * p.n_chunks < Np) p.BN += 16;
  int cols = 32;
  while (cols < p.n_bufs * p.BN) cols <<= 1;
  p.tmem_cols = cols;
  const int n_cc = p.BN >> 4;
  p.epi_pitch = (n_cc + parts - 1) / parts * 32 + 16;           // odd multiple of 16 bytes: conflict-free row stores
  bar_bytes = tc::kTailBytes + epi_warps * 32 * p.epi_pitch + conv_tab;
  b_res = (p.BN * Kp * 2 + pad) & ~pad;
  // weights resident when that still leaves >= 6 A stages (the large-M layers all qualify)
  p.b_resident = (b_res + 6 * p.a_bytes + bar_bytes <= kBudget) ? 1 : 0;
  p.b_tma = (!p.b_resident && b_tma_ok) ? 1 : 0;
  p.b_bytes = p.b_resident ? 0 : slice_bytes(p.BN);
  p.stage_bytes = p.a_bytes + p.b_bytes;
  const int avail = kBudget - bar_bytes - (p.b_resident ? b_res : 0);
  p.n_stages = std::max(2, std::min(tc::kMaxStages, avail / p.stage_bytes));
  const long long tiles = static_cast<long long>(p.m_tiles) * p.n_chunks;
  // persistent grid: a multiple of n_chunks (a CTA keeps one column chunk -> statistics stay in registers)
  long long grid = std::min<long long>(tiles, kNumSMs);
  grid = std::max<long long>(p.n_chunks, grid / p.n_chunks * p.n_chunks);
  const size_t smem = static_cast<size_t>(p.b_resident ? b_res : 0) + static_cast<size_t>(p.n_stages) * p.stage_bytes + bar_bytes +
                      1024;                          // + room to round the base up to 1 KB
  CUtensorMap tm_a;
  memset(&tm_a, 0, sizeof(tm_a));
  if (use_tma)
    if (int st = tma::make_map_2d_sw128(&tm_a, a.in1, static_cast<unsigned long long>(K), static_cast<unsigned long long>(M), tc::BM))
      return st;
  if (conv_tma) {
    const unsigned long long cin = a.cv_cin, wo = a.cv_w, ho = a.cv_h, frames = M / a.hw;
    const unsigned long long dims[4] = {cin, wo, ho, frames};
    const unsigned long long strides[3] = {cin * 2, wo * cin * 2, ho * wo * cin * 2};
    const unsigned box[4] = {p.cv_sw64 ? 32u : 64u, static_cast<unsigned>(wo), static_cast<unsigned>(p.cv_hbox),
                             static_cast<unsigned>(p.cv_nbox)};
    if (int st = tma::make_map_4d(&tm_a, 2, a.in1, dims, strides, box, p.cv_sw64 ? 64 : 128)) return st;
  }
  CUtensorMap tm_b;
  memset(&tm_b, 0, sizeof(tm_b));
  if (p.b_tma) {
    const int st = w_is_kn ? tma::make_map_2d_sw128(&tm_b, w16, static_cast<unsigned long long>(N), static_cast<unsigned long long>(K), 64)
                           : tma::make_map_2d_sw128(&tm_b, w16, static_cast<unsigned long long>(K), static_cast<unsigned long long>(N),
                                                    static_cast<unsigned>(p.BN));
    if (st) return st;
  }
  constexpr int kSmemMax = kBudget + 1024;
  auto go = [&](auto mode_tag, auto tma_tag) {
    constexpr int kMode = decltype(mode_tag)::value;
    constexpr bool kTma = decltype(tma_tag)::value;
    if (epi_warps == 16) {
      ensure_smem(tc::pw_gemm_tc_kernel<kMode, 16, kTma>, kSmemMax);
      tc::pw_gemm_tc_kernel<kMode, 16, kTma><<<static_cast<unsigned>(grid), tc::threads_of(16), smem, s>>>(p, tm_a, tm_b);
    } else {
      ensure_smem(tc::pw_gemm_tc_kernel<kMode, 8, kTma>, kSmemMax);
      tc::pw_gemm_tc_kernel<kMode, 8, kTma><<<static_cast<unsigned>(grid), tc::threads_of(8), smem, s>>>(p, tm_a, tm_b);
    }
  };
  using T = std::true_type;
  using F = std::false_type;
  switch (a.mode) {
    case EHGR_ROW_PLAIN:
      if (use_tma) go(std::integral_constant<int, EHGR_ROW_PLAIN>{}, T{});
      else go(std::integral_constant<int, EHGR_ROW_PLAIN>{}, F{});
      break;
    case EHGR_ROW_AFFINE:
      if (use_tma) go(std::integral_constant<int, EHGR_ROW_AFFINE>{}, T{});
      else go(std::integral_constant<int, EHGR_ROW_AFFINE>{}, F{});
      break;
    case EHGR_ROW_GATE:
      if (use_tma) go(std::integral_constant<int, EHGR_ROW_GATE>{}, T{});
      else go(std::integral_constant<int, EHGR_ROW_GATE>{}, F{});
      break;
    case EHGR_ROW_SHIFT: go(std::integral_constant<int, EHGR_ROW_SHIFT>{}, F{}); break;
    case EHGR_ROW_CONV3:   // K = 9*cin >= N for every decoder layer: the 8-epilogue-warp form only
      if (conv_tma) {
        ensure_smem(tc::pw_gemm_tc_kernel<EHGR_ROW_CONV3, 8, true>, kSmemMax);
        tc::pw_gemm_tc_kernel<EHGR_ROW_CONV3, 8, true><<<static_cast<unsigned>(grid), tc::threads_of(8), smem, s>>>(p, tm_a, tm_b);
      } else {
        ensure_smem(tc::pw_gemm_tc_kernel<EHGR_ROW_CONV3, 8, false>, kSmemMax);
        tc::pw_gemm_tc_kernel<EHGR_ROW_CONV3, 8, false><<<static_cast<unsigned>(grid), tc::threads_of(8), smem, s>>>(p, tm_a, tm_b);
      }
      break;
    default: go(std::integral_constant<int, EHGR_ROW_BNBWD>{}, F{}); break;
  }
  return launch_status();
}

}  // namespace ehgr
