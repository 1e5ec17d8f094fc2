// Fused ResBlock1 pair on tcgen05, second generation of tc_pair.cu (A operand in shared memory, "SS" form):
//     y = conv2(lrelu(conv1(lrelu(x)) + b1)) + b2 + x      C = N in {32, 64}
// (one iteration of the loop at vietTTS/hifigan/model.py:44-51: dilated conv, conv, residual add).
//
// tc_pair.cu runs conv1 -> epilogue -> conv2 -> epilogue of a tile one after the other from ONE issuing warp; measured
// (profiles/r2_umma_probe2.txt) an `if (elect_one())` issue region costs ~180 clk before its first MMA and every wait
// ~84 clk, during which the tensor pipe of a single issuer idles, while several issuing warps together reach the pipe's
// own cost of an M=128, K=16 shared-memory-operand MMA (47-57 clk for N <= 64).  Here the two convs are two DECOUPLED
// pipelines with one issuing warp each:
//   conv1:  converters (global fp32 -> lrelu -> hi/lo bf16, smem stage per 16-channel chunk) -> issuer A -> D1
//   conv2:  E1 warps (D1 -> +b1, lrelu, zero padding, hi/lo, smem) -> issuer B -> D2 -> E2 warps (+b2 +x -> y)
// so conv1 of tile i+1 overlaps conv2 of tile i and both epilogues; the accumulators are copied to registers and
// released before the epilogue arithmetic, the intermediate never leaves the chip (8 instead of 20 B of HBM traffic
// per element pair).  No collector hints: MMAs of the two issuers interleave in the tensor pipe.
// With both pipelines busy the tensor pipe's cost per shared-memory-operand MMA is the limit (measured: the issuers are
// blocked on the pipe 78 % of the time), so the three products of bf16x3 are issued as TWO instructions:
//   [main | aux] (+)= a_hi . [W_hi | W_lo]   (N' = 2N, the packed weight stage keeps W_hi and W_lo rows adjacent)
//    main         += a_lo . W_hi             (N' = N)
// and the epilogues add main + aux (143 instead of 171 clk per triple at C = 64, 104 instead of 141 at C = 32).
// Arithmetic: bf16x3 exactly as tc_conv.cu / tc_pair.cu (same products, same summation order).
//
// Tile: conv1 produces R = 128*MT rows [o0-h2, o0-h2+R); conv2 consumes them and yields V = R-(k-1) valid rows.
#include <cuda_bf16.h>

#include <algorithm>

#include "tc_common.cuh"
#include "vtts_internal.cuh"

namespace {

using namespace tcx;

constexpr int NTHREADS = 896;      // 28 warps (72 registers per thread)
// 0-3 E1 (D1 -> conv2 operand) | 4-11 E2 (D2 -> y, two warps per TMEM lane quadrant) | 12,13 MMA issuers | 14,15 weight
// producers | 16-27 converters (three groups of four warps)                               (warp % 4 = TMEM lane quadrant)
constexpr int W_E2 = 4, W_ISSA = 12, W_ISSB = 13, W_WP1 = 14, W_WP2 = 15, W_CONV = 16;
constexpr int NGRP = 3, GRP_THREADS = 128;
constexpr int NA = 4;              // conv1 activation stages (one 16-channel chunk each)

// PAIR = 1: the CTA-pair form (cf. tc_conv.cu): a cluster of two CTAs works on two tiles of the SAME problem, rank 0 issues
// `tcgen05.mma.cta_group::2` (M = 256: rank r's 128 rows are its own tile) and each CTA holds only its half of the
// weight operand:  MMA 1 (N' = 2N):  rank 0 supplies W_hi, rank 1 W_lo;   MMA 2 (N' = N): rank r supplies W_hi[rN/2, (r+1)N/2).
// The weights come pre-packed per rank ([chunk][rank][tap][k-half][N rows | N/2 rows][8], vtts_tc_pack_weights_pc), so a
// group of taps is still ONE bulk copy per CTA; shared-memory operand bytes per product: 11 instead of 14 KB at C = 64.
template <int N, int PAIR_ = 0>
struct P2Cfg {
  static constexpr int PAIR = PAIR_;
  static constexpr int CPP = PAIR ? 2 : 1;
  static constexpr int MT = N == 64 ? 1 : 2;
  static constexpr int R = 128 * MT;
  static constexpr int G = 4;                        // taps per weight group (one bulk copy, one issue region)
  static constexpr int NCH = N / 16;
  static constexpr int RA1 = R + 64;                 // conv1 operand rows per stage (halo <= 50)
  static constexpr int RA2 = R + 16;                 // conv2 operand rows (halo <= 10)
  static constexpr int A1_STAGE = RA1 * 64;
  static constexpr int A2_CHUNK = RA2 * 64;
  static constexpr int A2_BUF = NCH * A2_CHUNK;
  static constexpr int W_STAGE = PAIR ? N * 48 : N * 64;   // one (chunk, tap): [k-half][hi|lo][n][8 bf16]; pair form: [k-half][N | N/2 rows][8]
  static constexpr int KH_ROWS = PAIR ? N + N / 2 : 2 * N;  // rows of one k-half block
  static constexpr int W_GROUP = G * W_STAGE;
  static constexpr int NWG = N == 64 ? 2 : 3;        // weight groups in flight per ring
  static constexpr int EPI_PITCH = 80;               // 16 floats + 16 B pad
  static constexpr int EPI_STAGE = 8 * 32 * EPI_PITCH;
  static constexpr int NBAR = 2 * NA + 4 * NWG + 4 + 4;
  static constexpr int SMEM_BYTES = NA * A1_STAGE + 2 * A2_BUF + 2 * NWG * W_GROUP + EPI_STAGE + NBAR * 8 + 16 + 1024;
  static constexpr int ACC = MT * 2 * N;             // columns of one accumulator: [main | aux] per M tile (= 128)
  static constexpr int NIT = MT * (N / 16);          // 16-column pieces of one accumulator (= 4)
  static_assert(SMEM_BYTES <= 232448, "shared memory");
};

// CL = 1: acquire at cluster scope (barriers of the pair form whose arrivals come from both CTAs)
template <bool PROF, int CL = 0>
__device__ __forceinline__ void mbar_wait_p(uint64_t* bar, uint32_t parity, int* err, int code, long long& acc) {
  if constexpr (CL) {
    long long dummy = 0;
    mbar_wait_tc<0>(bar, parity, err, code, PROF ? acc : dummy);
  } else if constexpr (PROF) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
      if (clock64() - t0 > SPIN_TIMEOUT) spin_fail(err, code);
    }
    acc += clock64() - t0;
  } else {
    (void)acc;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
      if (++spins > (1u << 24)) spin_fail(err, code);       // a try_wait suspends ~100+ clk: > 1 s in total
    }
  }
}

template <int N, bool PROF, int PAIR>
__global__ void __launch_bounds__(NTHREADS, 1) tc_pair2_kernel(const __grid_constant__ TcPairLaunch L) {
  using Cfg = P2Cfg<N, PAIR>;
  constexpr int MT = Cfg::MT, R = Cfg::R, G = Cfg::G, NCH = Cfg::NCH, RA1 = Cfg::RA1, RA2 = Cfg::RA2, NWG = Cfg::NWG;
  constexpr int CPP = Cfg::CPP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a1_st = smem;
  uint8_t* a2_st = a1_st + NA * Cfg::A1_STAGE;
  uint8_t* w_st = a2_st + 2 * Cfg::A2_BUF;                       // [ring][NWG][W_GROUP]
  uint8_t* epi_st = w_st + 2 * NWG * Cfg::W_GROUP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_st + Cfg::EPI_STAGE);
  uint64_t* a1_full = bars;                  // [NA]   converters -> issuer A
  uint64_t* a1_empty = a1_full + NA;         // [NA]
  uint64_t* w_full = a1_empty + NA;          // [ring][NWG]
  uint64_t* w_empty = w_full + 2 * NWG;      // [ring][NWG]
  uint64_t* d_full = w_empty + 2 * NWG;      // [conv]  issuer -> epilogue (single-buffered accumulators: the epilogue warps copy
  uint64_t* d_empty = d_full + 2;            // [conv]  them to registers and release them before their arithmetic)
  uint64_t* a2_full = d_empty + 2;           // [2]    E1 -> issuer B
  uint64_t* a2_empty = a2_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a2_empty + 2);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  // pair form: every barrier an issuer waits on lives in rank 0 and counts the arrivals of both CTAs (rank 1 arrives
  // through the cluster address space); the *_empty / d_full barriers are signalled in both CTAs by multicast commits
  const uint32_t prank = PAIR ? cluster_rank() : 0u;

  if (warp == W_WP1 && lane == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&a1_full[i], CPP * GRP_THREADS); mbar_init(&a1_empty[i], 1); }
    for (int i = 0; i < 2 * NWG; ++i) { mbar_init(&w_full[i], (PAIR && prank == 0) ? 2 : 1); mbar_init(&w_empty[i], 1); }
    mbar_init(&d_full[0], 1); mbar_init(&d_full[1], 1);
    mbar_init(&d_empty[0], CPP * 128); mbar_init(&d_empty[1], CPP * 256);
    for (int i = 0; i < 2; ++i) { mbar_init(&a2_full[i], CPP * 128); mbar_init(&a2_empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_ISSA) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  const int ntiles = L.ntiles;
// tile -> problem map: round robin (problem = tile % nprob), so row tile r of every problem is in flight at the same time
// and a tensor the problems SHARE (the three ResBlocks of a stage read the same input in their first pair) is fetched
// from DRAM once and served from L2 to the others; problems with fewer tiles (larger V) skip their surplus indices.
// Pair form: tile index -> two consecutive tiles (rest = 2 q + rank) of one problem; a CTA whose own tile does not exist or
// lies beyond its row's length still walks the pipeline with valid = 0 (zero operand, no stores) as long as its peer's
// tile is live -- the skip decision must be the same in both CTAs.
#define P2_TILE_BEGIN                                                                 \
  for (int tile = blockIdx.x / CPP; tile < ntiles; tile += gridDim.x / CPP) {         \
    const int pi = tile % L.nprob;                                                    \
    const TcPairProb& P = L.p[pi];                                                    \
    const int tpr = L.tiles_per_row[pi];                                              \
    const int k = P.k, dil = P.dil;                                                   \
    const int V = R - (k - 1);                                                        \
    int tt, b, o0, valid;                                                             \
    if constexpr (PAIR) {                                                             \
      bool any_live = false;                                                          \
      tt = 0; b = 0; o0 = 0; valid = 0;                                               \
      _Pragma("unroll") for (int rk = 0; rk < 2; ++rk) {                              \
        const int rest_ = (tile / L.nprob) * 2 + rk;                                  \
        int tt_ = 0, b_ = 0, valid_ = 0;                                              \
        if (rest_ < tpr * L.B) {                                                      \
          tt_ = rest_ % tpr;                                                          \
          b_ = rest_ / tpr;                                                           \
          valid_ = L.T_rows;                                                          \
          if (L.len) {                                                                \
            const int v_ = L.len[b_] * L.len_mul;                                     \
            valid_ = v_ < valid_ ? v_ : valid_;                                       \
          }                                                                           \
          if (tt_ * V >= valid_) valid_ = 0;                                          \
        }                                                                             \
        any_live |= valid_ > 0;                                                       \
        if (rk == (int)prank) { tt = tt_; b = b_; o0 = tt_ * V; valid = valid_; }     \
      }                                                                               \
      if (!any_live) continue;                                                        \
    } else {                                                                          \
      const int rest = tile / L.nprob;                                                \
      if (rest >= tpr * L.B) continue;                                                \
      tt = rest % tpr;                                                                \
      b = rest / tpr;                                                                 \
      o0 = tt * V;                                                                    \
      valid = L.T_rows;                                                               \
      if (L.len) {                                                                    \
        const int v_ = L.len[b] * L.len_mul;                                          \
        valid = v_ < valid ? v_ : valid;                                              \
      }                                                                               \
      if (o0 >= valid) continue;                                                      \
    }                                                                                 \
    const int h2 = (k - 1) / 2, h1 = ((k - 1) * dil) / 2;                             \
    const int ng = (k + G - 1) / G;
#define P2_TILE_END }

  if ((warp == W_ISSA || warp == W_ISSB) && PAIR && prank != 0) {
    // ============================ pair form, rank 1: weight-group forwarders (one per ring) ============================
    // this CTA's half of a weight group lands on its own w_full; tell the issuer of that ring in rank 0
    const int ring = warp == W_ISSA ? 0 : 1;
    uint32_t ws = 0, wph = 0;
    long long c_w = 0;
    P2_TILE_BEGIN
      (void)b; (void)h1; (void)h2; (void)o0; (void)tt; (void)dil; (void)V; (void)valid;
      for (int s_ = 0; s_ < NCH * ng; ++s_) {
        mbar_wait_p<PROF>(&w_full[ring * NWG + ws], wph, L.err, 34 + ring, c_w);
        if (elect_one()) mbar_arrive_rank<0>(&w_full[ring * NWG + ws], 0);
        __syncwarp();
        if (++ws == NWG) { ws = 0; wph ^= 1; }
      }
    P2_TILE_END
  } else if (warp == W_ISSA || warp == W_ISSB) {
    // ============================ MMA issuers: ring 0 = conv1, ring 1 = conv2 ============================
    const int ring = warp == W_ISSA ? 0 : 1;
    constexpr uint32_t idesc = PAIR ? make_idesc2(N) : make_idesc(N), idesc2 = PAIR ? make_idesc2(2 * N) : make_idesc(2 * N);
    const uint64_t b_tmpl = make_desc(0, Cfg::KH_ROWS * 16, 128);   // [k-half][hi|lo][n][8]: k-half blocks 2N rows apart (pair form: 1.5 N)
    const uint64_t a_tmpl = make_desc(0, (ring == 0 ? RA1 : RA2) * 16, 128);
    const uint32_t RAx = ring == 0 ? RA1 : RA2;
    const uint32_t w_ring_u32 = smem_u32(w_st + (size_t)ring * NWG * Cfg::W_GROUP);
    const uint32_t a1_u32 = smem_u32(a1_st), a2_u32 = smem_u32(a2_st);
    const uint32_t d0 = tmem_base + ring * Cfg::ACC;
    uint32_t dph = 0, ws = 0, wph = 0, item = 0, buf = 0, bph = 0;
    long long c_d = 0, c_a = 0, c_w = 0;
    const long long t_begin = PROF ? clock64() : 0;
    P2_TILE_BEGIN
      (void)b; (void)h1; (void)h2; (void)o0; (void)tt; (void)V; (void)valid;
      const int dl = ring == 0 ? dil : 1;
      mbar_wait_p<PROF, PAIR>(&d_empty[ring], dph ^ 1, L.err, 30 + ring, c_d);
      if (ring == 1) mbar_wait_p<PROF, PAIR>(&a2_full[buf], bph, L.err, 32, c_a);
      tc_fence_after();
      for (int c = 0; c < NCH; ++c) {
        uint32_t a_base16, sa = 0;
        if (ring == 0) {
          sa = item % NA;
          mbar_wait_p<PROF, PAIR>(&a1_full[sa], (item / NA) & 1, L.err, 33, c_a);
          tc_fence_after();
          a_base16 = (a1_u32 + sa * Cfg::A1_STAGE) >> 4;
          ++item;
        } else {
          a_base16 = (a2_u32 + buf * Cfg::A2_BUF + c * Cfg::A2_CHUNK) >> 4;
        }
        for (int g = 0; g < ng; ++g) {
          const int nt = (k - g * G) < G ? (k - g * G) : G;
          mbar_wait_p<PROF, PAIR>(&w_full[ring * NWG + ws], wph, L.err, 34 + ring, c_w);
          tc_fence_after();
          const uint32_t w_base16 = (w_ring_u32 + ws * Cfg::W_GROUP) >> 4;
          const uint32_t first_grp = (c | g) != 0 ? 1u : 0u;
          if (elect_one()) {
            for (int t = 0; t < nt; ++t) {
              const uint64_t b_hi = b_tmpl | (uint64_t)(w_base16 + t * (Cfg::W_STAGE >> 4));     // rows [0,N) = W_hi, [N,2N) = W_lo
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint32_t row = a_base16 + mt * 128 + (g * G + t) * dl;
                const uint64_t a_hi = a_tmpl | (uint64_t)row;
                const uint64_t a_lo = a_tmpl | (uint64_t)(row + 2 * RAx);
                const uint32_t d = d0 + mt * 2 * N;
                if constexpr (PAIR) {
                  // rows [0, N) of the k-half block: this rank's half of [W_hi | W_lo]; rows [N, 1.5 N): its half of W_hi
                  umma2<0>(d, a_hi, b_hi, idesc2, (first_grp | (uint32_t)t) != 0 ? 1u : 0u);
                  umma2<0>(d, a_lo, b_hi + N, idesc, 1u);
                } else {
                  umma<0>(d, a_hi, b_hi, idesc2, (first_grp | (uint32_t)t) != 0 ? 1u : 0u);   // [main | aux] (+)= a_hi . [W_hi | W_lo]
                  umma<0>(d, a_lo, b_hi, idesc, 1u);                                           //  main        += a_lo . W_hi
                }
              }
            }
            if constexpr (PAIR) umma_commit2(&w_empty[ring * NWG + ws]); else umma_commit(&w_empty[ring * NWG + ws]);
          }
          __syncwarp();
          if (++ws == NWG) { ws = 0; wph ^= 1; }
        }
        if (ring == 0) {
          if (elect_one()) { if constexpr (PAIR) umma_commit2(&a1_empty[sa]); else umma_commit(&a1_empty[sa]); }
          __syncwarp();
        }
      }
      if (elect_one()) {
        if constexpr (PAIR) {
          if (ring == 1) umma_commit2(&a2_empty[buf]);
          umma_commit2(&d_full[ring]);
        } else {
          if (ring == 1) umma_commit(&a2_empty[buf]);
          umma_commit(&d_full[ring]);
        }
      }
      __syncwarp();
      dph ^= 1;
      if (ring == 1 && ++buf == 2) { buf = 0; bph ^= 1; }
    P2_TILE_END
    if (PROF && L.dbg && lane == 0 && blockIdx.x < 128) {
      long long* d = L.dbg + (size_t)blockIdx.x * 32 + ring * 4;      // 0..3 issuer A, 4..7 issuer B: total, wait acc, wait operand, wait weights
      d[0] = clock64() - t_begin; d[1] = c_d; d[2] = c_a; d[3] = c_w;
    }
  } else if (warp == W_WP1 || warp == W_WP2) {
    // ============================ weight producers (one per ring) ============================
    if (lane == 0) {
      const int ring = warp == W_WP1 ? 0 : 1;
      uint8_t* wr = w_st + (size_t)ring * NWG * Cfg::W_GROUP;
      uint32_t ws = 0, wph = 0;
      long long c_e = 0;
      P2_TILE_BEGIN
        (void)b; (void)h1; (void)h2; (void)o0; (void)dil; (void)tt; (void)V; (void)valid;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(PAIR ? (ring == 0 ? P.w1pc : P.w2pc) : (ring == 0 ? P.w1pk : P.w2pk));
        for (int c = 0; c < NCH; ++c)
          for (int g = 0; g < ng; ++g) {
            const int nt = (k - g * G) < G ? (k - g * G) : G;
            const uint32_t bytes = (uint32_t)nt * Cfg::W_STAGE;
            mbar_wait_p<PROF>(&w_empty[ring * NWG + ws], wph ^ 1, L.err, 36 + ring, c_e);
            mbar_expect_tx(&w_full[ring * NWG + ws], bytes);
            // pair form: blocks are ordered [chunk][rank][tap], so this rank's taps of a chunk are contiguous as well
            const size_t blk = PAIR ? ((size_t)(c * 2 + (int)prank) * k + (size_t)g * G) : ((size_t)c * k + (size_t)g * G);
            bulk_g2s(wr + ws * Cfg::W_GROUP, wsrc + blk * Cfg::W_STAGE, bytes, &w_full[ring * NWG + ws]);
            if (++ws == NWG) { ws = 0; wph ^= 1; }
          }
      P2_TILE_END
    }
    __syncwarp();
  } else if (warp >= W_CONV) {
    // ============================ activation converters (conv1 input), three groups of four warps ============================
    const int ct = tid - W_CONV * 32;
    const int grp = ct / GRP_THREADS;
    const int gt = ct - grp * GRP_THREADS;
    const int q = gt & 3;
    const int r0 = gt >> 2;
    const float slope = L.slope;
    uint32_t item = 0;
    long long c_e = 0;
    const long long t_begin = PROF ? clock64() : 0;
    P2_TILE_BEGIN
      (void)ng; (void)tt; (void)V;
      const int rows = R + (k - 1) * dil;
      const float* x0 = P.x + (size_t)b * L.T_rows * N;
      const int row_base = o0 - h2 - h1;
      for (int c = 0; c < NCH; ++c, ++item) {
        if ((int)(item % NGRP) != grp) continue;
        const uint32_t sa = item % NA, pa = (item / NA) & 1;
        mbar_wait_p<PROF>(&a1_empty[sa], pa ^ 1, L.err, 38, c_e);
        uint8_t* st = a1_st + sa * Cfg::A1_STAGE + ((q >> 1) * RA1) * 16 + (q & 1) * 8;
        const int coff = c * 16 + q * 4;
        constexpr int U = 10;
        for (int rr0 = r0; rr0 < rows; rr0 += 32 * U) {
          float4 v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int rr = rr0 + u * 32;
            const int t = row_base + rr;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rr < rows && t >= 0 && t < valid) v[u] = ldg_pf256(x0 + (size_t)t * N + coff);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int rr = rr0 + u * 32;
            if (rr < rows) {
              float4 x = v[u];
              x.x = lrelu(x.x, slope); x.y = lrelu(x.y, slope); x.z = lrelu(x.z, slope); x.w = lrelu(x.w, slope);
              uint2 hi, lo;
              split4(x, hi, lo);
              *reinterpret_cast<uint2*>(st + (size_t)rr * 16) = hi;
              *reinterpret_cast<uint2*>(st + (size_t)(2 * RA1 + rr) * 16) = lo;
            }
          }
        }
        fence_proxy_async();                                   // generic-proxy stores -> visible to the tensor core's operand fetch
        if constexpr (PAIR) mbar_arrive_rank<0>(&a1_full[sa], 0); else mbar_arrive(&a1_full[sa]);
      }
    P2_TILE_END
    if (PROF && L.dbg && gt == 0 && blockIdx.x < 128) {
      long long* d = L.dbg + (size_t)blockIdx.x * 32 + 8 + grp * 2;     // 8..13 converter groups: total, wait stage free
      d[0] = clock64() - t_begin; d[1] = c_e;
    }
  } else if (warp < W_E2) {
    // ============================ E1: D1 -> + b1, leaky_relu, zero padding, hi/lo split -> conv2 operand ============================
    // The accumulator is copied to registers in two halves; D1 is released after the second copy, before its conversion.
    uint32_t dph = 0, buf = 0, bph = 0;
    const float slope = L.slope;
    long long c_df = 0, c_ae = 0;
    const long long t_begin = PROF ? clock64() : 0;
    P2_TILE_BEGIN
      (void)b; (void)h1; (void)dil; (void)ng; (void)V; (void)tt;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
      mbar_wait_p<PROF>(&d_full[0], dph, L.err, 44, c_df);
      mbar_wait_p<PROF>(&a2_empty[buf], bph ^ 1, L.err, 45, c_ae);
      tc_fence_after();
      uint8_t* a2b = a2_st + (size_t)buf * Cfg::A2_BUF;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[2][16];
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const int it = 2 * half + sub;
          const int mt = it / (N / 16), c0 = (it - mt * (N / 16)) * 16;
          uint32_t ax[16];
          tmem_ld16(taddr + mt * 2 * N + c0, r[sub]);
          tmem_ld16(taddr + mt * 2 * N + N + c0, ax);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[sub][i] = __float_as_uint(__uint_as_float(r[sub][i]) + __uint_as_float(ax[i]));
        }
        if (half == 1) {
          tc_fence_before();
          if constexpr (PAIR) mbar_arrive_rank<0>(&d_empty[0], 0); else mbar_arrive(&d_empty[0]);   // D1 is in registers: conv1 of the next tile may start
        }
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const int it = 2 * half + sub;                       // 16-column piece: accumulator columns [16 it, 16 it + 16)
          const int mt = it / (N / 16), c0 = (it - mt * (N / 16)) * 16;
          const int i2 = mt * 128 + warp * 32 + lane;          // conv2 operand row of this thread
          const int s_ = o0 - h2 + i2;                         // global conv1 output row
          const bool live = s_ >= 0 && s_ < valid;
#pragma unroll
          for (int g2 = 0; g2 < 2; ++g2) {                     // two 8-channel pieces = the two k-halves of chunk c0/16
            const float4 ba = __ldg(reinterpret_cast<const float4*>(P.b1 + c0 + g2 * 8));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(P.b1 + c0 + g2 * 8 + 4));
            float4 va, vb;
            va.x = __uint_as_float(r[sub][g2 * 8 + 0]) + ba.x; va.y = __uint_as_float(r[sub][g2 * 8 + 1]) + ba.y;
            va.z = __uint_as_float(r[sub][g2 * 8 + 2]) + ba.z; va.w = __uint_as_float(r[sub][g2 * 8 + 3]) + ba.w;
            vb.x = __uint_as_float(r[sub][g2 * 8 + 4]) + bb.x; vb.y = __uint_as_float(r[sub][g2 * 8 + 5]) + bb.y;
            vb.z = __uint_as_float(r[sub][g2 * 8 + 6]) + bb.z; vb.w = __uint_as_float(r[sub][g2 * 8 + 7]) + bb.w;
            va.x = live ? lrelu(va.x, slope) : 0.f; va.y = live ? lrelu(va.y, slope) : 0.f;
            va.z = live ? lrelu(va.z, slope) : 0.f; va.w = live ? lrelu(va.w, slope) : 0.f;
            vb.x = live ? lrelu(vb.x, slope) : 0.f; vb.y = live ? lrelu(vb.y, slope) : 0.f;
            vb.z = live ? lrelu(vb.z, slope) : 0.f; vb.w = live ? lrelu(vb.w, slope) : 0.f;
            uint2 ha, la, hb, lb;
            split4(va, ha, la);
            split4(vb, hb, lb);
            uint8_t* dst = a2b + (size_t)(c0 >> 4) * Cfg::A2_CHUNK + ((size_t)g2 * RA2 + i2) * 16;
            *reinterpret_cast<uint4*>(dst) = make_uint4(ha.x, ha.y, hb.x, hb.y);
            *reinterpret_cast<uint4*>(dst + (size_t)2 * RA2 * 16) = make_uint4(la.x, la.y, lb.x, lb.y);
          }
        }
      }
      fence_proxy_async();                                     // conv2 operand: generic-proxy stores -> tensor core
      if constexpr (PAIR) mbar_arrive_rank<0>(&a2_full[buf], 0); else mbar_arrive(&a2_full[buf]);
      dph ^= 1;
      if (++buf == 2) { buf = 0; bph ^= 1; }
    P2_TILE_END
    if (PROF && L.dbg && tid == 0 && blockIdx.x < 128) {
      long long* d = L.dbg + (size_t)blockIdx.x * 32 + 22;              // 22..24 E1: total, wait D1, wait operand buffer free
      d[0] = clock64() - t_begin; d[1] = c_df; d[2] = c_ae;
    }
  } else {
    // ============================ E2 (warps 4-11): D2 -> + b2 + x -> y, coalesced through a per-warp slab ============================
    // Two warps per TMEM lane quadrant, each owns two of the four 16-column pieces; the pieces are copied to registers and D2
    // released before the (global-memory bound) residual add and store.
    const int ew = (warp - W_E2) & 3, eh = (warp - W_E2) >> 2;
    uint32_t dph = 0;
    uint8_t* slab = epi_st + (warp - W_E2) * (32 * Cfg::EPI_PITCH);
    const int trow = lane >> 2, tch = lane & 3;
    long long c_df = 0;
    const long long t_begin = PROF ? clock64() : 0;
    P2_TILE_BEGIN
      (void)h1; (void)h2; (void)dil; (void)ng; (void)tt;
      const size_t base = (size_t)b * L.T_rows * N;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + Cfg::ACC;
      float4 rs[2][4];
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {                      // residual rows of both pieces: in flight while conv2 still runs
        const int it = 2 * eh + sub;
        const int mt = it / (N / 16), c0 = (it - mt * (N / 16)) * 16;
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          const int rl = mt * 128 + ew * 32 + s4 * 8 + trow;
          const int tau = o0 + rl;
          rs[sub][s4] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rl < V && tau < valid) rs[sub][s4] = __ldg(reinterpret_cast<const float4*>(P.x + base + (size_t)tau * N + c0 + tch * 4));
        }
      }
      mbar_wait_p<PROF>(&d_full[1], dph, L.err, 46, c_df);
      tc_fence_after();
      uint32_t r[2][16];
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const int it = 2 * eh + sub;
        const int mt = it / (N / 16), c0 = (it - mt * (N / 16)) * 16;
        uint32_t ax[16];
        tmem_ld16(taddr + mt * 2 * N + c0, r[sub]);
        tmem_ld16(taddr + mt * 2 * N + N + c0, ax);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) r[sub][i] = __float_as_uint(__uint_as_float(r[sub][i]) + __uint_as_float(ax[i]));
      }
      tc_fence_before();
      if constexpr (PAIR) mbar_arrive_rank<0>(&d_empty[1], 0); else mbar_arrive(&d_empty[1]);   // D2 is in registers: conv2 of the next tile may start
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const int it = 2 * eh + sub;
        const int mt = it / (N / 16), c0 = (it - mt * (N / 16)) * 16;
        const float4 bi = __ldg(reinterpret_cast<const float4*>(P.b2 + c0 + tch * 4));
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
          *reinterpret_cast<uint4*>(slab + lane * Cfg::EPI_PITCH + qq * 16) = make_uint4(r[sub][qq * 4], r[sub][qq * 4 + 1], r[sub][qq * 4 + 2], r[sub][qq * 4 + 3]);
        __syncwarp();
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          const int rl = mt * 128 + ew * 32 + s4 * 8 + trow;
          const int tau = o0 + rl;
          const float4 a = *reinterpret_cast<const float4*>(slab + (s4 * 8 + trow) * Cfg::EPI_PITCH + tch * 16);
          float4 o;
          o.x = (a.x + bi.x) + rs[sub][s4].x; o.y = (a.y + bi.y) + rs[sub][s4].y;
          o.z = (a.z + bi.z) + rs[sub][s4].z; o.w = (a.w + bi.w) + rs[sub][s4].w;
          if (rl < V && tau < valid) *reinterpret_cast<float4*>(P.out + base + (size_t)tau * N + c0 + tch * 4) = o;
        }
        __syncwarp();
      }
      dph ^= 1;
    P2_TILE_END
    if (PROF && L.dbg && warp == W_E2 && lane == 0 && blockIdx.x < 128) {
      long long* d = L.dbg + (size_t)blockIdx.x * 32 + 25;              // 25..26 E2: total, wait D2
      d[0] = clock64() - t_begin; d[1] = c_df;
    }
  }
#undef P2_TILE_BEGIN
#undef P2_TILE_END

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();      // nobody leaves while the peer may still signal into this CTA's shared memory
  if (warp == W_ISSA) {
    tc_fence_after();
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// fp32 Haiku conv weight w[k][C][C] -> CTA-pair layout: [chunk c = C/16][rank][tap j][k-half][1.5 C rows][8 bf16]
//   rank 0 rows: W_hi[0, C)  then W_hi[0, C/2)        (its halves of [W_hi | W_lo] and of W_hi)
//   rank 1 rows: W_lo[0, C)  then W_hi[C/2, C)
__global__ void pack_w_pc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int k, int C) {
  const int rows = C + C / 2;
  const size_t total = (size_t)(C / 16) * 2 * k * 2 * rows * 8;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r_ = idx;
    const int e = r_ % 8; r_ /= 8;
    const int q = r_ % rows; r_ /= rows;
    const int kh = r_ % 2; r_ /= 2;
    const int j = r_ % k; r_ /= k;
    const int rank = r_ % 2; r_ /= 2;
    const int c = (int)r_;
    int n, lo;
    if (q < C) { n = q; lo = rank; }
    else { n = (q - C) + rank * (C / 2); lo = 0; }
    const int i = c * 16 + kh * 8 + e;
    const float v = w[((size_t)j * C + i) * C + n];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    dst[idx] = lo ? __float2bfloat16_rn(v - __bfloat162float(hi)) : hi;
  }
}

template <int N, int PAIR = 0>
int launch_pair2(vtts_ctx* ctx, TcPairLaunch& L, cudaStream_t st) {
  using Cfg = P2Cfg<N, PAIR>;
  // function attributes and cluster occupancy are per device (a process may hold contexts on several GPUs)
  static bool attr_done_dev[64] = {};
  static int max_pairs_dev[64] = {};
  bool& attr_done = attr_done_dev[ctx->device & 63];
  int& max_pairs = max_pairs_dev[ctx->device & 63];
  if (!attr_done) {
    VTTS_CUDA(cudaFuncSetAttribute(tc_pair2_kernel<N, false, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    VTTS_CUDA(cudaFuncSetAttribute(tc_pair2_kernel<N, true, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    if (PAIR) {
      cudaLaunchConfig_t qc = {};
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      qc.gridDim = dim3(ctx->sm_count & ~1); qc.blockDim = dim3(NTHREADS); qc.dynamicSmemBytes = Cfg::SMEM_BYTES; qc.attrs = qa; qc.numAttrs = 1;
      VTTS_CUDA(cudaOccupancyMaxActiveClusters(&max_pairs, tc_pair2_kernel<N, false, PAIR>, &qc));
      if (max_pairs < 1) return ctx->fail(VTTS_ERR_CUDA, "tc_pair2: no CTA pair fits on this device");
      if (max_pairs > ctx->sm_count / 2) max_pairs = ctx->sm_count / 2;
    }
    attr_done = true;
  }
  if (PAIR)
    for (int i = 0; i < L.nprob; ++i)
      if (!L.p[i].w1pc || !L.p[i].w2pc) return ctx->fail(VTTS_ERR_BAD_ARG, "tc_pair2: the CTA-pair form needs the pair-layout weights");
  // expensive problems (large k) first: the last, partial wave of tiles is made of cheap ones
  std::stable_sort(L.p, L.p + L.nprob, [](const TcPairProb& a, const TcPairProb& b) { return a.k > b.k; });
  int most = 0;
  for (int i = 0; i < 3; ++i) {
    L.tile_start[i] = 0;
    L.tiles_per_row[i] = 1;
    if (i < L.nprob) {
      const int V = Cfg::R - (L.p[i].k - 1);
      L.tiles_per_row[i] = (L.T_rows + V - 1) / V;
      most = std::max(most, L.tiles_per_row[i] * L.B);
    }
  }
  if (PAIR) most = (most + 1) / 2;       // a pair tile = two consecutive tiles of one problem
  const int total = most * L.nprob;      // round-robin index space (surplus indices of the problems with fewer tiles are skipped)
  L.ntiles = total;
  if (PAIR) {
    cudaLaunchConfig_t lc = {};
    cudaLaunchAttribute la[1];
    la[0].id = cudaLaunchAttributeClusterDimension;
    la[0].val.clusterDim.x = 2; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
    const int pairs = total < max_pairs ? total : max_pairs;
    lc.gridDim = dim3(2 * pairs); lc.blockDim = dim3(NTHREADS); lc.dynamicSmemBytes = Cfg::SMEM_BYTES; lc.stream = st; lc.attrs = la; lc.numAttrs = 1;
    if (L.dbg) VTTS_CUDA(cudaLaunchKernelEx(&lc, tc_pair2_kernel<N, true, PAIR>, L));
    else VTTS_CUDA(cudaLaunchKernelEx(&lc, tc_pair2_kernel<N, false, PAIR>, L));
  } else {
    const int grid = total < ctx->sm_count ? total : ctx->sm_count;
    if (L.dbg) tc_pair2_kernel<N, true, PAIR><<<grid, NTHREADS, Cfg::SMEM_BYTES, st>>>(L);
    else tc_pair2_kernel<N, false, PAIR><<<grid, NTHREADS, Cfg::SMEM_BYTES, st>>>(L);
  }
  ctx->launches++;
  VTTS_CUDA(cudaGetLastError());
  return VTTS_OK;
}

}  // namespace

size_t vtts_tc_packed_pc_bytes(int k, int C) { return (size_t)k * C * C * 6; }   // 1.5 x the plain hi/lo packing

int vtts_tc_pack_weights_pc(vtts_ctx* ctx, const float* w, void* dst, int k, int C) {
  if (C % 16 != 0) return ctx->fail(VTTS_ERR_BAD_ARG, "tc_pair2: C %d", C);
  pack_w_pc_kernel<<<256, 256>>>(w, reinterpret_cast<__nv_bfloat16*>(dst), k, C);
  VTTS_CUDA(cudaGetLastError());
  return VTTS_OK;
}

static int launch_pair2_any(vtts_ctx* ctx, TcPairLaunch& L, cudaStream_t st, int pair);

int vtts_launch_tc_pair2(vtts_ctx* ctx, TcPairLaunch& L, cudaStream_t st) { return launch_pair2_any(ctx, L, st, 0); }
int vtts_launch_tc_pair2c(vtts_ctx* ctx, TcPairLaunch& L, cudaStream_t st) { return launch_pair2_any(ctx, L, st, 1); }

static int launch_pair2_any(vtts_ctx* ctx, TcPairLaunch& L, cudaStream_t st, int pair) {
  if (L.nprob < 1 || L.nprob > 3) return ctx->fail(VTTS_ERR_BAD_ARG, "tc_pair2: nprob %d", L.nprob);
  for (int i = 0; i < L.nprob; ++i) {
    const TcPairProb& p = L.p[i];
    if (p.k < 1 || (p.k & 1) == 0 || (p.k - 1) * p.dil > 50 || p.k - 1 > 15) return ctx->fail(VTTS_ERR_BAD_ARG, "tc_pair2: k=%d dil=%d", p.k, p.dil);
    if (p.x == p.out) return ctx->fail(VTTS_ERR_BAD_ARG, "tc_pair2: in-place not supported (tiles read halo rows of x)");
  }
  L.err = ctx->d_err;
  L.dbg = ctx->tc_dbg_on ? ctx->d_tc_dbg : nullptr;
  switch (L.N) {
    case 64: return pair ? launch_pair2<64, 1>(ctx, L, st) : launch_pair2<64, 0>(ctx, L, st);
    case 32: return pair ? launch_pair2<32, 1>(ctx, L, st) : launch_pair2<32, 0>(ctx, L, st);
    default: return ctx->fail(VTTS_ERR_BAD_ARG, "tc_pair2: N %d unsupported", L.N);
  }
}
