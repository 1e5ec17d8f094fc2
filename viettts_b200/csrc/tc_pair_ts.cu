// Fused ResBlock1 pair on tcgen05 with the A operand in TENSOR MEMORY ("TS" form of tcgen05.mma):
//     y = conv2(lrelu(conv1(lrelu(x)) + b1)) + b2 + x      C = N in {32, 64}
// (one iteration of the loop at vietTTS/hifigan/model.py:44-51: dilated conv, conv, residual add).
//
// Why TS.  Measured on B200 (profiles/r2_umma_probe2.txt): with the activations as a shared-memory A operand an
// M=128, K=16 MMA costs ~48-57 clk whatever N is (operand fetch), so the C <= 64 layers run the tensor pipe at 1/3 to
// 1/2 of its rate; with A in TMEM the same MMA runs at its math floor (N/2 clk), provided several warps issue.  The
// price is that a tap shift can no longer be a descriptor offset (TMEM lane = row): every (chunk, tap, M-tile) needs
// its own shifted copy of the activation tile in TMEM.  "Replicator" warps make those copies: LDS.128 from the staged
// bf16 hi/lo operand (row r + tap*dilation) -> tcgen05.st into a ring of 16-column A slots.
//
// Two decoupled pipelines share the tensor pipe, each with its own issuing warp, slot ring, weight ring, accumulators:
//   conv1:  converters (global fp32 -> lrelu -> hi/lo bf16, smem)  -> replicators-1 -> issuer A -> D1
//   conv2:  E1 warps (D1 -> +b1, lrelu, zero padding, hi/lo, smem) -> replicators-2 -> issuer B -> D2 -> E2 warps (+b2 +x -> y)
// so conv1 of tile i+1 overlaps conv2 of tile i and both epilogues.  Arithmetic: bf16x3 exactly as tc_conv.cu
// (a_hi.w_hi + a_hi.w_lo + a_lo.w_hi, fp32 accumulation in TMEM, same summation order as tc_pair.cu).
//
// Tile: conv1 produces R = 128*MT rows [o0-h2, o0-h2+R); conv2 consumes them and yields V = R-(k-1) valid rows.
// N=64: MT=1, slot groups of 4 taps;  N=32: MT=2, groups of 2 taps  (4 slots = 12 MMAs per issue region either way).
// TMEM (512 columns): [0,256) D1[2], D2[2] (MT*N columns each);  [256,384) ring 1;  [384,512) ring 2  (2 groups x 4 slots x 16).
#include <cuda_bf16.h>

#include <algorithm>

#include "tc_common.cuh"
#include "vtts_internal.cuh"

namespace {

using namespace tcx;

constexpr int NTHREADS = 896;      // 28 warps
// 0-3 E1 (D1 -> conv2 operand) | 4-11 E2 (D2 -> y, two warps per TMEM lane quadrant) | 12,13 MMA issuers | 14,15 weight
// producers | 16-19 converters | 20-23 replicators of ring 1 | 24-27 replicators of ring 2   (warp % 4 = TMEM lane quadrant)
constexpr int W_E2 = 4, W_ISSA = 12, W_ISSB = 13, W_WP1 = 14, W_WP2 = 15, W_CONV = 16, W_REP1 = 20, W_REP2 = 24;
constexpr int NCONV = 128;         // converter threads
constexpr int NA = 4;              // conv1 activation stages (one 16-channel chunk each)
constexpr int NSG = 3;             // slot groups per ring: tcgen05.wait::st of group i+1 completes only after the MMAs queued
                                   // before it, so two groups leave the tensor pipe idle between items (measured: 62 % busy)

template <int N>
struct TsCfg {
  static constexpr int MT = N == 64 ? 1 : 2;
  static constexpr int R = 128 * MT;
  static constexpr int G = 4 / MT;                   // taps per slot group
  static constexpr int NCH = N / 16;
  static constexpr int RA1 = R + 64;                 // conv1 operand rows per stage (halo <= 50)
  static constexpr int RA2 = R + 16;                 // conv2 operand rows (halo <= 10)
  static constexpr int A1_STAGE = RA1 * 64;
  static constexpr int A2_CHUNK = RA2 * 64;
  static constexpr int A2_BUF = NCH * A2_CHUNK;
  static constexpr int W_STAGE = N * 64;             // one (chunk, tap): [k-half][hi|lo][n][8 bf16]
  static constexpr int W_GROUP = G * W_STAGE;
  static constexpr int NWG = N == 64 ? 2 : 4;        // weight groups in flight per ring
  static constexpr int EPI_PITCH = 80;               // 16 floats + 16 B pad
  static constexpr int EPI_STAGE = 8 * 32 * EPI_PITCH;
  static constexpr int NBAR = 2 * NA + 4 * NSG + 4 * NWG + 4 + 4;
  static constexpr int SMEM_BYTES = NA * A1_STAGE + 2 * A2_BUF + 2 * NWG * W_GROUP + EPI_STAGE + NBAR * 8 + 16 + 1024;
  static constexpr int ACC = MT * N;                 // columns of one accumulator (= 64)
  static constexpr int RING_COLS = NSG * 64;         // 3 groups x 4 slots x 16 columns
  static_assert(2 * ACC + 2 * RING_COLS <= 512, "TMEM");
  static constexpr int NIT = MT * (N / 16);          // 16-column pieces of one accumulator (= 4)
};

// PROF = true adds the stalled cycles to a role counter (vtts_debug_tc_stats); the production instantiation keeps no
// counters at all: at 896 threads the kernel has 72 registers per thread and every 64-bit counter costs two of them
// (with the counters compiled in, the replicator loop spilled: 40 M local-memory instructions per launch, ncu).
template <bool PROF>
__device__ __forceinline__ void mbar_wait_c(uint64_t* bar, uint32_t parity, int* err, int code, long long& acc) {
  if constexpr (PROF) {
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
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint4& a, const uint4& b, const uint4& c, const uint4& d) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w), "r"(d.x),
      "r"(d.y), "r"(d.z), "r"(d.w)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int N, bool PROF>
__global__ void __launch_bounds__(NTHREADS, 1) tc_pair_ts_kernel(const __grid_constant__ TcPairLaunch L) {
  using Cfg = TsCfg<N>;
  constexpr int MT = Cfg::MT, R = Cfg::R, G = Cfg::G, NCH = Cfg::NCH, RA1 = Cfg::RA1, RA2 = Cfg::RA2, NWG = Cfg::NWG, NIT = Cfg::NIT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a1_st = smem;
  uint8_t* a2_st = a1_st + NA * Cfg::A1_STAGE;
  uint8_t* w_st = a2_st + 2 * Cfg::A2_BUF;                       // [ring][NWG][W_GROUP]
  uint8_t* epi_st = w_st + 2 * NWG * Cfg::W_GROUP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_st + Cfg::EPI_STAGE);
  uint64_t* a1_full = bars;                  // [NA]   converters -> replicators-1
  uint64_t* a1_empty = a1_full + NA;         // [NA]
  uint64_t* s_full = a1_empty + NA;          // [ring][NSG]  replicators -> issuer
  uint64_t* s_empty = s_full + 2 * NSG;      // [ring][NSG]
  uint64_t* w_full = s_empty + 2 * NSG;      // [ring][NWG]
  uint64_t* w_empty = w_full + 2 * NWG;      // [ring][NWG]
  uint64_t* d_full = w_empty + 2 * NWG;      // [conv]  issuer -> epilogue (accumulators are single buffered: the epilogue
  uint64_t* d_empty = d_full + 2;            // [conv]  warps copy them to registers and release them at once)
  uint64_t* a2_full = d_empty + 2;           // [2]    E1 -> replicators-2
  uint64_t* a2_empty = a2_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a2_empty + 2);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (warp == W_WP1 && lane == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&a1_full[i], NCONV); mbar_init(&a1_empty[i], 4); }
    for (int i = 0; i < 2 * NSG; ++i) { mbar_init(&s_full[i], 4); mbar_init(&s_empty[i], 1); }
    for (int i = 0; i < 2 * NWG; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(&d_full[0], 1); mbar_init(&d_full[1], 1);
    mbar_init(&d_empty[0], 128); mbar_init(&d_empty[1], 256);
    for (int i = 0; i < 2; ++i) { mbar_init(&a2_full[i], 128); mbar_init(&a2_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_ISSA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t ring_col0 = 2 * Cfg::ACC;                       // first slot column

  const int ntiles = L.ntiles;
#define TS_TILE_BEGIN                                                                 \
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {                     \
    const int pi = tile >= L.tile_start[2] ? 2 : (tile >= L.tile_start[1] ? 1 : 0);   \
    const TcPairProb& P = L.p[pi];                                                    \
    const int rest = tile - L.tile_start[pi];                                         \
    const int tpr = L.tiles_per_row[pi];                                              \
    const int tt = rest % tpr;                                                        \
    const int b = rest / tpr;                                                         \
    const int k = P.k, dil = P.dil;                                                   \
    const int V = R - (k - 1);                                                        \
    const int o0 = tt * V;                                                            \
    int valid = L.T_rows;                                                             \
    if (L.len) {                                                                      \
      const int v_ = L.len[b] * L.len_mul;                                            \
      valid = v_ < valid ? v_ : valid;                                                \
    }                                                                                 \
    if (o0 >= valid) continue;                                                        \
    const int h2 = (k - 1) / 2, h1 = ((k - 1) * dil) / 2;                             \
    const int ng = (k + G - 1) / G;
#define TS_TILE_END }

  if (warp == W_ISSA || warp == W_ISSB) {
    // ============================ MMA issuers: ring 0 = conv1, ring 1 = conv2 ============================
    const int ring = warp == W_ISSA ? 0 : 1;
    constexpr uint32_t idesc = make_idesc(N);
    const uint64_t b_tmpl = make_desc(0, 2 * N * 16, 128);          // [k-half][hi|lo][n][8]: k-half blocks 2N rows apart
    const uint32_t w_ring_u32 = smem_u32(w_st + (size_t)ring * NWG * Cfg::W_GROUP);
    const uint32_t slot_base = tmem_base + ring_col0 + ring * Cfg::RING_COLS;
    const uint32_t d0 = tmem_base + ring * Cfg::ACC;
    uint32_t dph = 0, sg = 0, sph = 0, ws = 0, wph = 0;
    long long c_d = 0, c_s = 0, c_w = 0;
    const long long t_begin = PROF ? clock64() : 0;
    TS_TILE_BEGIN
      (void)b; (void)h1; (void)h2; (void)o0; (void)dil;
      mbar_wait_c<PROF>(&d_empty[ring], dph ^ 1, L.err, 30 + ring, c_d);
      tc_fence_after();
      for (int c = 0; c < NCH; ++c) {
        for (int g = 0; g < ng; ++g) {
          const int nt = (k - g * G) < G ? (k - g * G) : G;
          mbar_wait_c<PROF>(&w_full[ring * NWG + ws], wph, L.err, 34 + ring, c_w);     // weights arrive early: off the critical path
          mbar_wait_c<PROF>(&s_full[ring * NSG + sg], sph, L.err, 32 + ring, c_s);
          tc_fence_after();
          const uint32_t w_base16 = (w_ring_u32 + ws * Cfg::W_GROUP) >> 4;
          const uint32_t sl0 = slot_base + sg * 64;
          const uint32_t first_grp = (c | g) != 0 ? 1u : 0u;
          if (elect_one()) {
            for (int t = 0; t < nt; ++t) {
              const uint64_t b_hi = b_tmpl | (uint64_t)(w_base16 + t * (Cfg::W_STAGE >> 4));
              const uint64_t b_lo = b_tmpl | (uint64_t)(w_base16 + t * (Cfg::W_STAGE >> 4) + N);
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint32_t sl = sl0 + (t * MT + mt) * 16;
                const uint32_t d = d0 + mt * N;
                umma_ts(d, sl, b_hi, idesc, (first_grp | (uint32_t)t) != 0 ? 1u : 0u);
                umma_ts(d, sl, b_lo, idesc, 1u);
                umma_ts(d, sl + 8, b_hi, idesc, 1u);
              }
            }
            umma_commit(&s_empty[ring * NSG + sg]);
            umma_commit(&w_empty[ring * NWG + ws]);
          }
          __syncwarp();
          if (++sg == NSG) { sg = 0; sph ^= 1; }
          if (++ws == NWG) { ws = 0; wph ^= 1; }
        }
      }
      if (elect_one()) umma_commit(&d_full[ring]);
      __syncwarp();
      dph ^= 1;
    TS_TILE_END
    if (PROF && L.dbg && lane == 0 && blockIdx.x < 128) {
      long long* d = L.dbg + (size_t)blockIdx.x * 32 + ring * 4;      // 0..3 issuer A, 4..7 issuer B: total, wait acc, wait slots, wait weights
      d[0] = clock64() - t_begin; d[1] = c_d; d[2] = c_s; d[3] = c_w;
    }
  } else if (warp == W_WP1 || warp == W_WP2) {
    // ============================ weight producers (one per ring) ============================
    if (lane == 0) {
      const int ring = warp == W_WP1 ? 0 : 1;
      uint8_t* wr = w_st + (size_t)ring * NWG * Cfg::W_GROUP;
      uint32_t ws = 0, wph = 0;
      long long c_e = 0;
      TS_TILE_BEGIN
        (void)b; (void)h1; (void)h2; (void)o0; (void)dil;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(ring == 0 ? P.w1pk : P.w2pk);
        for (int c = 0; c < NCH; ++c)
          for (int g = 0; g < ng; ++g) {
            const int nt = (k - g * G) < G ? (k - g * G) : G;
            const uint32_t bytes = (uint32_t)nt * Cfg::W_STAGE;
            mbar_wait_c<PROF>(&w_empty[ring * NWG + ws], wph ^ 1, L.err, 36 + ring, c_e);
            mbar_expect_tx(&w_full[ring * NWG + ws], bytes);
            bulk_g2s(wr + ws * Cfg::W_GROUP, wsrc + ((size_t)c * k + (size_t)g * G) * Cfg::W_STAGE, bytes, &w_full[ring * NWG + ws]);
            if (++ws == NWG) { ws = 0; wph ^= 1; }
          }
      TS_TILE_END
    }
    __syncwarp();
  } else if (warp >= W_CONV && warp < W_REP1) {
    // ============================ activation converters (conv1 input) ============================
    const int gt = tid - W_CONV * 32;      // 0..127
    const int q = gt & 3;
    const int r0 = gt >> 2;
    const float slope = L.slope;
    uint32_t item = 0;
    long long c_e = 0;
    const long long t_begin = PROF ? clock64() : 0;
    TS_TILE_BEGIN
      (void)ng;
      const int rows = R + (k - 1) * dil;
      const float* x0 = P.x + (size_t)b * L.T_rows * N;
      const int row_base = o0 - h2 - h1;
      for (int c = 0; c < NCH; ++c, ++item) {
        const uint32_t sa = item % NA, pa = (item / NA) & 1;
        mbar_wait_c<PROF>(&a1_empty[sa], pa ^ 1, L.err, 38, c_e);
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
            if (rr < rows && t >= 0 && t < valid) v[u] = __ldg(reinterpret_cast<const float4*>(x0 + (size_t)t * N + coff));
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
        mbar_arrive(&a1_full[sa]);
      }
    TS_TILE_END
    if (PROF && L.dbg && gt == 0 && blockIdx.x < 128) {
      long long* d = L.dbg + (size_t)blockIdx.x * 32 + 8;               // 8..9 converters: total, wait stage free
      d[0] = clock64() - t_begin; d[1] = c_e;
    }
  } else if (warp >= W_REP1) {
    // ============================ replicators: smem operand rows (shifted per tap) -> TMEM A slots ============================
    const int ring = warp >= W_REP2 ? 1 : 0;
    const int q = warp & 3;                                   // TMEM lane quadrant of this warp
    const uint32_t slot_base = tmem_base + ((uint32_t)(q * 32) << 16) + ring_col0 + ring * Cfg::RING_COLS;
    uint32_t sg = 0, sph = 0, item = 0, buf = 0, bph = 0;
    long long c_in = 0, c_se = 0, c_fill = 0, c_ws = 0;
    const long long t_begin = PROF ? clock64() : 0;
    TS_TILE_BEGIN
      (void)b; (void)h1; (void)h2; (void)o0;
      const int dl = ring == 0 ? dil : 1;
      const int RA = ring == 0 ? RA1 : RA2;
      if (ring == 1) mbar_wait_c<PROF>(&a2_full[buf], bph, L.err, 40, c_in);
      for (int c = 0; c < NCH; ++c) {
        const uint8_t* src;
        uint32_t sa = 0;
        if (ring == 0) {
          sa = item % NA;
          mbar_wait_c<PROF>(&a1_full[sa], (item / NA) & 1, L.err, 41, c_in);
          src = a1_st + sa * Cfg::A1_STAGE;
          ++item;
        } else {
          src = a2_st + (size_t)buf * Cfg::A2_BUF + (size_t)c * Cfg::A2_CHUNK;
        }
        const uint32_t rowp = smem_u32(src) + (uint32_t)(q * 32 + lane) * 16u;
        const uint32_t ra16 = (uint32_t)RA * 16u;
        for (int g = 0; g < ng; ++g) {
          const int nslots = ((k - g * G) < G ? (k - g * G) : G) * MT;      // slot s = tap_local * MT + mt
          mbar_wait_c<PROF>(&s_empty[ring * NSG + sg], sph ^ 1, L.err, 42 + ring, c_se);
          tc_fence_after();
          const long long tf0 = PROF ? clock64() : 0;
          const uint32_t st0 = slot_base + sg * 64;
          // software pipelined over the slots: the four 16 B loads of slot s+1 are in flight while slot s is stored
          uint4 a0, a1, a2, a3;
          {
            const uint32_t p = rowp + (uint32_t)((g * G) * dl) * 16u;       // slot 0: tap g*G, M tile 0
            a0 = lds128(p); a1 = lds128(p + ra16); a2 = lds128(p + 2 * ra16); a3 = lds128(p + 3 * ra16);
          }
#pragma unroll
          for (int sidx = 0; sidx < 4; ++sidx) {
            if (sidx < nslots) {
              uint4 b0 = a0, b1 = a1, b2 = a2, b3 = a3;
              if (sidx + 1 < nslots) {
                const int sn = sidx + 1;
                const uint32_t p = rowp + (uint32_t)((sn % MT) * 128 + (g * G + sn / MT) * dl) * 16u;
                a0 = lds128(p); a1 = lds128(p + ra16); a2 = lds128(p + 2 * ra16); a3 = lds128(p + 3 * ra16);
              }
              tmem_st16(st0 + sidx * 16, b0, b1, b2, b3);
            }
          }
          const long long tf1 = PROF ? clock64() : 0;
          tmem_st_wait();
          if constexpr (PROF) { c_fill += tf1 - tf0; c_ws += clock64() - tf1; }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_full[ring * NSG + sg]);
          if (++sg == NSG) { sg = 0; sph ^= 1; }
        }
        if (ring == 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&a1_empty[sa]);
        }
      }
      if (ring == 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&a2_empty[buf]);
        if (++buf == 2) { buf = 0; bph ^= 1; }
      }
    TS_TILE_END
    if (PROF && L.dbg && lane == 0 && q == 0 && blockIdx.x < 128) {
      long long* d = L.dbg + (size_t)blockIdx.x * 32 + 12 + ring * 5;   // 12..16 / 17..21: total, wait operand, wait slots free, lds+st issue, wait::st
      d[0] = clock64() - t_begin; d[1] = c_in; d[2] = c_se; d[3] = c_fill; d[4] = c_ws;
    }
  } else if (warp < W_E2) {
    // ============================ E1: D1 -> + b1, leaky_relu, zero padding, hi/lo split -> conv2 operand ============================
    // The accumulator is copied to registers in two halves; D1 is released after the second copy, before its conversion.
    uint32_t dph = 0, buf = 0, bph = 0;
    const float slope = L.slope;
    long long c_df = 0, c_ae = 0;
    const long long t_begin = PROF ? clock64() : 0;
    TS_TILE_BEGIN
      (void)b; (void)h1; (void)dil; (void)ng; (void)V;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
      mbar_wait_c<PROF>(&d_full[0], dph, L.err, 44, c_df);
      mbar_wait_c<PROF>(&a2_empty[buf], bph ^ 1, L.err, 45, c_ae);
      tc_fence_after();
      uint8_t* a2b = a2_st + (size_t)buf * Cfg::A2_BUF;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[2][16];
        tmem_ld16(taddr + (2 * half) * 16, r[0]);
        tmem_ld16(taddr + (2 * half + 1) * 16, r[1]);
        tmem_ld_wait();
        if (half == 1) {
          tc_fence_before();
          mbar_arrive(&d_empty[0]);                            // D1 is in registers: conv1 of the next tile may start
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
      mbar_arrive(&a2_full[buf]);
      dph ^= 1;
      if (++buf == 2) { buf = 0; bph ^= 1; }
    TS_TILE_END
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
    TS_TILE_BEGIN
      (void)h1; (void)h2; (void)dil; (void)ng;
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
      mbar_wait_c<PROF>(&d_full[1], dph, L.err, 46, c_df);
      tc_fence_after();
      uint32_t r[2][16];
      tmem_ld16(taddr + (2 * eh) * 16, r[0]);
      tmem_ld16(taddr + (2 * eh + 1) * 16, r[1]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&d_empty[1]);                                // D2 is in registers: conv2 of the next tile may start
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
    TS_TILE_END
    if (PROF && L.dbg && warp == W_E2 && lane == 0 && blockIdx.x < 128) {
      long long* d = L.dbg + (size_t)blockIdx.x * 32 + 25;              // 25..26 E2: total, wait D2
      d[0] = clock64() - t_begin; d[1] = c_df;
    }
  }
#undef TS_TILE_BEGIN
#undef TS_TILE_END

  tc_fence_before();
  __syncthreads();
  if (warp == W_ISSA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int N>
int launch_pair_ts(vtts_ctx* ctx, TcPairLaunch& L, cudaStream_t st) {
  using Cfg = TsCfg<N>;
  static bool attr_done_dev[64] = {};   // function attributes are per device
  bool& attr_done = attr_done_dev[ctx->device & 63];
  if (!attr_done) {
    VTTS_CUDA(cudaFuncSetAttribute(tc_pair_ts_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    VTTS_CUDA(cudaFuncSetAttribute(tc_pair_ts_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  // expensive problems (large k) first: the last, partial wave of tiles is made of cheap ones
  std::stable_sort(L.p, L.p + L.nprob, [](const TcPairProb& a, const TcPairProb& b) { return a.k > b.k; });
  int total = 0;
  for (int i = 0; i < 3; ++i) {
    L.tile_start[i] = total;
    if (i < L.nprob) {
      const int V = Cfg::R - (L.p[i].k - 1);
      L.tiles_per_row[i] = (L.T_rows + V - 1) / V;
      total += L.tiles_per_row[i] * L.B;
    } else {
      L.tiles_per_row[i] = 1;
      L.tile_start[i] = 0x7fffffff;   // never selected
    }
  }
  L.ntiles = total;
  const int grid = total < ctx->sm_count ? total : ctx->sm_count;
  if (L.dbg) tc_pair_ts_kernel<N, true><<<grid, NTHREADS, Cfg::SMEM_BYTES, st>>>(L);
  else tc_pair_ts_kernel<N, false><<<grid, NTHREADS, Cfg::SMEM_BYTES, st>>>(L);
  ctx->launches++;
  VTTS_CUDA(cudaGetLastError());
  return VTTS_OK;
}

}  // namespace

int vtts_launch_tc_pair_ts(vtts_ctx* ctx, TcPairLaunch& L, cudaStream_t st) {
  if (L.nprob < 1 || L.nprob > 3) return ctx->fail(VTTS_ERR_BAD_ARG, "tc_pair_ts: nprob %d", L.nprob);
  for (int i = 0; i < L.nprob; ++i) {
    const TcPairProb& p = L.p[i];
    if (p.k < 1 || (p.k & 1) == 0 || (p.k - 1) * p.dil > 50 || p.k - 1 > 15) return ctx->fail(VTTS_ERR_BAD_ARG, "tc_pair_ts: k=%d dil=%d", p.k, p.dil);
    if (p.x == p.out) return ctx->fail(VTTS_ERR_BAD_ARG, "tc_pair_ts: in-place not supported (tiles read halo rows of x)");
  }
  L.err = ctx->d_err;
  L.dbg = ctx->tc_dbg_on ? ctx->d_tc_dbg : nullptr;
  switch (L.N) {
    case 64: return launch_pair_ts<64>(ctx, L, st);
    case 32: return launch_pair_ts<32>(ctx, L, st);
    default: return ctx->fail(VTTS_ERR_BAD_ARG, "tc_pair_ts: N %d unsupported", L.N);
  }
}
