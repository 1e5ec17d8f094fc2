// Fused ResBlock1 pair on tcgen05:  y = conv2(lrelu(conv1(lrelu(x)) + b1)) + b2 + x   for C = N in {32, 64}
// (one iteration of the loop at vietTTS/hifigan/model.py:44-51: dilated conv, conv, residual add).
//
// Why: with separate conv launches the fp32 intermediate costs 8 of the 20 bytes of HBM traffic per
// element-pair, and the stages with C <= 64 are bandwidth bound.  Here the conv1 accumulator goes
// TMEM -> registers (bias, leaky_relu, zero padding at the sequence ends, bf16 hi/lo split) -> the shared
// memory A operand of conv2, so only x is read and y written.
//
// Tile: conv1 produces R = 256 rows [o0-h2, o0-h2+R) (two M=128 tiles), conv2 consumes them and yields
// V = R-(k-1) valid output rows [o0, o0+V); arithmetic is bf16x3 as in tc_conv.cu.  One persistent CTA per
// SM, 14 warps: 0-3 epilogue (phase 1: D1 -> A2 operand, phase 2: D2 -> global), 4 MMA issue, 5 weight
// producer, 6-13 activation converters (two groups).  D1/D2 accumulator pairs are double buffered in TMEM.
#include <cuda_bf16.h>

#include <algorithm>

#include "tc_common.cuh"
#include "vtts_internal.cuh"

namespace {

using namespace tcx;

constexpr int NA = 4;          // conv1 activation stages
constexpr int NW = 8;          // weight stages (2-4 KB each)
constexpr int NTHREADS = 448;
constexpr int NCONV = 256, NGRP = 2, GRP_THREADS = NCONV / NGRP;
constexpr int MT = 2, R = 128 * MT, RA = R + 64, RA2 = R + 16;

template <int N>
struct PairCfg {
  static constexpr int NCH = N / 16;
  static constexpr int A_STAGE = RA * 64;
  static constexpr int A2_CHUNK = RA2 * 64;          // one 16-channel chunk of the conv2 operand
  static constexpr int W_STAGE = N * 64;
  static constexpr int ACC_COLS = 2 * MT * N;        // D1 + D2 of one tile
  static constexpr int TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : (2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512)));
  static constexpr int EPI_PITCH = 144;
  static constexpr int EPI_STAGE = 4 * 32 * EPI_PITCH;
  static constexpr int NBAR = 2 * NA + 2 * NW + 4 + 2 + 4;   // a, w, d1 full/empty[2], a2 full/empty, d2 full/empty[2]
  static constexpr int SMEM_BYTES = NA * A_STAGE + NCH * A2_CHUNK + NW * W_STAGE + EPI_STAGE + NBAR * 8 + 16 + 1024;
  static_assert(2 * ACC_COLS <= 512, "TMEM");
};

template <int N>
__global__ void __launch_bounds__(NTHREADS, 1) tc_pair_kernel(const __grid_constant__ TcPairLaunch L) {
  using Cfg = PairCfg<N>;
  constexpr int NCH = Cfg::NCH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_st = smem;
  uint8_t* a2_st = a_st + NA * Cfg::A_STAGE;
  uint8_t* w_st = a2_st + NCH * Cfg::A2_CHUNK;
  uint8_t* epi_st = w_st + NW * Cfg::W_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_st + Cfg::EPI_STAGE);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + NA;
  uint64_t* w_full = a_empty + NA;
  uint64_t* w_empty = w_full + NW;
  uint64_t* d1_full = w_empty + NW;      // [2]
  uint64_t* d1_empty = d1_full + 2;      // [2]
  uint64_t* a2_full = d1_empty + 2;
  uint64_t* a2_empty = a2_full + 1;
  uint64_t* d2_full = a2_empty + 1;      // [2]
  uint64_t* d2_empty = d2_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_empty + 2);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (warp == 5 && lane == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], GRP_THREADS); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], 128);
      mbar_init(&d2_full[i], 1); mbar_init(&d2_empty[i], 128);
    }
    mbar_init(a2_full, 128);
    mbar_init(a2_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  const int ntiles = L.ntiles;
  // tile -> (problem, batch row, time tile); each problem has its own V = R-(k-1) and tile count
#define PAIR_TILE_BEGIN                                                               \
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
    const int h2 = (k - 1) / 2, h1 = ((k - 1) * dil) / 2;
#define PAIR_TILE_END }

  if (warp == 4) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc = make_idesc(N);
    uint32_t sw = 0, pw = 0, acc = 0, aph = 0, item = 0, a2ph = 0;
    const uint32_t a_st_u32 = smem_u32(a_st), w_st_u32 = smem_u32(w_st), a2_u32 = smem_u32(a2_st);
    const uint64_t a_tmpl = make_desc(0, RA * 16, 128);
    const uint64_t a2_tmpl = make_desc(0, RA2 * 16, 128);
    const uint64_t b_tmpl = make_desc(0, 2 * N * 16, 128);   // weight stage layout [k-half][hi|lo][n][8]
    long long w_d = 0, w_a = 0, w_w = 0, w_a2 = 0;
    const long long t_begin = clock64();
    PAIR_TILE_BEGIN
      (void)b; (void)h1; (void)h2; (void)o0;
      const uint32_t d1 = tmem_base + acc * Cfg::ACC_COLS;
      const uint32_t d2 = d1 + MT * N;
      // ---- conv1: D1 += A1 . W1 ----
      mbar_wait_t(&d1_empty[acc], aph ^ 1, L.err, 11, w_d);
      tc_fence_after();
      for (int c = 0; c < NCH; ++c, ++item) {
        const uint32_t sa = item % NA, pa = (item / NA) & 1;
        mbar_wait_t(&a_full[sa], pa, L.err, 12, w_a);
        tc_fence_after();
        const uint32_t a_base16 = (a_st_u32 + sa * Cfg::A_STAGE) >> 4;
        for (int j = 0; j < k; ++j) {
          mbar_wait_t(&w_full[sw], pw, L.err, 13, w_w);
          tc_fence_after();
          const uint32_t w_base16 = (w_st_u32 + sw * Cfg::W_STAGE) >> 4;
          const uint64_t b_hi = b_tmpl | (uint64_t)w_base16;
          const uint64_t b_lo = b_tmpl | (uint64_t)(w_base16 + N);
          const uint32_t first = (c | j) != 0 ? 1u : 0u;
          if (elect_one()) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint32_t row = a_base16 + mt * 128 + j * dil;
              const uint64_t a_hi = a_tmpl | (uint64_t)row;
              const uint64_t a_lo = a_tmpl | (uint64_t)(row + 2 * RA);
              umma(d1 + mt * N, a_hi, b_hi, idesc, first);
              umma(d1 + mt * N, a_hi, b_lo, idesc, 1u);
              umma(d1 + mt * N, a_lo, b_hi, idesc, 1u);
            }
            umma_commit(&w_empty[sw]);
          }
          if (++sw == NW) { sw = 0; pw ^= 1; }
        }
        if (elect_one()) umma_commit(&a_empty[sa]);
      }
      if (elect_one()) umma_commit(&d1_full[acc]);
      // ---- conv2: D2 += A2 . W2  (A2 written by the epilogue warps from D1) ----
      mbar_wait_t(a2_full, a2ph, L.err, 14, w_a2);
      mbar_wait_t(&d2_empty[acc], aph ^ 1, L.err, 15, w_d);
      tc_fence_after();
      for (int c = 0; c < NCH; ++c) {
        const uint32_t a2_base16 = (a2_u32 + c * Cfg::A2_CHUNK) >> 4;
        for (int j = 0; j < k; ++j) {
          mbar_wait_t(&w_full[sw], pw, L.err, 16, w_w);
          tc_fence_after();
          const uint32_t w_base16 = (w_st_u32 + sw * Cfg::W_STAGE) >> 4;
          const uint64_t b_hi = b_tmpl | (uint64_t)w_base16;
          const uint64_t b_lo = b_tmpl | (uint64_t)(w_base16 + N);
          const uint32_t first = (c | j) != 0 ? 1u : 0u;
          if (elect_one()) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint32_t row = a2_base16 + mt * 128 + j;
              const uint64_t a_hi = a2_tmpl | (uint64_t)row;
              const uint64_t a_lo = a2_tmpl | (uint64_t)(row + 2 * RA2);
              umma(d2 + mt * N, a_hi, b_hi, idesc, first);
              umma(d2 + mt * N, a_hi, b_lo, idesc, 1u);
              umma(d2 + mt * N, a_lo, b_hi, idesc, 1u);
            }
            umma_commit(&w_empty[sw]);
          }
          if (++sw == NW) { sw = 0; pw ^= 1; }
        }
      }
      if (elect_one()) {
        umma_commit(a2_empty);
        umma_commit(&d2_full[acc]);
      }
      a2ph ^= 1;
      if (++acc == 2) { acc = 0; aph ^= 1; }
    PAIR_TILE_END
    if (L.dbg && lane == 0) {
      long long* d = L.dbg + (size_t)blockIdx.x * 16;
      d[0] = clock64() - t_begin; d[1] = w_d; d[2] = w_a; d[3] = w_w; d[11] = w_a2;
    }
    __syncwarp();
  } else if (warp == 5) {
    // ============================ weight producer ============================
    if (lane == 0) {
      uint32_t sw = 0, pw = 0;
      long long w_e = 0;
      PAIR_TILE_BEGIN
        (void)b; (void)h1; (void)h2; (void)o0; (void)dil;
        for (int which = 0; which < 2; ++which) {
          const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(which == 0 ? P.w1pk : P.w2pk);
          for (int s = 0; s < NCH * k; ++s) {
            mbar_wait_t(&w_empty[sw], pw ^ 1, L.err, 17, w_e);
            mbar_expect_tx(&w_full[sw], Cfg::W_STAGE);
            bulk_g2s(w_st + sw * Cfg::W_STAGE, wsrc + (size_t)s * Cfg::W_STAGE, Cfg::W_STAGE, &w_full[sw]);
            if (++sw == NW) { sw = 0; pw ^= 1; }
          }
        }
      PAIR_TILE_END
      if (L.dbg) L.dbg[(size_t)blockIdx.x * 16 + 4] = w_e;
    }
    __syncwarp();
  } else if (warp >= 6) {
    // ============================ activation converters (conv1 input) ============================
    const int ct = tid - 192;
    const int grp = ct / GRP_THREADS;
    const int gt = ct - grp * GRP_THREADS;
    const int q = gt & 3;
    const int r0 = gt >> 2;
    const float slope = L.slope;
    uint32_t item = 0;
    long long w_ae = 0, t_fill = 0;
    PAIR_TILE_BEGIN
      const int rows = R + (k - 1) * dil;
      const float* x0 = P.x + (size_t)b * L.T_rows * N;
      const int row_base = o0 - h2 - h1;
      for (int c = 0; c < NCH; ++c, ++item) {
        if ((int)(item % NGRP) != grp) continue;
        const uint32_t sa = item % NA, pa = (item / NA) & 1;
        mbar_wait_t(&a_empty[sa], pa ^ 1, L.err, 18, w_ae);
        const long long tf0 = clock64();
        uint8_t* st = a_st + sa * Cfg::A_STAGE + ((q >> 1) * RA) * 16 + (q & 1) * 8;
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
              *reinterpret_cast<uint2*>(st + (size_t)(2 * RA + rr) * 16) = lo;
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(&a_full[sa]);
        t_fill += clock64() - tf0;
      }
    PAIR_TILE_END
    if (L.dbg && gt == 0) { L.dbg[(size_t)blockIdx.x * 16 + 5 + 4 * grp] = w_ae; L.dbg[(size_t)blockIdx.x * 16 + 6 + 4 * grp] = t_fill; }
  } else {
    // ============================ epilogue warps 0-3 ============================
    uint32_t acc = 0, aph = 0, a2ph = 0;
    long long w_tf = 0, t_epi = 0;
    uint8_t* slab = epi_st + warp * (32 * Cfg::EPI_PITCH);
    const int trow = lane >> 3, tch = lane & 7;
    constexpr int NCHUNK = N / 32;
    const float slope = L.slope;
    PAIR_TILE_BEGIN
      (void)h1; (void)dil;
      const size_t base = (size_t)b * L.T_rows * N;
      const uint32_t taddr1 = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * Cfg::ACC_COLS;
      const uint32_t taddr2 = taddr1 + MT * N;
      // ---------------- phase 1: D1 -> bias, leaky_relu, zero padding, hi/lo split -> A2 operand ----------------
      mbar_wait_t(&d1_full[acc], aph, L.err, 19, w_tf);
      mbar_wait_t(a2_empty, a2ph ^ 1, L.err, 20, w_tf);      // conv2 MMAs of the previous tile no longer read A2
      const long long te0 = clock64();
      tc_fence_after();
#pragma unroll 1
      for (int it = 0; it < MT * NCHUNK; ++it) {
        const int mt = it / NCHUNK, c0 = (it - mt * NCHUNK) * 32;
        uint32_t r[32];
        tmem_ld16(taddr1 + mt * N + c0, r);
        tmem_ld16(taddr1 + mt * N + c0 + 16, r + 16);
        const int i2 = mt * 128 + warp * 32 + lane;          // A2 row of this thread
        const int s = o0 - h2 + i2;                          // global conv1 output row
        const bool live = s >= 0 && s < valid;
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {                        // four 8-channel pieces of the 32 columns
          float4 va, vb;
          const float4 ba = __ldg(reinterpret_cast<const float4*>(P.b1 + c0 + g * 8));
          const float4 bb = __ldg(reinterpret_cast<const float4*>(P.b1 + c0 + g * 8 + 4));
          va.x = __uint_as_float(r[g * 8 + 0]) + ba.x; va.y = __uint_as_float(r[g * 8 + 1]) + ba.y;
          va.z = __uint_as_float(r[g * 8 + 2]) + ba.z; va.w = __uint_as_float(r[g * 8 + 3]) + ba.w;
          vb.x = __uint_as_float(r[g * 8 + 4]) + bb.x; vb.y = __uint_as_float(r[g * 8 + 5]) + bb.y;
          vb.z = __uint_as_float(r[g * 8 + 6]) + bb.z; vb.w = __uint_as_float(r[g * 8 + 7]) + bb.w;
          va.x = live ? lrelu(va.x, slope) : 0.f; va.y = live ? lrelu(va.y, slope) : 0.f;
          va.z = live ? lrelu(va.z, slope) : 0.f; va.w = live ? lrelu(va.w, slope) : 0.f;
          vb.x = live ? lrelu(vb.x, slope) : 0.f; vb.y = live ? lrelu(vb.y, slope) : 0.f;
          vb.z = live ? lrelu(vb.z, slope) : 0.f; vb.w = live ? lrelu(vb.w, slope) : 0.f;
          uint2 ha, la, hb, lb;
          split4(va, ha, la);
          split4(vb, hb, lb);
          const int ch = c0 + g * 8;                         // first channel of this piece
          const int chunk = ch >> 4, kh = (ch >> 3) & 1;
          uint8_t* dst = a2_st + chunk * Cfg::A2_CHUNK + ((size_t)kh * RA2 + i2) * 16;
          *reinterpret_cast<uint4*>(dst) = make_uint4(ha.x, ha.y, hb.x, hb.y);
          *reinterpret_cast<uint4*>(dst + (size_t)2 * RA2 * 16) = make_uint4(la.x, la.y, lb.x, lb.y);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&d1_empty[acc]);
      mbar_arrive(a2_full);
      // ---------------- phase 2: D2 -> + b2 + x -> y ----------------
      const int row_w = o0 + warp * 32;
      float4 rs[8];
      auto load_resid = [&](int it, float4 (&dstv)[8]) {
        const int mt = it / NCHUNK, c0 = (it - mt * NCHUNK) * 32;
#pragma unroll
        for (int s8 = 0; s8 < 8; ++s8) {
          const int rl = mt * 128 + warp * 32 + s8 * 4 + trow;
          const int tau = o0 + rl;
          dstv[s8] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rl < V && tau < valid) dstv[s8] = __ldg(reinterpret_cast<const float4*>(P.x + base + (size_t)tau * N + c0 + tch * 4));
        }
      };
      load_resid(0, rs);
      mbar_wait_t(&d2_full[acc], aph, L.err, 21, w_tf);
      tc_fence_after();
#pragma unroll 1
      for (int it = 0; it < MT * NCHUNK; ++it) {
        const int mt = it / NCHUNK, c0 = (it - mt * NCHUNK) * 32;
        uint32_t r[32];
        tmem_ld16(taddr2 + mt * N + c0, r);
        tmem_ld16(taddr2 + mt * N + c0 + 16, r + 16);
        float4 rs_next[8];
        if (it + 1 < MT * NCHUNK) load_resid(it + 1, rs_next);
        const float4 bi = __ldg(reinterpret_cast<const float4*>(P.b2 + c0 + tch * 4));
        tmem_ld_wait();
#pragma unroll
        for (int qq = 0; qq < 8; ++qq)
          *reinterpret_cast<uint4*>(slab + lane * Cfg::EPI_PITCH + qq * 16) = make_uint4(r[qq * 4], r[qq * 4 + 1], r[qq * 4 + 2], r[qq * 4 + 3]);
        __syncwarp();
#pragma unroll
        for (int s8 = 0; s8 < 8; ++s8) {
          const int rl = mt * 128 + warp * 32 + s8 * 4 + trow;
          const int tau = o0 + rl;
          const float4 a = *reinterpret_cast<const float4*>(slab + (s8 * 4 + trow) * Cfg::EPI_PITCH + tch * 16);
          float4 o;
          o.x = (a.x + bi.x) + rs[s8].x; o.y = (a.y + bi.y) + rs[s8].y;
          o.z = (a.z + bi.z) + rs[s8].z; o.w = (a.w + bi.w) + rs[s8].w;
          if (rl < V && tau < valid) *reinterpret_cast<float4*>(P.out + base + (size_t)tau * N + c0 + tch * 4) = o;
        }
        __syncwarp();
        if (it + 1 < MT * NCHUNK) {
#pragma unroll
          for (int s8 = 0; s8 < 8; ++s8) rs[s8] = rs_next[s8];
        }
      }
      (void)row_w;
      tc_fence_before();
      mbar_arrive(&d2_empty[acc]);
      t_epi += clock64() - te0;
      a2ph ^= 1;
      if (++acc == 2) { acc = 0; aph ^= 1; }
    PAIR_TILE_END
    if (L.dbg && tid == 0) { L.dbg[(size_t)blockIdx.x * 16 + 7] = w_tf; L.dbg[(size_t)blockIdx.x * 16 + 8] = t_epi; }
  }
#undef PAIR_TILE_BEGIN
#undef PAIR_TILE_END

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

template <int N>
int launch_pair(vtts_ctx* ctx, TcPairLaunch& L, cudaStream_t st) {
  using Cfg = PairCfg<N>;
  static bool attr_done_dev[64] = {};   // function attributes are per device
  bool& attr_done = attr_done_dev[ctx->device & 63];
  if (!attr_done) {
    VTTS_CUDA(cudaFuncSetAttribute(tc_pair_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  int total = 0;
  for (int i = 0; i < 3; ++i) {
    L.tile_start[i] = total;
    if (i < L.nprob) {
      const int V = R - (L.p[i].k - 1);
      L.tiles_per_row[i] = (L.T_rows + V - 1) / V;
      total += L.tiles_per_row[i] * L.B;
    } else {
      L.tiles_per_row[i] = 1;
      L.tile_start[i] = 0x7fffffff;   // never selected
    }
  }
  L.ntiles = total;
  const int grid = total < ctx->sm_count ? total : ctx->sm_count;
  tc_pair_kernel<N><<<grid, NTHREADS, Cfg::SMEM_BYTES, st>>>(L);
  ctx->launches++;
  VTTS_CUDA(cudaGetLastError());
  return VTTS_OK;
}

}  // namespace

int vtts_launch_tc_pair(vtts_ctx* ctx, TcPairLaunch& L, cudaStream_t st) {
  if (ctx->pair_ts == 1) return vtts_launch_tc_pair_ts(ctx, L, st);
  if (ctx->pair_ts == 2) return vtts_launch_tc_pair2(ctx, L, st);
  if (ctx->pair_ts == 3) return vtts_launch_tc_pair2c(ctx, L, st);
  if (L.nprob < 1 || L.nprob > 3) return ctx->fail(VTTS_ERR_BAD_ARG, "tc_pair: nprob %d", L.nprob);
  for (int i = 0; i < L.nprob; ++i) {
    const TcPairProb& p = L.p[i];
    if (p.k < 1 || (p.k & 1) == 0 || (p.k - 1) * p.dil > 50 || p.k - 1 > 15) return ctx->fail(VTTS_ERR_BAD_ARG, "tc_pair: k=%d dil=%d", p.k, p.dil);
    if (p.x == p.out) return ctx->fail(VTTS_ERR_BAD_ARG, "tc_pair: in-place not supported (tiles read halo rows of x)");
  }
  L.err = ctx->d_err;
  L.dbg = ctx->tc_dbg_on ? ctx->d_tc_dbg : nullptr;
  switch (L.N) {
    case 64: return launch_pair<64>(ctx, L, st);
    case 32: return launch_pair<32>(ctx, L, st);
    default: return ctx->fail(VTTS_ERR_BAD_ARG, "tc_pair: N %d unsupported", L.N);
  }
}
