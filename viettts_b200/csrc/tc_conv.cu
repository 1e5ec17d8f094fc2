// Tensor-core NWC conv1d for sm_100a: implicit GEMM on tcgen05.mma with TMEM accumulators.
//
// Same operator as conv1d.cu (hk.Conv1D of vietTTS/hifigan/model.py:21-41 with the leaky_relu /
// 3-way-mean / bias / residual fusions), but the contraction runs on the 5th-gen tensor cores in
// "bf16x3" arithmetic:  every fp32 operand v is split into hi = bf16(v), lo = bf16(v - hi) and the
// product is accumulated in fp32 as  a_hi*w_hi + a_hi*w_lo + a_lo*w_hi  (the dropped a_lo*w_lo
// term and the split truncation are ~2^-17 relative; measured end-to-end waveform error of the
// whole generator vs float64: L-inf 1.7e-5, RMS 3.7e-6 -- see DESIGN.md).
//
// GEMM view per tap j:  D[time, cout] += A_j[time, cin] * W_j[cin, cout]
//   M = 128 time rows per MMA (TMEM lane = row), N = Cout (<= 256 TMEM columns), K = 16 channels.
//   A operand: activations in shared memory, K-major, NO swizzle, rows 16 B apart:
//       [plane hi|lo][k-half (8 ch)][row][8 x bf16]
//     so tap j / dilation d is just a start-address offset of j*d*16 bytes in the descriptor.
//   B operand: weights pre-split and pre-packed at load time into the same canonical layout,
//     streamed with cp.async.bulk (TMA bulk copy) through a 4-stage mbarrier ring.
//
// One persistent CTA per SM, 14 warps:
//   warps 0-3  epilogue   tcgen05.ld TMEM -> regs, + bias (+ residual), fp32 NWC store
//   warp  4    MMA issue  (one lane) + TMEM alloc/dealloc
//   warp  5    weight producer (one lane, cp.async.bulk + expect_tx)
//   warps 6-13 activation converters: fp32 global -> [mean3] -> leaky_relu -> hi/lo bf16 -> smem
// A "super tile" is MT = min(4, 512/N) M-tiles (128*MT rows) that share every weight stage, so a
// weight block fetched from L2 feeds MT MMAs.
#include <cuda_bf16.h>

#include "vtts_internal.cuh"

namespace {

constexpr int NA = 3;               // activation stages
constexpr int NW = 4;               // weight stages
constexpr int NTHREADS = 448;     // 4 epilogue + MMA + weight producer + 8 converter warps
constexpr int NCONV = 256;        // converter threads
constexpr long long SPIN_TIMEOUT = 4000000000LL;  // ~2 s of SM clocks: trap instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __noinline__ void spin_fail(int* err, int code) {
  if (err) atomicExch(err, code);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > SPIN_TIMEOUT) spin_fail(err, code);
  }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N
__host__ __device__ constexpr uint32_t make_idesc(int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float lrelu(float v, float s) { return v >= 0.f ? v : s * v; }

// split 4 floats into packed bf16 hi / lo
__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z),
                      h3 = __float2bfloat16_rn(v.w);
  const __nv_bfloat16 l0 = __float2bfloat16_rn(v.x - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v.y - __bfloat162float(h1)),
                      l2 = __float2bfloat16_rn(v.z - __bfloat162float(h2)), l3 = __float2bfloat16_rn(v.w - __bfloat162float(h3));
  hi.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  hi.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
  lo.x = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  lo.y = (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16);
}

template <int N>
struct TcCfg {
  static constexpr int MT = (512 / N) < 4 ? (512 / N) : 4;
  static constexpr int R = 128 * MT;          // output rows per super tile
  static constexpr int RA = R + 64;           // allocated activation rows per stage (halo <= 50)
  static constexpr int A_STAGE = RA * 64;     // bytes: 2 planes x 2 k-halves x RA rows x 16 B
  static constexpr int W_STAGE = N * 64;      // bytes: 2 planes x 2 k-halves x N rows x 16 B
  static constexpr int TMEM_COLS = MT * N;    // 512, 512, 256, 128
  static constexpr int NBAR = 2 * NA + 2 * NW + 2;
  static constexpr int SMEM_BYTES = NA * A_STAGE + NW * W_STAGE + NBAR * 8 + 16 + 1024;
};

template <int N>
__global__ void __launch_bounds__(NTHREADS, 1) tc_conv_kernel(const __grid_constant__ TcLaunch L) {
  using Cfg = TcCfg<N>;
  constexpr int MT = Cfg::MT, R = Cfg::R, RA = Cfg::RA;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_st = smem;
  uint8_t* w_st = smem + NA * Cfg::A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_st + NW * Cfg::W_STAGE);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + NA;
  uint64_t* w_full = bars + 2 * NA;
  uint64_t* w_empty = bars + 2 * NA + NW;
  uint64_t* tmem_full = bars + 2 * NA + 2 * NW;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 5 && lane == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], NCONV); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nprob = L.nprob;
  const int Cin = L.Cin;
  const int nch = Cin / 16;
  const int tiles_per_row = L.tiles_per_row;
  const int ntiles = L.ntiles;

  // every role walks the same tile sequence
#define TILE_LOOP_BEGIN                                                        \
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {              \
    const int pi = tile % nprob;                                               \
    const int rest = tile / nprob;                                             \
    const int tt = rest % tiles_per_row;                                       \
    const int b = rest / tiles_per_row;                                        \
    const int tau0 = tt * R;                                                   \
    int valid = L.T_rows;                                                      \
    if (L.len) {                                                               \
      const int v = L.len[b] * L.len_mul;                                      \
      valid = v < valid ? v : valid;                                           \
    }                                                                          \
    if (tau0 >= valid) continue;                                               \
    const TcProb& P = L.p[pi];
#define TILE_LOOP_END }

  if (warp == 4) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(N);
      uint32_t sa = 0, pa = 0, sw = 0, pw = 0, tph = 0;
      TILE_LOOP_BEGIN
        (void)b; (void)tau0;
        const int k = P.k, dil = P.dil;
        mbar_wait(tmem_empty, tph ^ 1, L.err, 1);
        tc_fence_after();
        for (int c = 0; c < nch; ++c) {
          mbar_wait(&a_full[sa], pa, L.err, 2);
          tc_fence_after();
          const uint32_t a_base = smem_u32(a_st + sa * Cfg::A_STAGE);
          for (int j = 0; j < k; ++j) {
            mbar_wait(&w_full[sw], pw, L.err, 3);
            tc_fence_after();
            const uint32_t w_base = smem_u32(w_st + sw * Cfg::W_STAGE);
            const uint64_t b_hi = make_desc(w_base, N * 16, 128);
            const uint64_t b_lo = make_desc(w_base + 2 * N * 16, N * 16, 128);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint32_t row = mt * 128 + j * dil;
              const uint64_t a_hi = make_desc(a_base + row * 16, RA * 16, 128);
              const uint64_t a_lo = make_desc(a_base + 2 * RA * 16 + row * 16, RA * 16, 128);
              const uint32_t d = tmem_base + mt * N;
              umma(d, a_hi, b_hi, idesc, (c | j) != 0 ? 1u : 0u);
              umma(d, a_hi, b_lo, idesc, 1u);
              umma(d, a_lo, b_hi, idesc, 1u);
            }
            umma_commit(&w_empty[sw]);
            if (++sw == NW) { sw = 0; pw ^= 1; }
          }
          umma_commit(&a_empty[sa]);
          if (++sa == NA) { sa = 0; pa ^= 1; }
        }
        umma_commit(tmem_full);
        tph ^= 1;
      TILE_LOOP_END
    }
    __syncwarp();
  } else if (warp == 5) {
    // ============================ weight producer ============================
    if (lane == 0) {
      uint32_t sw = 0, pw = 0;
      TILE_LOOP_BEGIN
        (void)b; (void)tau0;
        const int k = P.k;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(P.wpk);
        for (int s = 0; s < nch * k; ++s) {
          mbar_wait(&w_empty[sw], pw ^ 1, L.err, 4);
          mbar_expect_tx(&w_full[sw], Cfg::W_STAGE);
          bulk_g2s(w_st + sw * Cfg::W_STAGE, wsrc + (size_t)s * Cfg::W_STAGE, Cfg::W_STAGE, &w_full[sw]);
          if (++sw == NW) { sw = 0; pw ^= 1; }
        }
      TILE_LOOP_END
    }
    __syncwarp();
  } else if (warp >= 6) {
    // ============================ activation converters ============================
    const int ct = tid - 192;          // 0..255
    const int q = ct & 3;              // 4-channel group inside the 16-channel chunk
    const int r0 = ct >> 2;            // 0..63
    const int pre_mode = L.pre_mode;
    const float slope = L.pre_slope;
    const int ld = L.in_ld;
    uint32_t sa = 0, pa = 0;
    TILE_LOOP_BEGIN
      const int k = P.k, dil = P.dil;
      const int rows = R + (k - 1) * dil;
      const size_t in_base = (size_t)b * L.T_rows * ld;
      const float* x0 = P.x0 + in_base;
      const float* x1 = pre_mode == 2 ? P.x1 + in_base : nullptr;
      const float* x2 = pre_mode == 2 ? P.x2 + in_base : nullptr;
      const int row_base = tau0 + P.in_off;
      for (int c = 0; c < nch; ++c) {
        mbar_wait(&a_empty[sa], pa ^ 1, L.err, 5);
        uint8_t* st = a_st + sa * Cfg::A_STAGE + ((q >> 1) * RA) * 16 + (q & 1) * 8;
        const int coff = c * 16 + q * 4;
        constexpr int U = 8;           // loads in flight per thread (memory-level parallelism)
        for (int rr0 = r0; rr0 < rows; rr0 += 64 * U) {
          float4 v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int rr = rr0 + u * 64;
            const int t = row_base + rr;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rr < rows && t >= 0 && t < valid) {
              const size_t off = (size_t)t * ld + coff;
              v[u] = __ldg(reinterpret_cast<const float4*>(x0 + off));
              if (pre_mode == 2) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(x1 + off));
                const float4 bb = __ldg(reinterpret_cast<const float4*>(x2 + off));
                v[u].x = ((v[u].x + a.x) + bb.x) / 3.0f;
                v[u].y = ((v[u].y + a.y) + bb.y) / 3.0f;
                v[u].z = ((v[u].z + a.z) + bb.z) / 3.0f;
                v[u].w = ((v[u].w + a.w) + bb.w) / 3.0f;
              }
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int rr = rr0 + u * 64;
            if (rr < rows) {
              float4 x = v[u];
              if (pre_mode >= 1) {
                x.x = lrelu(x.x, slope); x.y = lrelu(x.y, slope); x.z = lrelu(x.z, slope); x.w = lrelu(x.w, slope);
              }
              uint2 hi, lo;
              split4(x, hi, lo);
              *reinterpret_cast<uint2*>(st + (size_t)rr * 16) = hi;
              *reinterpret_cast<uint2*>(st + (size_t)(2 * RA + rr) * 16) = lo;
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(&a_full[sa]);
        if (++sa == NA) { sa = 0; pa ^= 1; }
      }
    TILE_LOOP_END
  } else {
    // ============================ epilogue (warps 0-3) ============================
    uint32_t tph = 0;
    const int out_ld = L.out_ld;
    TILE_LOOP_BEGIN
      mbar_wait(tmem_full, tph, L.err, 6);
      tc_fence_after();
      const size_t out_base = (size_t)b * L.rows_out * out_ld;
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int tau = tau0 + mt * 128 + warp * 32 + lane;
        const bool ok = tau < valid;
        const size_t orow = out_base + (size_t)(tau * P.out_stride + P.out_off) * out_ld;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + mt * N;
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t r[32];
          tmem_ld16(taddr + c0, r);
          tmem_ld16(taddr + c0 + 16, r + 16);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const int n = c0 + g * 4;
              const float4 bi = __ldg(reinterpret_cast<const float4*>(P.bias + n));
              float4 o;
              o.x = __uint_as_float(r[g * 4 + 0]) + bi.x;
              o.y = __uint_as_float(r[g * 4 + 1]) + bi.y;
              o.z = __uint_as_float(r[g * 4 + 2]) + bi.z;
              o.w = __uint_as_float(r[g * 4 + 3]) + bi.w;
              if (P.resid) {
                const float4 rs = __ldg(reinterpret_cast<const float4*>(P.resid + orow + n));
                o.x += rs.x; o.y += rs.y; o.z += rs.z; o.w += rs.w;
              }
              *reinterpret_cast<float4*>(P.out + orow + n) = o;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty);
      tph ^= 1;
    TILE_LOOP_END
  }
#undef TILE_LOOP_BEGIN
#undef TILE_LOOP_END

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

// fp32 Haiku conv weight w[k][Cin][Cout_total] -> packed bf16 blocks for output columns [n0, n0+N):
//   [chunk c = Cin/16][tap j][plane hi|lo][k-half][n][8]
__global__ void pack_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int k, int Cin, int Cout_total, int n0, int N) {
  const size_t total = (size_t)k * Cin * N;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int n = idx % N;
    const int i = (idx / N) % Cin;
    const int j = idx / ((size_t)N * Cin);
    const float v = w[((size_t)j * Cin + i) * Cout_total + n0 + n];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const int c = i / 16, kh = (i % 16) / 8, e = i % 8;
    const size_t blk = ((size_t)c * k + j) * (size_t)(4 * N * 8);
    dst[blk + ((size_t)(0 * 2 + kh) * N + n) * 8 + e] = hi;
    dst[blk + ((size_t)(1 * 2 + kh) * N + n) * 8 + e] = lo;
  }
}

template <int N>
int launch_n(vtts_ctx* ctx, TcLaunch& L, cudaStream_t st) {
  using Cfg = TcCfg<N>;
  static bool attr_done = false;
  if (!attr_done) {
    VTTS_CUDA(cudaFuncSetAttribute(tc_conv_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  L.tiles_per_row = (L.T_rows + Cfg::R - 1) / Cfg::R;
  L.ntiles = L.nprob * L.tiles_per_row * L.B;
  const int grid = L.ntiles < ctx->sm_count ? L.ntiles : ctx->sm_count;
  tc_conv_kernel<N><<<grid, NTHREADS, Cfg::SMEM_BYTES, st>>>(L);
  ctx->launches++;
  VTTS_CUDA(cudaGetLastError());
  return VTTS_OK;
}

}  // namespace

size_t vtts_tc_packed_elems(int k, int Cin, int N) { return (size_t)k * Cin * N * 2; }

int vtts_tc_pack_weights(vtts_ctx* ctx, const float* w, void* dst, int k, int Cin, int Cout_total, int n0, int N) {
  pack_w_kernel<<<256, 256>>>(w, reinterpret_cast<__nv_bfloat16*>(dst), k, Cin, Cout_total, n0, N);
  VTTS_CUDA(cudaGetLastError());
  return VTTS_OK;
}

int vtts_launch_tc_conv(vtts_ctx* ctx, TcLaunch& L, cudaStream_t st) {
  if (L.nprob < 1 || L.nprob > 8) return ctx->fail(VTTS_ERR_BAD_ARG, "tc_conv: nprob %d", L.nprob);
  if (L.Cin % 16 != 0) return ctx->fail(VTTS_ERR_BAD_ARG, "tc_conv: Cin %d", L.Cin);
  for (int i = 0; i < L.nprob; ++i)
    if ((L.p[i].k - 1) * L.p[i].dil > 50 || L.p[i].k < 1) return ctx->fail(VTTS_ERR_BAD_ARG, "tc_conv: halo too large");
  L.err = ctx->d_err;
  switch (L.N) {
    case 256: return launch_n<256>(ctx, L, st);
    case 128: return launch_n<128>(ctx, L, st);
    case 64: return launch_n<64>(ctx, L, st);
    case 32: return launch_n<32>(ctx, L, st);
    default: return ctx->fail(VTTS_ERR_BAD_ARG, "tc_conv: N %d unsupported", L.N);
  }
}
