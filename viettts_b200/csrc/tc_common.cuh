// Device helpers shared by the tcgen05 kernels (tc_conv.cu, tc_pair.cu): mbarrier / bulk-copy / UMMA wrappers.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace tcx {

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
static __device__ __noinline__ void spin_fail(int* err, int code) {
  if (err) atomicExch(err, code);
  __trap();
}
// wait and add the stalled cycles to a per-role counter (profiling aid, see vtts_debug_tc_stats)
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, int* err, int code, long long& acc) {
  // mbarrier.try_wait suspends the thread in hardware for a while before it returns false, so the first probe is
  // part of the wait: time the whole thing (two clock reads, small next to the ~90 clk of a completed try_wait)
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > SPIN_TIMEOUT) spin_fail(err, code);
  }
  acc += clock64() - t0;
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
// COL selects the A-operand collector hint: 0 none, 1 fill (keep A for the next MMA), 2 lastuse (reuse the kept A).
// Consecutive MMAs that share the SAME A descriptor (a_hi x W_hi then a_hi x W_lo) read the 4 KB A tile from shared
// memory once instead of twice -- shared-memory read bandwidth, not the tensor pipe, bounds the N <= 64 layers.
template <int COL = 0>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (COL == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else if constexpr (COL == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- CTA-pair form (cta_group::2): two CTAs of a cluster issue ONE M=256 MMA; CTA r holds A rows [128r, 128r+128) and
// B rows (output columns) [N/2 r, N/2 r + N/2), each CTA's TMEM receives its 128 rows x N columns.  Only rank 0 issues;
// the commit is multicast to the barrier at the same shared-memory offset in both CTAs.  (scripts/umma_probe3.cu)
__host__ __device__ constexpr uint32_t make_idesc2(int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
template <int COL = 0>
__device__ __forceinline__ void umma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (COL == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else if constexpr (COL == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster (release at cluster scope)
// CLUSTER_REL = 0: default semantics (release at CTA scope) -- for signals that do not publish this thread's own writes
template <int CLUSTER_REL = 1>
__device__ __forceinline__ void mbar_arrive_rank(uint64_t* bar, uint32_t rank) {
  uint32_t addr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(smem_u32(bar)), "r"(rank));
  if constexpr (CLUSTER_REL) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
  else asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
// wait with acquire at cluster scope (the arrivals may come from the peer CTA).  POLL: test_wait (never suspends).
template <int POLL = 0>
__device__ __forceinline__ void mbar_wait_tc(uint64_t* bar, uint32_t parity, int* err, int code, long long& acc) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    if constexpr (POLL)
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(bar)), "r"(parity)
          : "memory");
    else
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(bar)), "r"(parity)
          : "memory");
    if (ok) break;
    if (clock64() - t0 > SPIN_TIMEOUT) spin_fail(err, code);
  }
  acc += clock64() - t0;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// one elected lane of a fully converged warp (elect.sync): lets ptxas issue uniform-datapath
// instructions (UTCHMMA, UTCBAR) without a per-thread waterfall loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// read-only 16 B load that asks L2 to fetch the whole 256 B block around it: the converters read 64 B of every 512 B (or
// 256 B, 128 B) activation row per 16-channel chunk, so the next three chunks of the same rows then hit L2 instead of DRAM
__device__ __forceinline__ float4 ldg_pf256(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// leaky_relu for 0 < s < 1: max(v, s*v) (2 instructions; identical to v >= 0 ? v : s*v for every finite v and +-0)
__device__ __forceinline__ float lrelu(float v, float s) { return fmaxf(v, s * v); }

// split 4 floats into packed bf16 hi / lo planes: hi = bf16_rn(v), lo = bf16_rn(v - hi).
// Packed conversions (cvt.rn.bf16x2.f32) do two values per instruction and deliver the pairs already packed.
__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
  const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y);
  const __nv_bfloat162 h23 = __floats2bfloat162_rn(v.z, v.w);
  const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01);
  const uint32_t u23 = *reinterpret_cast<const uint32_t*>(&h23);
  // bf16 -> f32 is a 16-bit shift
  const float r0 = v.x - __uint_as_float(u01 << 16);
  const float r1 = v.y - __uint_as_float(u01 & 0xffff0000u);
  const float r2 = v.z - __uint_as_float(u23 << 16);
  const float r3 = v.w - __uint_as_float(u23 & 0xffff0000u);
  const __nv_bfloat162 l01 = __floats2bfloat162_rn(r0, r1);
  const __nv_bfloat162 l23 = __floats2bfloat162_rn(r2, r3);
  hi.x = u01;
  hi.y = u23;
  lo.x = *reinterpret_cast<const uint32_t*>(&l01);
  lo.y = *reinterpret_cast<const uint32_t*>(&l23);
}

}  // namespace tcx
