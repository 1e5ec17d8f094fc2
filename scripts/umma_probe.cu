// Micro-benchmark of the tcgen05 operand paths on B200 (sm_100a).  Questions it answers (DESIGN.md "small-N MMA"):
//   1. what does one M=128, K=16 kind::f16 MMA cost as a function of N when A comes from shared memory (SS)?
//   2. ... when A comes from tensor memory (TS)?   (is the ~65 clk floor of the N <= 64 layers the smem A fetch?)
//   3. how is a K=16 A tile laid out in TMEM for the TS form (checked numerically, not assumed)
//   4. throughput of tcgen05.st, and of an LDS.128 -> tcgen05.st "row-shift replication" loop running beside TS MMAs
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/umma_probe scripts/umma_probe.cu
// Run on the GPU box: ./scripts/umma_probe > gpurun_out/umma_probe.txt
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../viettts_b200/csrc/tc_common.cuh"

using namespace tcx;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int SMEM_BYTES = 200 * 1024;

__host__ __device__ constexpr uint32_t idesc_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct Args {
  int mode;       // see main()
  int N;          // MMA N
  int M;          // MMA M (64 or 128)
  int iters;
  int nrep;       // replicator warps (mode 5/6)
  long long* out; // [grid][8]
  float* dump;    // mode 3: [128][32] accumulator dump
};

// modes: 0 SS one MMA per distinct A tile; 1 SS production triple (hi fill, hi lastuse, lo); 2 TS triple; 3 TS layout check;
//        4 tcgen05.st throughput; 5 TS triples + replicator warps; 6 replicator warps only; 7 SS triple + idle; 8 TS single
__global__ void __launch_bounds__(416, 1) probe_kernel(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  // fill smem with small bf16 values (1.0 / 0.5 pattern) so that accumulators stay finite
  for (int i = tid; i < (SMEM_BYTES - 2048) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // bf16 2^-7 pairs
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int N = a.N, M = a.M;
  const uint32_t idesc = idesc_mn(M, N);
  const uint32_t sbase = smem_u32(smem);
  // A region: 64 KB at offset 0 (K-major no swizzle: [k-half][row][16 B], 1024 rows); B region at 96 KB: [k-half][n][16 B]
  const uint32_t a16 = sbase >> 4, b16 = (sbase + 96 * 1024) >> 4;
  const uint64_t a_tmpl = make_desc(0, 1024 * 16, 128);
  const uint64_t b_tmpl = make_desc(0, 256 * 16, 128);
  long long t_role = 0;

  if (a.mode == 3) {
    // ---- TS layout check: A[r][k] written with tcgen05.st 32x32b.x8 (lane = row, reg c = packed (k=2c, 2c+1)), B[n][k] = (n==k)
    if (warp < 4) {
      uint32_t r[8];
      const int row = warp * 32 + lane;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const __nv_bfloat162 v = __floats2bfloat162_rn((float)(row * 16 + 2 * c), (float)(row * 16 + 2 * c + 1) );
        // values up to 2047: exact in bf16 only below 256 -> use small codes instead: row%8*16+k
        const __nv_bfloat162 v2 = __floats2bfloat162_rn((float)((row % 8) * 16 + 2 * c), (float)((row % 8) * 16 + 2 * c + 1));
        (void)v;
        r[c] = *reinterpret_cast<const uint32_t*>(&v2);
      }
      tmem_st8(tmem_base + ((uint32_t)(warp * 32) << 16) + 256, r);
      tmem_st_wait();
    }
    // B: identity, N=32 rows x 16 k: element (n,k) at k-half (k/8) block: [(k/8)][n][k%8]
    __nv_bfloat16* bsm = reinterpret_cast<__nv_bfloat16*>(smem + 96 * 1024);
    for (int i = tid; i < 2 * 256 * 8; i += blockDim.x) {
      const int kh = i / (256 * 8), n = (i / 8) % 256, e = i % 8;
      bsm[i] = __float2bfloat16_rn((n == kh * 8 + e) ? 1.0f : 0.0f);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
      if (elect_one()) {
        umma_ts(tmem_base, tmem_base + 256, b_tmpl | (uint64_t)b16, idesc_mn(128, 32), 0u);
        umma_commit(&bar);
      }
      __syncwarp();
    }
    long long acc = 0;
    mbar_wait_t(&bar, 0, nullptr, 0, acc);
    tc_fence_after();
    if (warp < 4) {
      uint32_t r[32];
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16), r);
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + 16, r + 16);
      tmem_ld_wait();
      if (blockIdx.x == 0)
        for (int c = 0; c < 32; ++c) a.dump[(warp * 32 + lane) * 32 + c] = __uint_as_float(r[c]);
    }
  } else if (a.mode == 4) {
    // ---- tcgen05.st throughput: warps 0..3, x16
    if (warp < 4) {
      uint32_t r[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) r[c] = 0x3c003c00u + lane;
      const long long t0 = clock64();
      for (int it = 0; it < a.iters; ++it) {
        tmem_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + 256 + (it & 7) * 16, r);
      }
      tmem_st_wait();
      t_role = clock64() - t0;
    }
  } else {
    const bool with_mma = a.mode != 6;
    const bool with_rep = a.mode == 5 || a.mode == 6;
    if (warp == 0 && with_mma) {
      const long long t0 = clock64();
      for (int it = 0; it < a.iters; ++it) {
        const uint32_t arow = (uint32_t)((it * 7) & 63) * 8;           // distinct A windows (row shift), 8-row aligned or not
        const uint32_t aoff = a16 + ((it * 5) & 7) + arow;              // arbitrary row shift like a conv tap
        const uint64_t ad_hi = a_tmpl | (uint64_t)aoff;
        const uint64_t ad_lo = a_tmpl | (uint64_t)(aoff + 2048);        // second plane 32 KB further
        const uint64_t bd0 = b_tmpl | (uint64_t)(b16 + (it & 3) * 1024);
        const uint64_t bd1 = b_tmpl | (uint64_t)(b16 + (it & 3) * 1024 + 512);
        const uint32_t slot = tmem_base + 256 + (uint32_t)(it & 7) * 16;
        if (elect_one()) {
          if (a.mode == 0) {
            umma<0>(tmem_base, ad_hi, bd0, idesc, 1u);
          } else if (a.mode == 1 || a.mode == 7) {
            umma<1>(tmem_base, ad_hi, bd0, idesc, 1u);
            umma<2>(tmem_base, ad_hi, bd1, idesc, 1u);
            umma<0>(tmem_base, ad_lo, bd0, idesc, 1u);
          } else if (a.mode == 8) {
            umma_ts(tmem_base, slot, bd0, idesc, 1u);
          } else {   // 2, 5: TS triple
            umma_ts(tmem_base, slot, bd0, idesc, 1u);
            umma_ts(tmem_base, slot, bd1, idesc, 1u);
            umma_ts(tmem_base, slot + 8, bd0, idesc, 1u);
          }
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      long long acc = 0;
      mbar_wait_t(&bar, 0, nullptr, 0, acc);
      t_role = clock64() - t0;
    } else if (warp >= 1 && warp <= a.nrep && with_rep) {
      // replicator: 4 x LDS.128 (hi kh0, hi kh1, lo kh0, lo kh1 of row lane+32q+shift) -> one tcgen05.st x16 (one A slot)
      const int rw = warp - 1, q = rw & 3, sub = rw >> 2, nsub = a.nrep / 4;
      const long long t0 = clock64();
      for (int it = sub; it < a.iters; it += nsub) {
        const int row = q * 32 + lane + ((it * 5) & 63);
        const uint8_t* p = smem + (size_t)row * 16;
        const uint4 v0 = *reinterpret_cast<const uint4*>(p);
        const uint4 v1 = *reinterpret_cast<const uint4*>(p + 1024 * 16);
        const uint4 v2 = *reinterpret_cast<const uint4*>(p + 2048 * 16);
        const uint4 v3 = *reinterpret_cast<const uint4*>(p + 3072 * 16);
        uint32_t r[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
        tmem_st16(tmem_base + ((uint32_t)(q * 32) << 16) + 256 + (uint32_t)(it & 7) * 16, r);
        if ((it & 7) == 7) tmem_st_wait();
      }
      tmem_st_wait();
      t_role = clock64() - t0;
    }
  }
  if (lane == 0 && warp < 8) a.out[(size_t)blockIdx.x * 8 + warp] = t_role;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static double run(int mode, int N, int M, int iters, int nrep, long long* d_out, float* d_dump, int grid, int role_warp) {
  Args a{mode, N, M, iters, nrep, d_out, d_dump};
  CK(cudaMemset(d_out, 0, sizeof(long long) * 148 * 8));
  probe_kernel<<<grid, 416, SMEM_BYTES>>>(a);
  CK(cudaDeviceSynchronize());
  static long long h[148 * 8];
  CK(cudaMemcpy(h, d_out, sizeof(long long) * 148 * 8, cudaMemcpyDeviceToHost));
  double s = 0;
  for (int b = 0; b < grid; ++b) s += (double)h[b * 8 + role_warp];
  return s / grid / iters;
}

int main() {
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  long long* d_out; float* d_dump;
  CK(cudaMalloc(&d_out, sizeof(long long) * 148 * 8));
  CK(cudaMalloc(&d_dump, sizeof(float) * 128 * 32));
  const int grid = 148, iters = 4096;
  printf("umma_probe: grid=%d CTAs (one per SM), %d iterations per CTA; clk = SM clocks per iteration (mean over CTAs)\n", grid, iters);
  run(0, 64, 128, 256, 0, d_out, d_dump, grid, 0);   // warm-up
  printf("\n[1] SS, one MMA per distinct A window (no collector), M=128 K=16:\n");
  for (int N : {16, 32, 64, 128, 256}) printf("  N=%3d  %.1f clk/MMA   (math floor 128*N/256 = %d)\n", N, run(0, N, 128, iters, 0, d_out, d_dump, grid, 0), N / 2);
  printf("[1b] SS, M=64 (A tile half the bytes):\n");
  for (int N : {32, 64, 256}) printf("  N=%3d  %.1f clk/MMA\n", N, run(0, N, 64, iters, 0, d_out, d_dump, grid, 0));
  printf("[2] SS production triple (a_hi.w_hi fill | a_hi.w_lo lastuse | a_lo.w_hi):\n");
  for (int N : {32, 64, 128, 256}) printf("  N=%3d  %.1f clk/triple  (math floor %d)\n", N, run(1, N, 128, iters, 0, d_out, d_dump, grid, 0), 3 * N / 2);
  printf("[3] TS (A from TMEM), one MMA:\n");
  for (int N : {16, 32, 64, 128, 256}) printf("  N=%3d  %.1f clk/MMA\n", N, run(8, N, 128, iters, 0, d_out, d_dump, grid, 0));
  printf("[4] TS triple (hi.w_hi | hi.w_lo | lo.w_hi):\n");
  for (int N : {32, 64, 128, 256}) printf("  N=%3d  %.1f clk/triple  (math floor %d)\n", N, run(2, N, 128, iters, 0, d_out, d_dump, grid, 0), 3 * N / 2);
  printf("[5] tcgen05.st 32x32b.x16, 4 warps (one 128-lane x 16-column tile = 8 KB per 4 instructions):\n");
  printf("  %.1f clk per warp-instruction\n", run(4, 32, 128, iters, 0, d_out, d_dump, grid, 0));
  printf("[6] replicator warps only (4 x LDS.128 + tcgen05.st.x16 per warp per A slot):\n");
  for (int nrep : {4, 8}) printf("  %d warps: %.1f clk per A slot (hi+lo tile, 8 KB of smem reads)\n", nrep, run(6, 32, 128, iters, nrep, d_out, d_dump, grid, 1));
  printf("[7] TS triples with the replicators running beside them (clk per triple seen by the MMA warp | per slot seen by replicator warp 1):\n");
  for (int N : {32, 64})
    for (int nrep : {4, 8}) {
      const double m = run(5, N, 128, iters, nrep, d_out, d_dump, grid, 0);
      const double r = run(5, N, 128, iters, nrep, d_out, d_dump, grid, 1);
      printf("  N=%3d %d warps: MMA %.1f | replicator %.1f\n", N, nrep, m, r);
    }
  // layout check
  run(3, 32, 128, 1, 0, d_out, d_dump, 1, 0);
  static float hd[128 * 32];
  CK(cudaMemcpy(hd, d_dump, sizeof(hd), cudaMemcpyDeviceToHost));
  printf("[8] TS operand layout check: A[r][k] = (r%%8)*16+k stored by tcgen05.st.32x32b.x8 as packed (k=2c | k=2c+1 <<16), B = identity: expect D[r][n] = (r%%8)*16+n for n<16\n");
  int bad = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < 32; ++n) {
      const float e = n < 16 ? (float)((r % 8) * 16 + n) : 0.f;
      if (hd[r * 32 + n] != e) ++bad;
    }
  printf("  mismatches: %d of 4096\n", bad);
  for (int r : {0, 1, 9, 127}) {
    printf("  D[%3d][0..17] =", r);
    for (int n = 0; n < 18; ++n) printf(" %g", hd[r * 32 + n]);
    printf("\n");
  }
  return 0;
}
