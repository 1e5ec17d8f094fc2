// Second tcgen05 probe: cost model of the MMA issue loop.  One `if (elect_one()) { ... }` region per iteration holding
// MT "triples" (a_hi.w_hi fill | a_hi.w_lo lastuse | a_lo.w_hi) on MT different accumulators, exactly the shape of the
// production loop in tc_conv.cu.  Fitting clk/iteration = a + b * (3 MT) separates the per-region cost (a) from the
// per-MMA cost (b); SS (A in smem) vs TS (A in TMEM) separates operand fetch from issue cost.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/umma_probe2 scripts/umma_probe2.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../viettts_b200/csrc/tc_common.cuh"
using namespace tcx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
constexpr int SMEM_BYTES = 200 * 1024;

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct Args { int N; int iters; int variant; long long* out; int nwarps; };

// variant: 0 SS triples, distinct A windows per M tile (production)      1 TS triples
//          2 SS, no collector hints                                      3 SS, every MMA the SAME A and B descriptor
//          4 SS triples + a completed mbarrier try_wait per region (production has one per group)
template <int MT>
__global__ void __launch_bounds__(128, 1) probe2(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  for (int i = tid; i < (SMEM_BYTES - 2048) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, a.nwarps); mbar_init(&bar2, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int N = a.N;
  const uint32_t idesc = make_idesc(N);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t a16 = sbase >> 4, b16 = (sbase + 96 * 1024) >> 4;
  const uint64_t a_tmpl = make_desc(0, 1024 * 16, 128);
  const uint64_t b_tmpl = make_desc(0, 256 * 16, 128);
  long long t_role = 0;
  if (warp < a.nwarps) {
    if (a.variant == 4 && lane == 0 && warp == 0) mbar_arrive(&bar2);    // phase 0 complete: try_wait(parity 0) returns true immediately
    __syncwarp();
    const long long t0 = clock64();
    long long dummy = 0;
    for (int it = 0; it < a.iters; ++it) {
      if (a.variant == 4) { mbar_wait_t(&bar2, 0, nullptr, 0, dummy); tc_fence_after(); }
      const uint32_t aoff = a16 + (uint32_t)((it * 5) & 31);
      const uint64_t bd0 = b_tmpl | (uint64_t)(b16 + (it & 3) * 1024);
      const uint64_t bd1 = b_tmpl | (uint64_t)(b16 + (it & 3) * 1024 + 512);
      if (elect_one()) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const uint32_t d = tmem_base + (uint32_t)((warp * MT + mt) * (N < 32 ? 32 : N)) % 256u;
          if (a.variant == 1) {
            const uint32_t slot = tmem_base + 256 + (uint32_t)((it + mt + warp * 2) & 7) * 16;
            umma_ts(d, slot, bd0, idesc, 1u);
            umma_ts(d, slot, bd1, idesc, 1u);
            umma_ts(d, slot + 8, bd0, idesc, 1u);
          } else if (a.variant == 3) {
            const uint64_t ad = a_tmpl | (uint64_t)a16;
            const uint64_t bd = b_tmpl | (uint64_t)b16;
            umma<0>(d, ad, bd, idesc, 1u); umma<0>(d, ad, bd, idesc, 1u); umma<0>(d, ad, bd, idesc, 1u);
          } else {
            const uint64_t ad_hi = a_tmpl | (uint64_t)(aoff + (mt + warp * MT) % 6 * 128);
            const uint64_t ad_lo = a_tmpl | (uint64_t)(aoff + (mt + warp * MT) % 6 * 128 + 2048);
            if (a.variant == 2) {
              umma<0>(d, ad_hi, bd0, idesc, 1u); umma<0>(d, ad_hi, bd1, idesc, 1u); umma<0>(d, ad_lo, bd0, idesc, 1u);
            } else {
              umma<1>(d, ad_hi, bd0, idesc, 1u); umma<2>(d, ad_hi, bd1, idesc, 1u); umma<0>(d, ad_lo, bd0, idesc, 1u);
            }
          }
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    long long acc = 0;
    mbar_wait_t(&bar, 0, nullptr, 0, acc);
    t_role = clock64() - t0;
    if (lane == 0 && warp == 0) a.out[blockIdx.x] = t_role + (dummy & 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int MT>
static double run(int N, int variant, int iters, long long* d_out, int nwarps = 1) {
  static bool attr = false;
  if (!attr) { CK(cudaFuncSetAttribute(probe2<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); attr = true; }
  Args a{N, iters, variant, d_out, nwarps};
  probe2<MT><<<148, 128, SMEM_BYTES>>>(a);
  CK(cudaDeviceSynchronize());
  static long long h[148];
  CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  double s = 0;
  for (int b = 0; b < 148; ++b) s += (double)h[b];
  return s / 148 / iters;
}

int main() {
  long long* d_out;
  CK(cudaMalloc(&d_out, sizeof(long long) * 148));
  const int iters = 2048;
  run<1>(64, 0, 64, d_out);
  const char* names[5] = {"SS triples (production shape)", "TS triples (A in TMEM)", "SS triples, no collector hints", "SS, same A and B every MMA",
                          "SS triples + one completed mbarrier try_wait per region"};
  printf("umma_probe2: clk per region (= one `if (elect_one())` block holding MT triples = 3*MT MMAs), 148 CTAs, %d regions each\n", iters);
  for (int v = 0; v < 5; ++v) {
    printf("\n[%d] %s\n", v, names[v]);
    for (int N : {32, 64, 128}) {
      const double c1 = run<1>(N, v, iters, d_out), c2 = run<2>(N, v, iters, d_out), c4 = run<4>(N, v, iters, d_out), c8 = run<8>(N, v, iters, d_out);
      printf("  N=%3d  MT=1 %7.1f  MT=2 %7.1f  MT=4 %7.1f  MT=8 %7.1f   -> per-MMA slope (MT 4->8) %.1f clk, math floor %d clk/MMA\n", N, c1, c2, c4,
             c8, (c8 - c4) / 12.0, N / 2);
    }
  }
  printf("\n[6] several issuing warps at once (each its own accumulators), clk per region-set and aggregate clk per MMA\n");
  for (int v : {0, 1}) {
    printf("  %s\n", names[v]);
    for (int N : {32, 64}) {
      for (int W : {1, 2, 4}) {
        const double c1 = run<1>(N, v, iters, d_out, W), c2 = run<2>(N, v, iters, d_out, W);
        printf("    N=%3d warps=%d  MT=1: %7.1f (%5.1f /MMA)  MT=2: %7.1f (%5.1f /MMA)\n", N, W, c1, c1 / (3.0 * W), c2, c2 / (6.0 * W));
      }
    }
  }
  return 0;
}
