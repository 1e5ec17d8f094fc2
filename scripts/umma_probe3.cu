// Third tcgen05 probe: the CTA-pair form (cta_group::2, M = 256 over two SMs of a cluster).
//   part A  numerics of ONE M=256 x N=128 x K=16 MMA with the production operand layout (K-major, no swizzle, rows 16 B
//           apart): CTA r of the pair holds A rows [128 r, 128 r + 128) and B rows (output columns) [64 r, 64 r + 64);
//           each CTA's TMEM receives its 128 rows x all 128 columns.
//   part B  issue cost: clk per region of MT bf16x3 triples (as umma_probe2 variant 0) for the pair form vs the
//           single-CTA form at the same N -- the pair fetches only half of the weight operand per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/umma_probe3 scripts/umma_probe3.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../viettts_b200/csrc/tc_common.cuh"
using namespace tcx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
constexpr int SMEM_BYTES = 200 * 1024;

struct Args { int N; int iters; int variant; long long* out; float* d_out; };

// variant 0: pair form, SS triples with collector hints     1: pair form, no hints
//         2: single-CTA form (cta_group::1) inside the same cluster kernel, both CTAs issue for themselves (control)
template <int MT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe3(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t rank = cluster_rank();
  const bool pair = a.variant != 2;
  for (int i = tid; i < (SMEM_BYTES - 2048) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); mbar_arrive(&bar3); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    if (pair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int N = a.N;
  const int NB = pair ? N / 2 : N;                    // weight rows held by this CTA
  const uint32_t idesc = pair ? make_idesc2(N) : make_idesc(N);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t a16 = sbase >> 4, b16 = (sbase + 96 * 1024) >> 4;
  const uint64_t a_tmpl = make_desc(0, 1024 * 16, 128);
  const uint64_t b_tmpl = make_desc(0, 2 * NB * 16, 128);       // [k-half][plane hi|lo][NB rows][16 B]
  if (warp == 0 && (rank == 0 || !pair)) {
    const long long t0 = clock64();
    for (int it = 0; it < a.iters; ++it) {
      const uint32_t aoff = a16 + (uint32_t)((it * 5) & 31);
      const uint64_t bd0 = b_tmpl | (uint64_t)(b16 + (it & 3) * 1024);
      const uint64_t bd1 = b_tmpl | (uint64_t)(b16 + (it & 3) * 1024 + NB);
      if (elect_one()) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const uint32_t d = tmem_base + (uint32_t)(mt * N) % 512u;
          const uint64_t ad_hi = a_tmpl | (uint64_t)(aoff + mt % 6 * 128);
          const uint64_t ad_lo = a_tmpl | (uint64_t)(aoff + mt % 6 * 128 + 2048);
          if (a.variant == 0 || a.variant >= 3) {
            umma2<1>(d, ad_hi, bd0, idesc, 1u); umma2<2>(d, ad_hi, bd1, idesc, 1u); umma2<0>(d, ad_lo, bd0, idesc, 1u);
          } else if (a.variant == 1) {
            umma2<0>(d, ad_hi, bd0, idesc, 1u); umma2<0>(d, ad_hi, bd1, idesc, 1u); umma2<0>(d, ad_lo, bd0, idesc, 1u);
          } else {
            umma<1>(d, ad_hi, bd0, idesc, 1u); umma<2>(d, ad_hi, bd1, idesc, 1u); umma<0>(d, ad_lo, bd0, idesc, 1u);
          }
        }
      }
      __syncwarp();
      if (a.variant == 3) { if (elect_one()) umma_commit2(&bar2); __syncwarp(); }       // a multicast commit per region (nobody waits on it)
      if (a.variant == 4) { if (elect_one()) umma_commit2(&bar2); __syncwarp(); long long dd = 0; mbar_wait_tc(&bar3, 0, nullptr, 0, dd); }
      if (a.variant == 5) { long long dd = 0; mbar_wait_tc(&bar3, 0, nullptr, 0, dd); }  // completed try_wait.acquire.cluster per region
      if (a.variant == 6) { long long dd = 0; mbar_wait_t(&bar3, 0, nullptr, 0, dd); }   // completed try_wait (acquire.cta) per region
    }
    if (elect_one()) { if (pair) umma_commit2(&bar); else umma_commit(&bar); }
    __syncwarp();
    long long acc = 0;
    mbar_wait_t(&bar, 0, nullptr, 0, acc);
    if (lane == 0) a.out[blockIdx.x] = clock64() - t0;
  } else if (warp == 0) {
    long long acc = 0;
    mbar_wait_t(&bar, 0, nullptr, 0, acc);              // the pair's commit arrives on the peer's barrier too
    if (lane == 0) a.out[blockIdx.x] = -acc;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// part A: one MMA, numerics.  out[cta][row 0..127][col 0..127]
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe3_check(float* out) {
  __shared__ __align__(1024) uint8_t a_s[128 * 32];     // [k-half][row][8 bf16]
  __shared__ __align__(1024) uint8_t b_s[64 * 32];      // [k-half][n local][8 bf16]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t rank = cluster_rank();
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(a_s);
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(b_s);
  for (int i = tid; i < 128 * 16; i += 128) {
    const int row = i / 16, k = i % 16, g = rank * 128 + row;
    A[((k / 8) * 128 + row) * 8 + k % 8] = __float2bfloat16((float)(g % 13) + 0.25f * k);
  }
  for (int i = tid; i < 64 * 16; i += 128) {
    const int nl = i / 16, k = i % 16, n = rank * 64 + nl;
    B[((k / 8) * 64 + nl) * 8 + k % 8] = __float2bfloat16(k == n % 16 ? (float)(1 + n / 16) : 0.f);
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0 && rank == 0) {
    if (elect_one()) {
      const uint64_t ad = make_desc(smem_u32(a_s), 128 * 16, 128);
      const uint64_t bd = make_desc(smem_u32(b_s), 64 * 16, 128);
      umma2<0>(tmem_base, ad, bd, make_idesc2(128), 0u);
      umma_commit2(&bar);
    }
    __syncwarp();
  }
  long long acc = 0;
  mbar_wait_t(&bar, 0, nullptr, 0, acc);
  tc_fence_after();
  for (int c0 = 0; c0 < 128; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int q = 0; q < 16; ++q) out[((size_t)blockIdx.x * 128 + warp * 32 + lane) * 128 + c0 + q] = __uint_as_float(r[q]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

template <int MT>
static double run(int N, int variant, int iters, long long* d_out) {
  static bool attr = false;
  if (!attr) { CK(cudaFuncSetAttribute(probe3<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); attr = true; }
  Args a{N, iters, variant, d_out, nullptr};
  probe3<MT><<<148, 128, SMEM_BYTES>>>(a);
  CK(cudaDeviceSynchronize());
  static long long h[148];
  CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  double s = 0;
  int n = 0;
  for (int b = 0; b < 148; ++b) if (h[b] > 0) { s += (double)h[b]; ++n; }
  return s / n / iters;
}

int main() {
  // ---- part A ----
  float* d_o;
  CK(cudaMalloc(&d_o, sizeof(float) * 2 * 128 * 128));
  CK(cudaMemset(d_o, 0xff, sizeof(float) * 2 * 128 * 128));
  probe3_check<<<2, 128>>>(d_o);
  CK(cudaDeviceSynchronize());
  static float h_o[2 * 128 * 128];
  CK(cudaMemcpy(h_o, d_o, sizeof(h_o), cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int g = 0; g < 256; ++g)
    for (int n = 0; n < 128; ++n) {
      const float want = ((float)(g % 13) + 0.25f * (n % 16)) * (float)(1 + n / 16);
      const float got = h_o[(size_t)g * 128 + n];
      if (want != got && bad++ < 8) printf("  mismatch row %d col %d: got %g want %g\n", g, n, got, want);
    }
  printf("umma_probe3 part A: one cta_group::2 MMA (M=256, N=128, K=16, A rows and B rows split over the CTA pair): %s (%d mismatches of 32768)\n",
         bad ? "FAIL" : "exact", bad);

  // ---- part B ----
  long long* d_out;
  CK(cudaMalloc(&d_out, sizeof(long long) * 148));
  const int iters = 2048;
  run<1>(128, 0, 64, d_out);
  const char* names[7] = {"pair form (cta_group::2, M=256), collector hints", "pair form, no hints", "single-CTA form (M=128), same kernel, both CTAs issue",
                          "pair form + one multicast commit per region", "pair form + commit + completed try_wait.acquire.cluster per region",
                          "pair form + completed try_wait.acquire.cluster per region", "pair form + completed try_wait (acquire.cta) per region"};
  printf("\npart B: clk per region of MT bf16x3 triples (3*MT MMAs), 74 CTA pairs\n");
  for (int v = 0; v < 7; ++v) {
    printf("[%d] %s\n", v, names[v]);
    for (int N : {64, 128, 256}) {
      const double c1 = run<1>(N, v, iters, d_out), c2 = run<2>(N, v, iters, d_out), c4 = run<4>(N, v, iters, d_out);
      printf("  N=%3d  MT=1 %7.1f  MT=2 %7.1f  MT=4 %7.1f   -> per-MMA (MT=4) %.1f clk, math floor %d clk/MMA\n", N, c1, c2, c4, c4 / 12.0, N / 2);
    }
  }
  return 0;
}
