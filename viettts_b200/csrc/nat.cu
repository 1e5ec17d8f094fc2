// NAT acoustic model forward -- restates AcousticModel.inference
// (vietTTS/nat/model.py:123-144) with batch semantics "row b == reference run on row b alone".
//
//   TokenEncoder            model.py:26-47   embed_kernel, generic conv (+BN+relu), lstm_scan_kernel
//   upsample                model.py:102-111 upsample_kernel
//   loop_fn (AR decoder)    model.py:129-142 decoder_scan_kernel  (persistent, cooperative)
//   postnet + residual      model.py:113-121,143-144 generic conv (+BN+tanh)
//
// Recurrent kernels: one cooperative grid of 128 CTAs; CTA c owns 4 hidden units of every LSTM
// layer (16 gate columns i,g,f,o) and keeps that slice of the RECURRENT weights resident in
// shared memory for the whole scan; the input projections that do not depend on the recurrence
// (token embedding path for the encoder, cond_t for the decoder) are hoisted into one GEMM
// before the scan (generic conv kernel with k=1).  State vectors live in global memory (L2) and
// are exchanged with one grid barrier per dependent phase.
#include <cooperative_groups.h>

#include "vtts_internal.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int SCAN_CTAS = 128;
constexpr int SCAN_THREADS = 256;
constexpr int UPC = 4;        // hidden units per CTA per layer
constexpr int NCOL = 16;      // 4 gates x UPC
constexpr int RG = 8;         // batch rows per register tile
constexpr int NSLICE = 64;    // K slices (SCAN_THREADS / 4 column groups)
constexpr int MAX_ROWS = 128; // batch rows per scan launch

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---- threefry2x32 (20 rounds), the counter-based generator used for VTTS_DROPOUT_SEED ----------
__host__ __device__ inline void threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t& o0, uint32_t& o1) {
  const uint32_t ks2 = 0x1BD11BDAu ^ k0 ^ k1;
  uint32_t x0 = c0 + k0, x1 = c1 + k1;
  const int R0[4] = {13, 15, 26, 6}, R1[4] = {17, 29, 16, 24};
  const uint32_t ks[3] = {k0, k1, ks2};
#pragma unroll
  for (int blk = 0; blk < 5; ++blk) {
    const int* R = (blk & 1) ? R1 : R0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      x0 += x1;
      x1 = (x1 << R[r]) | (x1 >> (32 - R[r]));
      x1 ^= x0;
    }
    x0 += ks[(blk + 1) % 3];
    x1 += ks[(blk + 2) % 3] + (uint32_t)(blk + 1);
  }
  o0 = x0;
  o1 = x1;
}

__device__ __forceinline__ float keep_scale(int mode, const uint8_t* keep, uint64_t seed, int b, int t, int N, int layer, int unit) {
  if (mode == VTTS_DROPOUT_OFF) return 1.f;
  if (mode == VTTS_DROPOUT_MASK) return keep[(((size_t)b * N + t) * 2 + layer) * vc::PRENET + unit] ? 2.f : 0.f;
  uint32_t o0, o1;
  threefry2x32((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)(b * (uint32_t)N + t), (uint32_t)(layer * vc::PRENET + unit), o0, o1);
  return (o0 < 0x80000000u) ? 2.f : 0.f;
}

// ---- small kernels ---------------------------------------------------------------------------------
__global__ void embed_kernel(const int32_t* __restrict__ tokens, const float* __restrict__ emb, float* __restrict__ out, int n_tok) {
  // out[tok][256] = emb[tokens[tok]][256]
  const int tok = blockIdx.x * 4 + threadIdx.x / 64;
  if (tok >= n_tok) return;
  int id = tokens[tok];
  id = id < 0 ? 0 : (id >= vc::VOCAB ? vc::VOCAB - 1 : id);
  const int c4 = (threadIdx.x % 64) * 4;
  *reinterpret_cast<float4*>(out + (size_t)tok * vc::ENC_D + c4) = __ldg(reinterpret_cast<const float4*>(emb + (size_t)id * vc::ENC_D + c4));
}

__global__ void bn_inv_kernel(const float* __restrict__ scale, const float* __restrict__ var, float* __restrict__ inv, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) inv[i] = scale[i] * rsqrtf(var[i] + 1e-5f);
}

// dst[c][r][q] = src[(row0 + r)*ld + (q / upc)*gate_stride + c*upc + (q % upc)]
__global__ void repack_cols_kernel(const float* __restrict__ src, int ld, int row0, int nrows, float* __restrict__ dst,
                                   int ncta, int ncols, int upc, int gate_stride) {
  const size_t total = (size_t)ncta * nrows * ncols;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    int q = idx % ncols;
    int r = (idx / ncols) % nrows;
    int c = idx / ((size_t)ncols * nrows);
    int col = (q / upc) * gate_stride + c * upc + (q % upc);
    dst[idx] = src[(size_t)(row0 + r) * ld + col];
  }
}

__global__ void fill_kernel(int32_t* p, int v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---- Gaussian upsampling (model.py:102-111) --------------------------------------------------------
// out[b,n,:] = sum_l softmax_l(-(mid_l - n)^2/10) enc[b,l,:],  mid = cumsum(dur) - dur/2
constexpr int UP_F = 8;  // frames per CTA
__global__ void __launch_bounds__(256) upsample_kernel(const float* __restrict__ enc, const float* __restrict__ dur,
                                                       const int32_t* __restrict__ lengths, const int32_t* __restrict__ n_frames,
                                                       int L, int N, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* mid = sm;            // [L]
  float* w = sm + L;          // [UP_F][L]
  const int b = blockIdx.y, f0 = blockIdx.x * UP_F, tid = threadIdx.x;
  const int len = lengths ? min(lengths[b], L) : L;
  const int nf = n_frames ? min(n_frames[b], N) : N;
  if (f0 >= nf) return;
  if (tid < 32) {
    // sequential cumsum in fp32 (jnp.cumsum), done by one warp with a carried scan
    float carry = 0.f;
    for (int l0 = 0; l0 < len; l0 += 32) {
      int l = l0 + tid;
      float d = l < len ? dur[(size_t)b * L + l] : 0.f;
      float s = d;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float v = __shfl_up_sync(0xffffffffu, s, o);
        if (tid >= o) s += v;
      }
      s += carry;
      if (l < len) mid[l] = s - d / 2.f;
      carry = __shfl_sync(0xffffffffu, s, 31);
    }
  }
  __syncthreads();
  const int warp = tid / 32, lane = tid % 32;  // 8 warps = UP_F frames
  {
    const float n = (float)(f0 + warp);
    float mx = -INFINITY;
    for (int l = lane; l < len; l += 32) {
      float d = mid[l] - n;
      float v = -(d * d) / 10.0f;
      w[warp * L + l] = v;
      mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int l = lane; l < len; l += 32) {
      float e = expf(w[warp * L + l] - mx);
      w[warp * L + l] = e;
      s += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    for (int l = lane; l < len; l += 32) w[warp * L + l] = w[warp * L + l] / s;  // jax.nn.softmax: exp / sum
  }
  __syncthreads();
  // each thread: 2 channels x UP_F frames
  float acc[UP_F][2];
#pragma unroll
  for (int f = 0; f < UP_F; ++f) acc[f][0] = acc[f][1] = 0.f;
  const float* e = enc + (size_t)b * L * vc::ENC_OUT + tid * 2;
  for (int l = 0; l < len; ++l) {
    const float2 x = __ldg(reinterpret_cast<const float2*>(e + (size_t)l * vc::ENC_OUT));
#pragma unroll
    for (int f = 0; f < UP_F; ++f) {
      const float ww = w[f * L + l];
      acc[f][0] = fmaf(ww, x.x, acc[f][0]);
      acc[f][1] = fmaf(ww, x.y, acc[f][1]);
    }
  }
#pragma unroll
  for (int f = 0; f < UP_F; ++f) {
    const int n = f0 + f;
    if (n < nf) {
      float2 o = make_float2(acc[f][0], acc[f][1]);
      *reinterpret_cast<float2*>(out + ((size_t)b * N + n) * vc::ENC_OUT + tid * 2) = o;
    }
  }
}

// ---- shared device pieces of the scan kernels -------------------------------------------------------
struct Seg {
  const float* p;   // p[row*stride + i]
  int n;            // multiple of 4
  int stride;
};

// Copy resident weights [K][16] from global (this CTA's slice) into shared memory with one pad row
// per K slice (slice length SL = K/64): physical row = k + k/SL, so that the two slices read in the
// same quarter-warp phase land in different bank halves.
__device__ void load_w16(float* wsm, const float* __restrict__ g, int K) {
  const int SL = K / NSLICE;
  for (int e = threadIdx.x; e < K * 4; e += SCAN_THREADS) {
    int k = e >> 2, q = (e & 3) * 4;
    float4 v = __ldg(reinterpret_cast<const float4*>(g + (size_t)k * NCOL + q));
    *reinterpret_cast<float4*>(wsm + (size_t)(k + k / SL) * NCOL + q) = v;
  }
}

// Stage `nr` (<= RG) batch rows of the concatenated input vector into xs[r][Kpad].
// zero_row[r] != 0 zeroes the LAST segment of that row (ResetCore on the recurrent state).
__device__ void stage_rows(float* xs, int Kpad, const Seg* segs, int nseg, int row0, int nr, const int* zero_last) {
  int koff = 0;
  for (int s = 0; s < nseg; ++s) {
    const int n4 = segs[s].n >> 2;
    for (int e = threadIdx.x; e < nr * n4; e += SCAN_THREADS) {
      int r = e / n4, i4 = (e - r * n4) * 4;
      float4 v = make_float4(0, 0, 0, 0);
      const bool z = segs[s].p == nullptr || (zero_last && s == nseg - 1 && zero_last[r]);
      if (!z) v = __ldcg(reinterpret_cast<const float4*>(segs[s].p + (size_t)(row0 + r) * segs[s].stride + i4));
      *reinterpret_cast<float4*>(xs + (size_t)r * Kpad + koff + i4) = v;
    }
    koff += segs[s].n;
  }
  for (int e = threadIdx.x; e < (RG - nr) * (koff >> 2); e += SCAN_THREADS) {
    int r = nr + e / (koff >> 2), i4 = (e % (koff >> 2)) * 4;
    *reinterpret_cast<float4*>(xs + (size_t)r * Kpad + i4) = make_float4(0, 0, 0, 0);
  }
}

// z[r][col] (RG x 16) = xs[r][0:K] . wsm[0:K][col]; result written to zs[r*16 + col].
// Thread (ks = tid/4, cgp = tid%4) reduces K slice ks for 4 columns x 8 rows.
__device__ void matmul16(const float* __restrict__ xs, int Kpad, const float* __restrict__ wsm, int K, float* part, float* zs) {
  const int tid = threadIdx.x;
  const int cgp = tid & 3, ks = tid >> 2;
  const int SL = K / NSLICE;
  float acc[RG][4];
#pragma unroll
  for (int r = 0; r < RG; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
  const float* xk = xs + ks * SL;
  const float* wk = wsm + (size_t)(ks * SL + ks) * NCOL + cgp * 4;
  for (int kk = 0; kk < SL; kk += 4) {
    float4 xv[RG];
#pragma unroll
    for (int r = 0; r < RG; ++r) xv[r] = *reinterpret_cast<const float4*>(xk + (size_t)r * Kpad + kk);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float4 wv = *reinterpret_cast<const float4*>(wk + (size_t)(kk + e) * NCOL);
#pragma unroll
      for (int r = 0; r < RG; ++r) {
        const float x = e == 0 ? xv[r].x : (e == 1 ? xv[r].y : (e == 2 ? xv[r].z : xv[r].w));
        acc[r][0] = fmaf(x, wv.x, acc[r][0]);
        acc[r][1] = fmaf(x, wv.y, acc[r][1]);
        acc[r][2] = fmaf(x, wv.z, acc[r][2]);
        acc[r][3] = fmaf(x, wv.w, acc[r][3]);
      }
    }
  }
  // reduce the 8 slices held by one warp (lane bits 2..4), then across the 8 warps through smem
#pragma unroll
  for (int r = 0; r < RG; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v = acc[r][c];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      acc[r][c] = v;
    }
  const int warp = tid >> 5, lane = tid & 31;
  if (lane < 4) {
#pragma unroll
    for (int r = 0; r < RG; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) part[warp * (RG * NCOL) + r * NCOL + lane * 4 + c] = acc[r][c];
  }
  __syncthreads();
  if (tid < RG * NCOL) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) s += part[w * (RG * NCOL) + tid];
    zs[tid] = s;
  }
  __syncthreads();
}

// LSTM cell update for (row r, unit uu) from zs (hk.LSTM: i,g,f,o; forget bias +1)
__device__ __forceinline__ float lstm_cell(const float* zs, int r, int uu, float zadd_i, float zadd_g, float zadd_f, float zadd_o, float& c) {
  const float zi = zs[r * NCOL + 0 * UPC + uu] + zadd_i;
  const float zg = zs[r * NCOL + 1 * UPC + uu] + zadd_g;
  const float zf = zs[r * NCOL + 2 * UPC + uu] + zadd_f;
  const float zo = zs[r * NCOL + 3 * UPC + uu] + zadd_o;
  const float f = sigmoidf_(zf + 1.f);
  c = f * c + sigmoidf_(zi) * tanhf(zg);
  return sigmoidf_(zo) * tanhf(c);
}

// ---- encoder BiLSTM scan (model.py:36-46) ---------------------------------------------------------
struct EncScanArgs {
  const float* zx;        // [2][B][L][1024] hoisted x.Wx + b per direction
  const float* whr;       // [2][64][256][16] recurrent weights, per direction / CTA
  const int32_t* lengths; // [B] or null
  float* out;             // [B][L][512]  (fwd | bwd)
  int B, L;
};

__global__ void __launch_bounds__(SCAN_THREADS, 1) enc_scan_kernel(const EncScanArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float sm[];
  constexpr int K = vc::ENC_D, H = vc::ENC_D, Kpad = K + 4;
  float* wsm = sm;                               // [(K+64)][16]
  float* xs = wsm + (K + NSLICE) * NCOL;         // [RG][Kpad]
  float* part = xs + RG * Kpad;                  // [8][RG*16]
  float* zs = part + 8 * RG * NCOL;              // [RG*16]
  float* cst = zs + RG * NCOL;                   // [MAX_ROWS][UPC]
  __shared__ int zero_row[RG];
  const int dir = blockIdx.x / 64, c = blockIdx.x % 64, tid = threadIdx.x;
  const int B = a.B, L = a.L;
  load_w16(wsm, a.whr + ((size_t)dir * 64 + c) * K * NCOL, K);
  for (int e = tid; e < MAX_ROWS * UPC; e += SCAN_THREADS) cst[e] = 0.f;
  __syncthreads();
  for (int s = 0; s < L; ++s) {
    const int t = dir == 0 ? s : L - 1 - s;
    const int tprev = dir == 0 ? t - 1 : t + 1;
    for (int row0 = 0; row0 < B; row0 += RG) {
      const int nr = min(RG, B - row0);
      if (tid < RG) {
        int z = 0;
        if (tid < nr) {
          const int len = a.lengths ? min(a.lengths[row0 + tid], L) : L;
          // fwd: zero state at t=0.  bwd: ResetCore mask = (t >= len-1) (model.py:37,40)
          z = dir == 0 ? (s == 0) : (s == 0 || t >= len - 1);
        }
        zero_row[tid] = z;
      }
      __syncthreads();
      Seg seg;
      seg.p = (s == 0) ? nullptr : a.out + (size_t)tprev * vc::ENC_OUT + dir * H;
      seg.n = H;
      seg.stride = L * vc::ENC_OUT;
      stage_rows(xs, Kpad, &seg, 1, row0, nr, zero_row);
      __syncthreads();
      matmul16(xs, Kpad, wsm, K, part, zs);
      if (tid < nr * UPC) {
        const int r = tid / UPC, uu = tid % UPC, b = row0 + r;
        const float* zx = a.zx + (((size_t)dir * B + b) * L + t) * (4 * H) + c * UPC + uu;
        float cc = zero_row[r] ? 0.f : cst[b * UPC + uu];
        const float h = lstm_cell(zs, r, uu, __ldg(zx), __ldg(zx + H), __ldg(zx + 2 * H), __ldg(zx + 3 * H), cc);
        cst[b * UPC + uu] = cc;
        a.out[((size_t)b * L + t) * vc::ENC_OUT + dir * H + c * UPC + uu] = h;
      }
      __syncthreads();
    }
    grid.sync();
  }
}

// ---- autoregressive decoder scan (model.py:129-142) ------------------------------------------------
// Per frame t (B rows at once), four grid-wide phases:
//   EA  p1(t) = drop(relu(mel_{t-1} . W1))  computed as relu([h0,h1]_{t-1} . (Wo.W1) + bo.W1)  (2 cols / CTA)
//       and mel_{t-1} = [h0,h1]_{t-1} . Wo + bo  (the output frame, CTAs 0..79)
//   B   p2(t) = drop(relu(p1 . W2))                                                        (2 cols / CTA)
//   C   LSTM0: z = zc0[t] + [p2, h0_{t-1}] . W0r ; gates -> h0_t                             (4 units / CTA)
//   D   LSTM1: z = zc1[t] + [p2, h0_t, h1_{t-1}] . W1r ; gates -> h1_t
// The CTA's slices of the recurrent matrices (768x16 and 1280x16 fp32 = 128 KB) live in REGISTERS:
// thread (ks, cg) owns K-slice ks (12 / 20 rows) x 4 gate columns of both layers, so shared memory is
// free to hold the input vectors of all 32 batch rows and each phase costs a single L2 round trip.
struct DecScanArgs {
  const float* zc0;      // [B][N][2048] cond.W0[0:512] + b0
  const float* zc1;      // [B][N][2048] cond.W1[0:512] + b1
  const float* w0r;      // [128][768][16]   rows 512..1279 of lstm0   ([p2, h0])
  const float* w1r;      // [128][1280][16]  rows 512..1791 of lstm1   ([p2, h0, h1])
  const float* wc;       // [32][1024][8]    (Wo . W1) columns, 8 per EA CTA
  const float* bc;       // [256]            bo . W1
  const float* wp2;      // [128][256][2]    prenet fc2 columns
  const float* wo;       // [20][1024][4]    projection columns, 4 per EA CTA
  const float* bo;       // [80]
  const uint8_t* keep;   // [B][N][2][256] or null
  uint64_t seed;
  int mode;
  float* p1;             // [B][256]
  float* p2;             // [B][256]
  float* h0;             // [2][B][512]
  float* h1;             // [2][B][512]
  float* mel;            // [B][N][80]  (pre-postnet output)
  int B, N;
};

constexpr int DEC_XR = 32;                 // batch rows staged at once
constexpr int DEC_KPAD = vc::PRENET + 2 * vc::DEC_H + 4;   // 1284

// stage rows [row0,row0+nr) of up to 3 concatenated segments into xs[r][DEC_KPAD]
__device__ __forceinline__ void dec_stage(float* xs, const Seg* segs, int nseg, int row0, int nr) {
  int koff = 0;
  for (int s = 0; s < nseg; ++s) {
    const int n4 = segs[s].n >> 2;
    const int total = nr * n4;
    const float* p = segs[s].p;
    const int stride = segs[s].stride;
    for (int e0 = threadIdx.x; e0 < total; e0 += SCAN_THREADS * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = e0 + u * SCAN_THREADS;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < total && p) {
          const int r = e / n4, i4 = (e - r * n4) * 4;
          v[u] = __ldcg(reinterpret_cast<const float4*>(p + (size_t)(row0 + r) * stride + i4));
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = e0 + u * SCAN_THREADS;
        if (e < total) {
          const int r = e / n4, i4 = (e - r * n4) * 4;
          *reinterpret_cast<float4*>(xs + (size_t)r * DEC_KPAD + koff + i4) = v[u];
        }
      }
    }
    koff += segs[s].n;
  }
}

// acc[8 rows][4 cols] partial sums over this thread's K slice -> butterfly over the 8 slices of the
// warp (28 shuffles); afterwards the thread holds the warp-level sums of row (lane>>2), cols cg*4..+3.
__device__ __forceinline__ float4 warp_reduce_rows(float (&acc)[RG][4], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float a4[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float send = b4 ? acc[r][c] : acc[r + 4][c];
      const float keep = b4 ? acc[r + 4][c] : acc[r][c];
      a4[r][c] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  float a2[2][4];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float send = b3 ? a4[r][c] : a4[r + 2][c];
      const float keep = b3 ? a4[r + 2][c] : a4[r][c];
      a2[r][c] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  float a1[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float send = b2 ? a2[0][c] : a2[1][c];
    const float keep = b2 ? a2[1][c] : a2[0][c];
    a1[c] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  return make_float4(a1[0], a1[1], a1[2], a1[3]);
}

// one LSTM layer for up to DEC_XR staged rows: z partials -> part[warp][row][16] -> zs[row][16]
template <int SL>
__device__ __forceinline__ void dec_matmul(const float* __restrict__ xs, int koff_unused, const float (&w)[SL][4], int ngroups,
                                           float* part, float* zs) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ks = tid >> 2;
  for (int g = 0; g < ngroups; ++g) {
    float acc[RG][4];
#pragma unroll
    for (int r = 0; r < RG; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    const float* xk = xs + (size_t)(g * RG) * DEC_KPAD + ks * SL;
#pragma unroll
    for (int kk = 0; kk < SL; kk += 4) {
#pragma unroll
      for (int r = 0; r < RG; ++r) {
        const float4 xv = *reinterpret_cast<const float4*>(xk + (size_t)r * DEC_KPAD + kk);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc[r][c] = fmaf(xv.x, w[kk + 0][c], acc[r][c]);
          acc[r][c] = fmaf(xv.y, w[kk + 1][c], acc[r][c]);
          acc[r][c] = fmaf(xv.z, w[kk + 2][c], acc[r][c]);
          acc[r][c] = fmaf(xv.w, w[kk + 3][c], acc[r][c]);
        }
      }
    }
    const float4 s = warp_reduce_rows(acc, lane);
    // row (lane>>2) of group g, columns (lane&3)*4..+3
    *reinterpret_cast<float4*>(part + ((size_t)warp * DEC_XR + g * RG + (lane >> 2)) * NCOL + (lane & 3) * 4) = s;
  }
  __syncthreads();
  for (int o = tid; o < ngroups * RG * NCOL; o += SCAN_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < SCAN_THREADS / 32; ++wv) s += part[(size_t)wv * DEC_XR * NCOL + o];
    zs[o] = s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(SCAN_THREADS, 1) decoder_scan_kernel(const DecScanArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float sm[];
  constexpr int H = vc::DEC_H, K0 = vc::PRENET + H, K1 = vc::PRENET + 2 * H;
  constexpr int SL0 = K0 / NSLICE, SL1 = K1 / NSLICE;   // 12, 20
  float* xs = sm;                                   // [DEC_XR][DEC_KPAD]
  float* part = xs + DEC_XR * DEC_KPAD;             // [8][DEC_XR][16]
  float* zs = part + 8 * DEC_XR * NCOL;             // [DEC_XR][16]
  float* cst = zs + DEC_XR * NCOL;                  // [2][MAX_ROWS][UPC]
  float* wea = cst + 2 * MAX_ROWS * UPC;            // [1024][8]: CTAs 0..31 (Wo.W1) cols 8c..8c+7; CTAs 32..51 Wo cols 4(c-32)..+3 ([1024][4])
  float* wp2 = wea + 2 * H * 8;                     // [256][2]
  const int c = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ks = tid >> 2, cgp = tid & 3;
  const int B = a.B, N = a.N;

  // ---- register-resident recurrent weight slices ----
  float w0[SL0][4], w1[SL1][4];
  {
    const float* g0 = a.w0r + ((size_t)c * K0 + ks * SL0) * NCOL + cgp * 4;
#pragma unroll
    for (int i = 0; i < SL0; ++i) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g0 + (size_t)i * NCOL));
      w0[i][0] = v.x; w0[i][1] = v.y; w0[i][2] = v.z; w0[i][3] = v.w;
    }
    const float* g1 = a.w1r + ((size_t)c * K1 + ks * SL1) * NCOL + cgp * 4;
#pragma unroll
    for (int i = 0; i < SL1; ++i) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g1 + (size_t)i * NCOL));
      w1[i][0] = v.x; w1[i][1] = v.y; w1[i][2] = v.z; w1[i][3] = v.w;
    }
  }
  for (int e = tid; e < 2 * MAX_ROWS * UPC; e += SCAN_THREADS) cst[e] = 0.f;
  constexpr int EA_P1 = 32, EA_MEL = 20;            // CTAs computing p1 (8 cols each) / mel (4 cols each)
  if (c < EA_P1)
    for (int e = tid; e < 2 * H * 8; e += SCAN_THREADS) wea[e] = a.wc[(size_t)c * 2 * H * 8 + e];
  else if (c < EA_P1 + EA_MEL)
    for (int e = tid; e < 2 * H * 4; e += SCAN_THREADS) wea[e] = a.wo[(size_t)(c - EA_P1) * 2 * H * 4 + e];
  for (int e = tid; e < vc::PRENET * 2; e += SCAN_THREADS) wp2[e] = a.wp2[(size_t)c * vc::PRENET * 2 + e];
  __syncthreads();

  for (int t = 0; t <= N; ++t) {
    const int cur = t & 1, prv = cur ^ 1;
    // ---- phase EA: mel_{t-1} (CTAs 32..51, 4 cols each) and p1(t) (CTAs 0..31, 8 cols each) from [h0,h1]_{t-1} ----
    if (t > 0 && c < EA_P1 + EA_MEL) {
      const bool is_p1 = c < EA_P1;
      for (int row0 = 0; row0 < B; row0 += DEC_XR) {
        const int nr = min(DEC_XR, B - row0);
        Seg segs[2];
        segs[0] = Seg{a.h0 + (size_t)prv * B * H, H, H};
        segs[1] = Seg{a.h1 + (size_t)prv * B * H, H, H};
        dec_stage(xs, segs, 2, row0, nr);
        __syncthreads();
        for (int r = warp; r < nr; r += SCAN_THREADS / 32) {
          const float* x = xs + (size_t)r * DEC_KPAD;
          const int b = row0 + r;
          if (is_p1) {
            float s[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) s[q] = 0.f;
#pragma unroll 4
            for (int i = lane; i < 2 * H; i += 32) {
              const float xv = x[i];
              const float4 wa = *reinterpret_cast<const float4*>(wea + i * 8);
              const float4 wb = *reinterpret_cast<const float4*>(wea + i * 8 + 4);
              s[0] = fmaf(xv, wa.x, s[0]); s[1] = fmaf(xv, wa.y, s[1]); s[2] = fmaf(xv, wa.z, s[2]); s[3] = fmaf(xv, wa.w, s[3]);
              s[4] = fmaf(xv, wb.x, s[4]); s[5] = fmaf(xv, wb.y, s[5]); s[6] = fmaf(xv, wb.z, s[6]); s[7] = fmaf(xv, wb.w, s[7]);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
            }
            if (lane < 8 && t < N) {
              float v = s[0];
#pragma unroll
              for (int q = 1; q < 8; ++q) v = lane == q ? s[q] : v;
              const int u = c * 8 + lane;
              v = fmaxf(v + __ldg(a.bc + u), 0.f);
              a.p1[(size_t)b * vc::PRENET + u] = v * keep_scale(a.mode, a.keep, a.seed, b, t, N, 0, u);
            }
          } else {
            float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
            for (int i = lane; i < 2 * H; i += 32) {
              const float xv = x[i];
              const float4 wa = *reinterpret_cast<const float4*>(wea + i * 4);
              s[0] = fmaf(xv, wa.x, s[0]); s[1] = fmaf(xv, wa.y, s[1]); s[2] = fmaf(xv, wa.z, s[2]); s[3] = fmaf(xv, wa.w, s[3]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
            }
            if (lane < 4) {
              float v = s[0];
#pragma unroll
              for (int q = 1; q < 4; ++q) v = lane == q ? s[q] : v;
              const int m = (c - EA_P1) * 4 + lane;
              a.mel[((size_t)b * N + (t - 1)) * vc::MEL + m] = v + __ldg(a.bo + m);
            }
          }
        }
        __syncthreads();
      }
    } else if (t == 0) {
      for (int e = tid; e < B * 2; e += SCAN_THREADS) a.p1[(size_t)(e >> 1) * vc::PRENET + c * 2 + (e & 1)] = 0.f;  // prenet(0) = 0
    }
    if (t == N) break;
    grid.sync();
    // ---- phase B: p2 = dropout(relu(p1 . W2)) ----
    for (int b = warp; b < B; b += SCAN_THREADS / 32) {
      float s0 = 0.f, s1 = 0.f;
      const float* p = a.p1 + (size_t)b * vc::PRENET;
#pragma unroll
      for (int i = lane; i < vc::PRENET; i += 32) {
        const float x = __ldcg(p + i);
        s0 = fmaf(x, wp2[i * 2 + 0], s0);
        s1 = fmaf(x, wp2[i * 2 + 1], s1);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      }
      if (lane < 2) {
        const float v = fmaxf(lane == 0 ? s0 : s1, 0.f);
        const int u = c * 2 + lane;
        a.p2[(size_t)b * vc::PRENET + u] = v * keep_scale(a.mode, a.keep, a.seed, b, t, N, 1, u);
      }
    }
    grid.sync();
    // ---- phase C: LSTM0 on [cond_t (hoisted), p2, h0_prev] ----
    for (int row0 = 0; row0 < B; row0 += DEC_XR) {
      const int nr = min(DEC_XR, B - row0);
      Seg segs[2];
      segs[0] = Seg{a.p2, vc::PRENET, vc::PRENET};
      segs[1] = Seg{t == 0 ? nullptr : a.h0 + (size_t)prv * B * H, H, H};
      dec_stage(xs, segs, 2, row0, nr);
      __syncthreads();
      dec_matmul<SL0>(xs, 0, w0, (nr + RG - 1) / RG, part, zs);
      if (tid < nr * UPC) {
        const int r = tid / UPC, uu = tid % UPC, b = row0 + r;
        const float* zc = a.zc0 + ((size_t)b * N + t) * (4 * H) + c * UPC + uu;
        float cc = cst[b * UPC + uu];
        const float h = lstm_cell(zs, r, uu, __ldg(zc), __ldg(zc + H), __ldg(zc + 2 * H), __ldg(zc + 3 * H), cc);
        cst[b * UPC + uu] = cc;
        a.h0[((size_t)cur * B + b) * H + c * UPC + uu] = h;
      }
      __syncthreads();
    }
    grid.sync();
    // ---- phase D: LSTM1 on [cond_t (hoisted), p2, h0_t, h1_prev] ----
    for (int row0 = 0; row0 < B; row0 += DEC_XR) {
      const int nr = min(DEC_XR, B - row0);
      Seg segs[3];
      segs[0] = Seg{a.p2, vc::PRENET, vc::PRENET};
      segs[1] = Seg{a.h0 + (size_t)cur * B * H, H, H};
      segs[2] = Seg{t == 0 ? nullptr : a.h1 + (size_t)prv * B * H, H, H};
      dec_stage(xs, segs, 3, row0, nr);
      __syncthreads();
      dec_matmul<SL1>(xs, 0, w1, (nr + RG - 1) / RG, part, zs);
      if (tid < nr * UPC) {
        const int r = tid / UPC, uu = tid % UPC, b = row0 + r;
        const float* zc = a.zc1 + ((size_t)b * N + t) * (4 * H) + c * UPC + uu;
        float cc = cst[(MAX_ROWS + b) * UPC + uu];
        const float h = lstm_cell(zs, r, uu, __ldg(zc), __ldg(zc + H), __ldg(zc + 2 * H), __ldg(zc + 3 * H), cc);
        cst[(MAX_ROWS + b) * UPC + uu] = cc;
        a.h1[((size_t)cur * B + b) * H + c * UPC + uu] = h;
      }
      __syncthreads();
    }
    grid.sync();
  }
}

// Wc[i][o] = sum_m Wo[i][m] * W1[m][o]   (1024 x 80) . (80 x 256), double accumulation; bc = bo . W1
__global__ void precompose_kernel(const float* __restrict__ wo, const float* __restrict__ bo, const float* __restrict__ w1,
                                  float* __restrict__ wc, float* __restrict__ bc) {
  const int o = threadIdx.x;            // 256
  const int i = blockIdx.x;             // 0..1024 (last block computes bc)
  double s = 0.0;
  if (i < 2 * vc::DEC_H) {
    for (int m = 0; m < vc::MEL; ++m) s += (double)wo[(size_t)i * vc::MEL + m] * (double)w1[(size_t)m * vc::PRENET + o];
    wc[(size_t)i * vc::PRENET + o] = (float)s;
  } else {
    for (int m = 0; m < vc::MEL; ++m) s += (double)bo[m] * (double)w1[(size_t)m * vc::PRENET + o];
    bc[o] = (float)s;
  }
}

constexpr size_t enc_scan_smem() {
  return ((size_t)(vc::ENC_D + NSLICE) * NCOL + RG * (vc::ENC_D + 4) + 8 * RG * NCOL + RG * NCOL + MAX_ROWS * UPC) * 4;
}
constexpr size_t dec_scan_smem() {
  return ((size_t)DEC_XR * DEC_KPAD + 8 * DEC_XR * NCOL + DEC_XR * NCOL + 2 * MAX_ROWS * UPC + 2 * vc::DEC_H * 8 + vc::PRENET * 2) * 4;
}

// derived-weight slots (ctx->ac_d)
enum {
  D_ENC_BNINV0 = 0, D_ENC_BNINV1, D_ENC_BNINV2,
  D_POST_BNINV0, D_POST_BNINV1, D_POST_BNINV2, D_POST_BNINV3,
  D_ENC_WHR,     // [2][64][256][16]
  D_DEC_W0R,     // [128][768][16]
  D_DEC_W1R,     // [128][1280][16]
  D_DEC_WC,      // [32][1024][8]  (Wo . W1) columns
  D_DEC_WCFULL,  // [1024][256] scratch
  D_DEC_BC,      // [256]
  D_DEC_WP2,     // [128][256][2]
  D_DEC_WO,      // [20][1024][4]
  D_COUNT
};

}  // namespace

int vtts_acoustic_prepare(vtts_ctx* ctx) {
  const size_t sizes[D_COUNT] = {256, 256, 256, 512, 512, 512, 512,
                                 (size_t)2 * 64 * 256 * 16, (size_t)128 * 768 * 16, (size_t)128 * 1280 * 16,
                                 (size_t)128 * 1024 * 2, (size_t)1024 * 256, 256, (size_t)128 * 256 * 2, (size_t)80 * 1024};
  size_t total = 0;
  std::vector<size_t> offs(D_COUNT);
  for (int i = 0; i < D_COUNT; ++i) {
    offs[i] = total;
    total += (sizes[i] + 63) & ~size_t(63);
  }
  if (ctx->ac_derived) cudaFree(ctx->ac_derived);
  VTTS_CUDA(cudaMalloc(&ctx->ac_derived, total * sizeof(float)));
  ctx->ac_d.resize(D_COUNT);
  for (int i = 0; i < D_COUNT; ++i) ctx->ac_d[i] = ctx->ac_derived + offs[i];
  auto& T = ctx->ac_t;
  for (int i = 0; i < 3; ++i) bn_inv_kernel<<<1, 256>>>(T[aci::ENC_CONV(i, 2)], T[aci::ENC_CONV(i, 5)], ctx->ac_d[D_ENC_BNINV0 + i], 256);
  for (int i = 0; i < 4; ++i) bn_inv_kernel<<<2, 256>>>(T[aci::POST_CONV(i, 2)], T[aci::POST_CONV(i, 5)], ctx->ac_d[D_POST_BNINV0 + i], 512);
  // encoder recurrent weights: rows 256..511 of w[512][1024]
  repack_cols_kernel<<<256, 256>>>(T[aci::ENC_LSTM_F_W], 1024, 256, 256, ctx->ac_d[D_ENC_WHR], 64, 16, UPC, 256);
  repack_cols_kernel<<<256, 256>>>(T[aci::ENC_LSTM_B_W], 1024, 256, 256, ctx->ac_d[D_ENC_WHR] + (size_t)64 * 256 * 16, 64, 16, UPC, 256);
  // decoder: rows after the 512 cond rows
  repack_cols_kernel<<<512, 256>>>(T[aci::DEC_L0_W], 2048, 512, 768, ctx->ac_d[D_DEC_W0R], 128, 16, UPC, 512);
  repack_cols_kernel<<<512, 256>>>(T[aci::DEC_L1_W], 2048, 512, 1280, ctx->ac_d[D_DEC_W1R], 128, 16, UPC, 512);
  precompose_kernel<<<2 * vc::DEC_H + 1, 256>>>(T[aci::PROJ_W], T[aci::PROJ_B], T[aci::PRE1_W], ctx->ac_d[D_DEC_WCFULL], ctx->ac_d[D_DEC_BC]);
  repack_cols_kernel<<<256, 256>>>(ctx->ac_d[D_DEC_WCFULL], 256, 0, 1024, ctx->ac_d[D_DEC_WC], 32, 8, 8, 0);
  repack_cols_kernel<<<64, 256>>>(T[aci::PRE2_W], 256, 0, 256, ctx->ac_d[D_DEC_WP2], 128, 2, 2, 0);
  repack_cols_kernel<<<64, 256>>>(T[aci::PROJ_W], 80, 0, 1024, ctx->ac_d[D_DEC_WO], 20, 4, 4, 0);
  VTTS_CUDA(cudaGetLastError());
  VTTS_CUDA(cudaDeviceSynchronize());
  VTTS_CUDA(cudaFuncSetAttribute(enc_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc_scan_smem()));
  VTTS_CUDA(cudaFuncSetAttribute(decoder_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_scan_smem()));
  return VTTS_OK;
}

int vtts_acoustic_run(vtts_ctx* ctx, const int32_t* tokens, const int32_t* lengths, const float* dur,
                      const int32_t* n_frames, const uint8_t* keep, int mode, uint64_t seed, int B, int L, int N,
                      float* mel, cudaStream_t st, void* ws_base, size_t ws_cap, size_t* ws_need) {
  const bool measure = ws_need != nullptr;
  if (!measure) {
    if (!ctx->ac_loaded) return ctx->fail(VTTS_ERR_NOT_LOADED, "acoustic weights not loaded");
    if (B < 1 || L < 1 || N < 1 || B > MAX_ROWS)
      return ctx->fail(VTTS_ERR_BAD_ARG, "acoustic: B=%d L=%d N=%d (1 <= B <= %d rows per call; the host layer chunks larger batches)", B, L, N, MAX_ROWS);
    if (mode < 0 || mode > 2) return ctx->fail(VTTS_ERR_BAD_ARG, "acoustic: dropout_mode %d", mode);
    if (mode == VTTS_DROPOUT_MASK && !keep) return ctx->fail(VTTS_ERR_BAD_ARG, "acoustic: dropout_mode MASK needs keep_mask");
    if (ctx->sm_count < SCAN_CTAS) return ctx->fail(VTTS_ERR_NO_DEVICE, "scan kernels need %d SMs, device has %d", SCAN_CTAS, ctx->sm_count);
  }
  Arena ar(ws_base, ws_cap, measure);
  const size_t BL = (size_t)B * L, BN = (size_t)B * N;
  float* e0 = ar.take<float>(BL * 256);
  float* e1 = ar.take<float>(BL * 256);
  float* zx = ar.take<float>(2 * BL * 1024);
  float* enc = ar.take<float>(BL * 512);
  float* cond = ar.take<float>(BN * 512);
  float* zc0 = ar.take<float>(BN * 2048);
  float* zc1 = ar.take<float>(BN * 2048);
  float* melpre = ar.take<float>(BN * 80);
  float* q0 = ar.take<float>(BN * 512);
  float* q1 = ar.take<float>(BN * 512);
  float* p1 = ar.take<float>((size_t)B * 256);
  float* p2 = ar.take<float>((size_t)B * 256);
  float* h0 = ar.take<float>((size_t)2 * B * 512);
  float* h1 = ar.take<float>((size_t)2 * B * 512);
  if (measure) {
    *ws_need = ar.off + 256;
    return VTTS_OK;
  }
  auto& T = ctx->ac_t;
  auto& D = ctx->ac_d;
  ctx->tap_enc = enc; ctx->tap_enc_n = BL * 512;
  ctx->tap_cond = cond; ctx->tap_cond_n = BN * 512;
  ctx->tap_melpre = melpre; ctx->tap_melpre_n = BN * 80;

  VTTS_CUDA(cudaMemsetAsync(mel, 0, BN * 80 * sizeof(float), st));
  VTTS_CUDA(cudaMemsetAsync(cond, 0, BN * 512 * sizeof(float), st));
  VTTS_CUDA(cudaMemsetAsync(melpre, 0, BN * 80 * sizeof(float), st));

  // ---- TokenEncoder: embed -> 3 x [conv k3, BN(eval), relu] ----
  embed_kernel<<<(unsigned)((BL + 3) / 4), 256, 0, st>>>(tokens, T[aci::EMBED], e0, (int)BL);
  ctx->launches++;
  VTTS_CUDA(cudaGetLastError());
  ConvLaunch Lc;
  float* cur = e0;
  float* nxt = e1;
  for (int i = 0; i < 3; ++i) {
    memset(&Lc, 0, sizeof(Lc));
    Lc.nprob = 1; Lc.Cin = 256; Lc.Cout = 256; Lc.B = B; Lc.T_rows = L; Lc.rows_out = L;
    Lc.len = lengths; Lc.len_mul = 1; Lc.pre_mode = 0; Lc.pre_slope = 1.f; Lc.post_act = 2;
    Lc.p[0] = ConvProb{cur, nullptr, nullptr, T[aci::ENC_CONV(i, 0)], T[aci::ENC_CONV(i, 1)], nullptr,
                       T[aci::ENC_CONV(i, 4)], D[D_ENC_BNINV0 + i], T[aci::ENC_CONV(i, 3)], nxt, 3, 1, -1, 1, 0};
    int rc = vtts_launch_conv(ctx, Lc, st);
    if (rc) return rc;
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
  // rows past len[b] of `cur` were never written: the scans mask them, but the hoisted GEMM reads them
  // -> harmless garbage confined to rows that are never consumed (k=1 GEMM has no row mixing).
  // ---- hoisted input projections of the two encoder LSTMs: zx[dir] = x . W[0:256] + b ----
  memset(&Lc, 0, sizeof(Lc));
  Lc.nprob = 2; Lc.Cin = 256; Lc.Cout = 1024; Lc.B = 1; Lc.T_rows = (int)BL; Lc.rows_out = (int)BL;
  Lc.len = nullptr; Lc.len_mul = 1; Lc.pre_mode = 0; Lc.pre_slope = 1.f; Lc.post_act = 0;
  Lc.p[0] = ConvProb{cur, nullptr, nullptr, T[aci::ENC_LSTM_F_W], T[aci::ENC_LSTM_F_B], nullptr, nullptr, nullptr, nullptr, zx, 1, 1, 0, 1, 0};
  Lc.p[1] = ConvProb{cur, nullptr, nullptr, T[aci::ENC_LSTM_B_W], T[aci::ENC_LSTM_B_B], nullptr, nullptr, nullptr, nullptr, zx + BL * 1024, 1, 1, 0, 1, 0};
  int rc = vtts_launch_conv(ctx, Lc, st);
  if (rc) return rc;
  // ---- BiLSTM scan ----
  {
    EncScanArgs ea;
    ea.zx = zx; ea.whr = D[D_ENC_WHR]; ea.lengths = lengths; ea.out = enc; ea.B = B; ea.L = L;
    void* args[] = {&ea};
    VTTS_CUDA(cudaLaunchCooperativeKernel((void*)enc_scan_kernel, dim3(SCAN_CTAS), dim3(SCAN_THREADS), args, enc_scan_smem(), st));
    ctx->launches++;
  }
  // ---- Gaussian upsampling ----
  {
    dim3 grid((N + UP_F - 1) / UP_F, B);
    size_t smem = (size_t)(L + UP_F * L) * sizeof(float);
    if (smem > 200 * 1024) return ctx->fail(VTTS_ERR_BAD_ARG, "acoustic: L=%d too long for the upsample kernel", L);
    if (smem > 48 * 1024) VTTS_CUDA(cudaFuncSetAttribute(upsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    upsample_kernel<<<grid, 256, smem, st>>>(enc, dur, lengths, n_frames, L, N, cond);
    ctx->launches++;
    VTTS_CUDA(cudaGetLastError());
  }
  // ---- hoisted cond projections of the decoder LSTMs ----
  memset(&Lc, 0, sizeof(Lc));
  Lc.nprob = 2; Lc.Cin = 512; Lc.Cout = 2048; Lc.B = 1; Lc.T_rows = (int)BN; Lc.rows_out = (int)BN;
  Lc.pre_mode = 0; Lc.pre_slope = 1.f; Lc.post_act = 0; Lc.len_mul = 1;
  Lc.p[0] = ConvProb{cond, nullptr, nullptr, T[aci::DEC_L0_W], T[aci::DEC_L0_B], nullptr, nullptr, nullptr, nullptr, zc0, 1, 1, 0, 1, 0};
  Lc.p[1] = ConvProb{cond, nullptr, nullptr, T[aci::DEC_L1_W], T[aci::DEC_L1_B], nullptr, nullptr, nullptr, nullptr, zc1, 1, 1, 0, 1, 0};
  rc = vtts_launch_conv(ctx, Lc, st);
  if (rc) return rc;
  // ---- autoregressive scan ----
  {
    DecScanArgs da;
    da.zc0 = zc0; da.zc1 = zc1;
    da.w0r = D[D_DEC_W0R]; da.w1r = D[D_DEC_W1R]; da.wc = D[D_DEC_WC]; da.bc = D[D_DEC_BC]; da.wp2 = D[D_DEC_WP2];
    da.wo = D[D_DEC_WO]; da.bo = T[aci::PROJ_B];
    da.keep = keep; da.seed = seed; da.mode = mode;
    da.p1 = p1; da.p2 = p2; da.h0 = h0; da.h1 = h1; da.mel = melpre;
    da.B = B; da.N = N;
    void* args[] = {&da};
    VTTS_CUDA(cudaLaunchCooperativeKernel((void*)decoder_scan_kernel, dim3(SCAN_CTAS), dim3(SCAN_THREADS), args, dec_scan_smem(), st));
    ctx->launches++;
  }
  // ---- postnet: 4 x [conv k5, BN, tanh], conv k5, + residual ----
  const float* pin = melpre;
  float* pout = q0;
  int cin = 80;
  for (int i = 0; i < 5; ++i) {
    const int cout = i < 4 ? 512 : 80;
    memset(&Lc, 0, sizeof(Lc));
    Lc.nprob = 1; Lc.Cin = cin; Lc.Cout = cout; Lc.B = B; Lc.T_rows = N; Lc.rows_out = N;
    Lc.len = n_frames; Lc.len_mul = 1; Lc.pre_mode = 0; Lc.pre_slope = 1.f; Lc.post_act = i < 4 ? 1 : 0;
    ConvProb p;
    memset(&p, 0, sizeof(p));
    p.x0 = pin; p.w = T[aci::POST_CONV(i, 0)]; p.bias = T[aci::POST_CONV(i, 1)];
    if (i < 4) { p.bn_mean = T[aci::POST_CONV(i, 4)]; p.bn_inv = D[D_POST_BNINV0 + i]; p.bn_off = T[aci::POST_CONV(i, 3)]; }
    if (i == 4) { p.resid = melpre; p.out = mel; } else { p.out = pout; }
    p.k = 5; p.dil = 1; p.in_off = -2; p.out_stride = 1; p.out_off = 0;
    Lc.p[0] = p;
    rc = vtts_launch_conv(ctx, Lc, st);
    if (rc) return rc;
    pin = pout;
    pout = (pout == q0) ? q1 : q0;
    cin = cout;
  }
  return VTTS_OK;
}
