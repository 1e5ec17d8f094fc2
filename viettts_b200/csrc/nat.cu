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
  // counter = (frame, row << 12 | entry): a row's stream depends on its row index in the call and the absolute frame,
  // not on the padded frame count N of the batch it happens to share
  (void)N;
  threefry2x32((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)t, ((uint32_t)b << 12) | (uint32_t)(layer * vc::PRENET + unit), o0, o1);
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

// ---- Gaussian upsampling (model.py:102-111) --------------------------------------------------------
// out[b,n,:] = sum_l softmax_l(-(mid_l - n)^2/10) enc[b,l,:],  mid = cumsum(dur) - dur/2
constexpr int UP_F = 8;  // frames per CTA
__global__ void __launch_bounds__(256) upsample_kernel(const float* __restrict__ enc, const float* __restrict__ dur,
                                                       const int32_t* __restrict__ lengths, const int32_t* __restrict__ n_frames,
                                                       int L, int N, float* __restrict__ out, int out_ld) {
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
      *reinterpret_cast<float2*>(out + ((size_t)b * N + n) * out_ld + tid * 2) = o;
    }
  }
}

// ---- shared device pieces of the scan kernels -------------------------------------------------------
struct Seg {
  const float* p;   // p[row*stride + i]
  int n;            // multiple of 4
  int stride;
};

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

// ---- autoregressive decoder scan (model.py:129-142) ------------------------------------------------
// One cooperative grid of 144 CTAs, up to 32 batch rows per launch.
//   CTAs 0..127   LSTM role: CTA c owns 4 hidden units (16 gate columns) of both layers; its slices of the
//                 recurrent matrices (768x16 + 1280x16 fp32) live in REGISTERS.  A persistent shared-memory
//                 buffer holds [p2 | h0 | h1] of all rows; the parts that are already final (h0_{t-1},
//                 h1_{t-1}) are prefetched while the prenet CTAs work, so only p2 (phase C) and h0_t (phase D)
//                 are fetched on the critical path.
//   CTAs 128..143 prenet role: 16 columns each of p1 = drop(relu([h0,h1]_{t-1}.(Wo.W1) + bo.W1))  (phase EA)
//                 and of p2 = drop(relu(p1.W2))                                                   (phase B)
// The output projection mel_t = [h0,h1]_t.Wo + bo is NOT part of the scan: nothing in the recurrence reads mel_t
// (the prenet consumes the precomposed Wo.W1), so every frame's [h0 | h1] is written to `hout` and one GEMM
// projects the whole sequence afterwards.  (A projection role inside the scan -- 4 CTAs taking part in every grid
// barrier -- was the slowest arrival at the C and D barriers for small batches: 15 us per frame at B = 1.)
// Four grid barriers per frame: EA | B | C (LSTM0) | D (LSTM1).
struct DecScanArgs {
  const float* zc0;      // [B][N][2048] cond.W0[0:512] + b0
  const float* zc1;      // [B][N][2048] cond.W1[0:512] + b1
  const float* w0r;      // [128][768][16]   rows 512..1279 of lstm0   ([p2, h0])
  const float* w1r;      // [128][1280][16]  rows 512..1791 of lstm1   ([p2, h0, h1])
  const float* wc;       // [16][1024][16]   (Wo . W1) columns, 16 per prenet CTA
  const float* bc;       // [256]            bo . W1
  const float* wp2;      // [16][256][16]    prenet fc2 columns
  const uint8_t* keep;   // [B][N][2][256] or null (indexed with row_base)
  uint64_t seed;
  int mode;
  float* p1;             // [B][256]
  float* p2;             // [B][256]
  float* h0;             // [2][B][512] compact double-buffered state the recurrence reads (frame parity)
  float* h1;             // [2][B][512]
  unsigned int* pre_bar; // arrival counter of the prenet CTAs' private barrier (zeroed before the launch)
  int* err;              // device int set before trapping on a barrier time-out
  float* hout;           // [B][N][1024] decoder outputs [h0_t | h1_t] of every frame (write only): the output projection
                         // runs over the whole tensor as ONE GEMM after the scan
  int B, N;              // rows of this launch (<= 32), frames
  int row_base;          // first row of this launch inside the full batch (dropout stream indexing)
  int N_total_rows;      // unused
  long long* dbg;        // optional per-CTA phase timers [grid][16]
};

constexpr int DEC_XR = 32;                 // batch rows per staging group (smem holds the state of one group)
constexpr int DEC_NG = 4;                  // row groups per launch: up to 128 rows share the three grid barriers of a frame
constexpr int DEC_KPAD = vc::PRENET + 2 * vc::DEC_H + 4;   // 1284: [p2 | h0 | h1] + pad
constexpr int DEC_LSTM = 128, DEC_PRE = 16;
constexpr int DEC_CTAS = DEC_LSTM + DEC_PRE;   // 144

// copy rows [0,nr) x [n floats] of a global matrix (row stride `stride`) into smem (row pitch `pitch`, column
// offset koff); p == nullptr writes zeros.  8 x 16 B loads in flight per thread.
__device__ __forceinline__ void dec_fetch(float* xs, int pitch, int koff, const float* p, int n, int stride, int nr) {
  const int n4 = n >> 2;
  const int total = nr * n4;
  for (int e0 = threadIdx.x; e0 < total; e0 += SCAN_THREADS * 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = e0 + u * SCAN_THREADS;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < total && p) {
        const int r = e / n4, i4 = (e - r * n4) * 4;
        v[u] = __ldcg(reinterpret_cast<const float4*>(p + (size_t)r * stride + i4));
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = e0 + u * SCAN_THREADS;
      if (e < total) {
        const int r = e / n4, i4 = (e - r * n4) * 4;
        *reinterpret_cast<float4*>(xs + (size_t)r * pitch + koff + i4) = v[u];
      }
    }
  }
}

// same copy with cp.async (16 B, L2 only): the data goes global -> shared without passing through registers, so a fetch
// can stay in flight while the CTA computes on operands that have already arrived (decoder_tf_scan_kernel)
__device__ __forceinline__ void dec_fetch_async(float* xs, int pitch, int koff, const float* p, int n, int stride, int nr) {
  const int n4 = n >> 2;
  const int total = nr * n4;
  for (int e = threadIdx.x; e < total; e += SCAN_THREADS) {
    const int r = e / n4, i4 = (e - r * n4) * 4;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(xs + (size_t)r * pitch + koff + i4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p + (size_t)r * stride + i4) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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

// partial LSTM pre-activations for the staged rows: zout[row][16] = xs[row][0:64*SL] . w  (+ zadd[row][16])
// (xs already points at the first column of the K segment; thread (ks, cg) owns rows ks*SL.. of the segment)
template <int SL, int PITCH = DEC_KPAD>
__device__ __forceinline__ void dec_matmul(const float* __restrict__ xs, const float (&w)[SL][4], int ngroups, float* part, float* zout,
                                           const float* zadd) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ks = tid >> 2;
  for (int g = 0; g < ngroups; ++g) {
    float acc[RG][4];
#pragma unroll
    for (int r = 0; r < RG; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    const float* xk = xs + (size_t)(g * RG) * PITCH + ks * SL;
#pragma unroll
    for (int kk = 0; kk < SL; kk += 4) {
#pragma unroll
      for (int r = 0; r < RG; ++r) {
        const float4 xv = *reinterpret_cast<const float4*>(xk + (size_t)r * PITCH + kk);
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
    *reinterpret_cast<float4*>(part + ((size_t)warp * DEC_XR + g * RG + (lane >> 2)) * NCOL + (lane & 3) * 4) = s;
  }
  __syncthreads();
  for (int o = tid; o < ngroups * RG * NCOL; o += SCAN_THREADS) {
    float s = zadd ? zadd[o] : 0.f;
#pragma unroll
    for (int wv = 0; wv < SCAN_THREADS / 32; ++wv) s += part[(size_t)wv * DEC_XR * NCOL + o];
    zout[o] = s;
  }
  __syncthreads();
}

// ---- encoder BiLSTM scan (model.py:36-46) ---------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS, 1) enc_scan_kernel(const EncScanArgs a) {
  // 128 CTAs: direction = cta / 64; CTA owns 4 hidden units (16 gate columns); its 256x16 slice of the recurrent
  // matrix lives in registers (4 rows x 4 columns per thread); h_{t-1} of up to 32 rows is staged per step.
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float sm[];
  constexpr int K = vc::ENC_D, H = vc::ENC_D, XP = K + 4, SL = K / NSLICE;   // 256, 256, 260, 4
  float* xs = sm;                                // [32][XP]
  float* part = xs + 32 * XP;                    // [8][32][16]
  float* zs = part + 8 * DEC_XR * NCOL;          // [32][16]
  float* cst = zs + DEC_XR * NCOL;               // [MAX_ROWS][UPC]
  __shared__ int zero_row[DEC_XR];
  const int dir = blockIdx.x / 64, c = blockIdx.x % 64, tid = threadIdx.x;
  const int ks = tid >> 2, cgp = tid & 3;
  const int B = a.B, L = a.L;
  float wh[SL][4];
  {
    const float* g = a.whr + (((size_t)dir * 64 + c) * K + ks * SL) * NCOL + cgp * 4;
#pragma unroll
    for (int i = 0; i < SL; ++i) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g + (size_t)i * NCOL));
      wh[i][0] = v.x; wh[i][1] = v.y; wh[i][2] = v.z; wh[i][3] = v.w;
    }
  }
  for (int e = tid; e < MAX_ROWS * UPC; e += SCAN_THREADS) cst[e] = 0.f;
  for (int e = tid; e < 32 * XP; e += SCAN_THREADS) xs[e] = 0.f;
  __syncthreads();
  for (int s = 0; s < L; ++s) {
    const int t = dir == 0 ? s : L - 1 - s;
    const int tprev = dir == 0 ? t - 1 : t + 1;
    for (int row0 = 0; row0 < B; row0 += DEC_XR) {
      const int nr = min(DEC_XR, B - row0);
      if (tid < DEC_XR) {
        int z = 1;
        if (tid < nr) {
          const int len = a.lengths ? min(a.lengths[row0 + tid], L) : L;
          // fwd: zero state at t=0.  bwd: ResetCore mask = (t >= len-1) (model.py:37,40)
          z = dir == 0 ? (s == 0) : (s == 0 || t >= len - 1);
        }
        zero_row[tid] = z;
      }
      __syncthreads();
      // stage h_{t-1} (zero for reset rows)
      for (int e = tid; e < nr * (H / 4); e += SCAN_THREADS) {
        const int r = e / (H / 4), i4 = (e - r * (H / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!zero_row[r]) v = __ldcg(reinterpret_cast<const float4*>(a.out + ((size_t)(row0 + r) * L + tprev) * vc::ENC_OUT + dir * H + i4));
        *reinterpret_cast<float4*>(xs + (size_t)r * XP + i4) = v;
      }
      __syncthreads();
      dec_matmul<SL, XP>(xs, wh, (nr + RG - 1) / RG, part, zs, nullptr);
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

// same, accumulating two K segments (xa: 64*SLA columns with weights wa; xb: 64*SLB columns with wb) before ONE reduction
template <int SLA, int SLB, int PITCH = DEC_KPAD>
__device__ __forceinline__ void dec_matmul2(const float* __restrict__ xa, const float (&wa)[SLA][4], const float* __restrict__ xb,
                                            const float (&wb)[SLB][4], int ngroups, float* part, float* zout, const float* zadd) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ks = tid >> 2;
  for (int g = 0; g < ngroups; ++g) {
    float acc[RG][4];
#pragma unroll
    for (int r = 0; r < RG; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    const float* xk = xa + (size_t)(g * RG) * PITCH + ks * SLA;
#pragma unroll
    for (int kk = 0; kk < SLA; kk += 4) {
#pragma unroll
      for (int r = 0; r < RG; ++r) {
        const float4 xv = *reinterpret_cast<const float4*>(xk + (size_t)r * PITCH + kk);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc[r][c] = fmaf(xv.x, wa[kk + 0][c], acc[r][c]);
          acc[r][c] = fmaf(xv.y, wa[kk + 1][c], acc[r][c]);
          acc[r][c] = fmaf(xv.z, wa[kk + 2][c], acc[r][c]);
          acc[r][c] = fmaf(xv.w, wa[kk + 3][c], acc[r][c]);
        }
      }
    }
    const float* xk2 = xb + (size_t)(g * RG) * PITCH + ks * SLB;
#pragma unroll
    for (int kk = 0; kk < SLB; kk += 4) {
#pragma unroll
      for (int r = 0; r < RG; ++r) {
        const float4 xv = *reinterpret_cast<const float4*>(xk2 + (size_t)r * PITCH + kk);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc[r][c] = fmaf(xv.x, wb[kk + 0][c], acc[r][c]);
          acc[r][c] = fmaf(xv.y, wb[kk + 1][c], acc[r][c]);
          acc[r][c] = fmaf(xv.z, wb[kk + 2][c], acc[r][c]);
          acc[r][c] = fmaf(xv.w, wb[kk + 3][c], acc[r][c]);
        }
      }
    }
    const float4 s = warp_reduce_rows(acc, lane);
    *reinterpret_cast<float4*>(part + ((size_t)warp * DEC_XR + g * RG + (lane >> 2)) * NCOL + (lane & 3) * 4) = s;
  }
  __syncthreads();
  for (int o = tid; o < ngroups * RG * NCOL; o += SCAN_THREADS) {
    float s = zadd ? zadd[o] : 0.f;
#pragma unroll
    for (int wv = 0; wv < SCAN_THREADS / 32; ++wv) s += part[(size_t)wv * DEC_XR * NCOL + o];
    zout[o] = s;
  }
  __syncthreads();
}

// ---- prenet / projection CTAs: out[32 rows][8 cols] = xs[32][K] . wsm[K][8] ------------------------------
// thread (rg = tid&3, ks = tid>>2): rows rg+4i (i<8), K slice ks of SL = K/64; weights in smem with one pad
// row per slice (physical row = k + k/SL).  Result: outv[row*8 + col] for 32 x 8 outputs.
template <int SL>
__device__ __forceinline__ void pre_gemm8(const float* __restrict__ xs, int pitch, const float* __restrict__ wsm, float* part, float* outv) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rg = tid & 3, ks = tid >> 2;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
  const float* xk = xs + (size_t)rg * pitch + ks * SL;
  const float* wk = wsm + (size_t)(ks * SL + ks) * 8;
#pragma unroll
  for (int kk = 0; kk < SL; kk += 4) {
    float4 xv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xv[i] = *reinterpret_cast<const float4*>(xk + (size_t)(4 * i) * pitch + kk);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float4 wa = *reinterpret_cast<const float4*>(wk + (size_t)(kk + e) * 8);
      const float4 wb = *reinterpret_cast<const float4*>(wk + (size_t)(kk + e) * 8 + 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x = e == 0 ? xv[i].x : (e == 1 ? xv[i].y : (e == 2 ? xv[i].z : xv[i].w));
        acc[i][0] = fmaf(x, wa.x, acc[i][0]); acc[i][1] = fmaf(x, wa.y, acc[i][1]);
        acc[i][2] = fmaf(x, wa.z, acc[i][2]); acc[i][3] = fmaf(x, wa.w, acc[i][3]);
        acc[i][4] = fmaf(x, wb.x, acc[i][4]); acc[i][5] = fmaf(x, wb.y, acc[i][5]);
        acc[i][6] = fmaf(x, wb.z, acc[i][6]); acc[i][7] = fmaf(x, wb.w, acc[i][7]);
      }
    }
  }
  // butterfly over the 8 K slices of the warp (lane bits 2..4), halving the row set each time
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float a4[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float send = b4 ? acc[i][c] : acc[i + 4][c];
      const float keep = b4 ? acc[i + 4][c] : acc[i][c];
      a4[i][c] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  float a2[2][8];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float send = b3 ? a4[i][c] : a4[i + 2][c];
      const float keep = b3 ? a4[i + 2][c] : a4[i][c];
      a2[i][c] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  float a1[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float send = b2 ? a2[0][c] : a2[1][c];
    const float keep = b2 ? a2[1][c] : a2[0][c];
    a1[c] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  // this thread now holds row rg + 4*(lane>>2 & 7) ... = rg + 4*ksl, all 8 columns, summed over the warp's slices
  const int row = rg + 4 * ((lane >> 2) & 7);
  float* pw = part + ((size_t)warp * 32 + row) * 8;
  *reinterpret_cast<float4*>(pw) = make_float4(a1[0], a1[1], a1[2], a1[3]);
  *reinterpret_cast<float4*>(pw + 4) = make_float4(a1[4], a1[5], a1[6], a1[7]);
  __syncthreads();
  {
    float sum = 0.f;
#pragma unroll
    for (int wv = 0; wv < SCAN_THREADS / 32; ++wv) sum += part[(size_t)wv * 256 + tid];
    outv[tid] = sum;   // tid = row*8 + col
  }
  __syncthreads();
}

// copy a [K][ncols_src] slice (cols c0..c0+8) of a global per-CTA weight block into smem [K + K/SL][8]
__device__ __forceinline__ void pre_load_w8(float* wsm, const float* g, int K, int SL, int ld, int c0) {
  for (int e = threadIdx.x; e < K * 2; e += SCAN_THREADS) {
    const int k = e >> 1, h = (e & 1) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(g + (size_t)k * ld + c0 + h));
    *reinterpret_cast<float4*>(wsm + (size_t)(k + k / SL) * 8 + h) = v;
  }
}

// Barrier among the DEC_PRE prenet CTAs only (all co-resident: cooperative launch).  The second prenet layer needs every
// column block of p1 from its 15 peers but nothing from the 128 LSTM CTAs, so a grid-wide barrier here would put the
// LSTM CTAs' pre-accumulation (6.2 us at 32 rows) on the critical path and cost a full grid.sync.  `target` is the
// monotonically growing arrival count; a 2 s time-out traps instead of hanging the GPU.
__device__ __forceinline__ void prenet_barrier(unsigned int* counter, unsigned int target, int* err) {
  __syncthreads();                                   // this CTA's p1 stores are issued
  if (threadIdx.x == 0) {
    __threadfence();                                 // ... and visible device-wide before the arrival is
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v < target && clock64() - t0 > 4000000000LL) {
        if (err) *err = 77;
        __threadfence_system();
        asm volatile("trap;");
      }
    } while (v < target);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(SCAN_THREADS, 1) decoder_scan_kernel(const DecScanArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float sm[];
  constexpr int H = vc::DEC_H, K0 = vc::PRENET + H, K1 = vc::PRENET + 2 * H;
  const int c = blockIdx.x, tid = threadIdx.x;
  const int B = a.B, N = a.N;
  long long tm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tq = clock64();
#define DEC_MARK(i) { const long long tn = clock64(); tm[i] += tn - tq; tq = tn; }

  if (c < DEC_LSTM) {
    // =============================== LSTM role ===============================
    float* xs = sm;                                   // [DEC_XR][DEC_KPAD] = [p2 | h0 | h1] of ONE row group at a time
    float* part = xs + DEC_XR * DEC_KPAD;             // [8][DEC_XR][16]
    float* zs = part + 8 * DEC_XR * NCOL;             // [DEC_XR][16]
    float* zp0 = zs + DEC_XR * NCOL;                  // [DEC_NG][DEC_XR][16] pre-accumulated h0_{t-1} . W0[h0 rows]
    float* zp1 = zp0 + DEC_NG * DEC_XR * NCOL;        // [DEC_NG][DEC_XR][16] pre-accumulated h1_{t-1} . W1[h1 rows]
    float* cst = zp1 + DEC_NG * DEC_XR * NCOL;        // [DEC_NG][2][DEC_XR][UPC]
    const int ks = tid >> 2, cgp = tid & 3;
    constexpr int SLP = vc::PRENET / NSLICE, SLH = H / NSLICE;   // 4, 8
    // register-resident weight slices, split by input segment so that each segment's product can be
    // accumulated as soon as that segment is final
    float w0p[SLP][4], w0h[SLH][4], w1p[SLP][4], w1h0[SLH][4], w1h1[SLH][4];
    {
      auto ld = [&](float (&dst)[4], const float* base, int row) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)row * NCOL + cgp * 4));
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
      };
      const float* g0 = a.w0r + (size_t)c * K0 * NCOL;
      const float* g1 = a.w1r + (size_t)c * K1 * NCOL;
#pragma unroll
      for (int i = 0; i < SLP; ++i) { ld(w0p[i], g0, ks * SLP + i); ld(w1p[i], g1, ks * SLP + i); }
#pragma unroll
      for (int i = 0; i < SLH; ++i) {
        ld(w0h[i], g0, vc::PRENET + ks * SLH + i);
        ld(w1h0[i], g1, vc::PRENET + ks * SLH + i);
        ld(w1h1[i], g1, vc::PRENET + H + ks * SLH + i);
      }
    }
    for (int e = tid; e < DEC_NG * 2 * DEC_XR * UPC; e += SCAN_THREADS) cst[e] = 0.f;
    for (int e = tid; e < DEC_NG * 2 * DEC_XR * NCOL; e += SCAN_THREADS) zp0[e] = 0.f;   // zp0 and zp1 are adjacent
    for (int e = tid; e < DEC_XR * DEC_KPAD; e += SCAN_THREADS) xs[e] = 0.f;    // rows >= B and the t=0 state are zero
    __syncthreads();
    // Rows are processed in groups of DEC_XR (the staging buffer holds one group); all groups of a launch share the
    // three grid barriers of a frame.  With one group, h0_{t-1} and p2 stay resident in xs between the phases.
    const int NG = (B + DEC_XR - 1) / DEC_XR;
    for (int t = 0; t < N; ++t) {
      // ---- EA window (prenet CTAs are busy): products of the state that is already final ----
      if (t > 0) {
        for (int rg = 0; rg < NG; ++rg) {
          const int r0 = rg * DEC_XR, nb = min(DEC_XR, B - r0), ngroups = (nb + RG - 1) / RG;
          if (NG > 1) dec_fetch(xs, DEC_KPAD, vc::PRENET, a.h0 + ((size_t)((t - 1) & 1) * B + r0) * H, H, H, nb);
          dec_fetch(xs, DEC_KPAD, vc::PRENET + H, a.h1 + ((size_t)((t - 1) & 1) * B + r0) * H, H, H, nb);
          if (NG > 1) __syncthreads();
          dec_matmul<SLH>(xs + vc::PRENET, w0h, ngroups, part, zp0 + rg * DEC_XR * NCOL, nullptr);     // its barriers also order the h1 fetch
          dec_matmul<SLH>(xs + vc::PRENET + H, w1h1, ngroups, part, zp1 + rg * DEC_XR * NCOL, nullptr);
        }
      }
      DEC_MARK(0)
      DEC_MARK(1)
      DEC_MARK(2)
      grid.sync();   // p2(t) is ready (the prenet CTAs order their two layers among themselves, see prenet_barrier)
      DEC_MARK(3)
      // ---- phase C: LSTM0 = zc0[t] + p2 . W0[p2 rows] + (h0_{t-1} part) ----
      for (int rg = 0; rg < NG; ++rg) {
        const int r0 = rg * DEC_XR, nb = min(DEC_XR, B - r0), ngroups = (nb + RG - 1) / RG;
        dec_fetch(xs, DEC_KPAD, 0, a.p2 + (size_t)r0 * vc::PRENET, vc::PRENET, vc::PRENET, nb);
        __syncthreads();
        dec_matmul<SLP>(xs, w0p, ngroups, part, zs, zp0 + rg * DEC_XR * NCOL);
        if (tid < nb * UPC) {
          const int r = tid / UPC, uu = tid % UPC, rb = r0 + r;
          const float* zc = a.zc0 + ((size_t)rb * N + t) * (4 * H) + c * UPC + uu;
          float* cs = cst + (size_t)rg * 2 * DEC_XR * UPC;
          float cc = cs[r * UPC + uu];
          const float h = lstm_cell(zs, r, uu, __ldg(zc), __ldg(zc + H), __ldg(zc + 2 * H), __ldg(zc + 3 * H), cc);
          cs[r * UPC + uu] = cc;
          a.h0[((size_t)(t & 1) * B + rb) * H + c * UPC + uu] = h;
          a.hout[((size_t)rb * N + t) * 2 * H + c * UPC + uu] = h;
        }
        if (NG > 1) __syncthreads();    // zs and xs are reused by the next group
      }
      DEC_MARK(4)
      grid.sync();
      DEC_MARK(5)
      // ---- phase D: LSTM1 = zc1[t] + p2 . W1[p2 rows] + h0_t . W1[h0 rows] + (h1_{t-1} part) ----
      for (int rg = 0; rg < NG; ++rg) {
        const int r0 = rg * DEC_XR, nb = min(DEC_XR, B - r0), ngroups = (nb + RG - 1) / RG;
        if (NG > 1) dec_fetch(xs, DEC_KPAD, 0, a.p2 + (size_t)r0 * vc::PRENET, vc::PRENET, vc::PRENET, nb);
        dec_fetch(xs, DEC_KPAD, vc::PRENET, a.h0 + ((size_t)(t & 1) * B + r0) * H, H, H, nb);
        __syncthreads();
        dec_matmul2<SLP, SLH>(xs, w1p, xs + vc::PRENET, w1h0, ngroups, part, zs, zp1 + rg * DEC_XR * NCOL);
        if (tid < nb * UPC) {
          const int r = tid / UPC, uu = tid % UPC, rb = r0 + r;
          const float* zc = a.zc1 + ((size_t)rb * N + t) * (4 * H) + c * UPC + uu;
          float* cs = cst + (size_t)rg * 2 * DEC_XR * UPC;
          float cc = cs[(DEC_XR + r) * UPC + uu];
          const float h = lstm_cell(zs, r, uu, __ldg(zc), __ldg(zc + H), __ldg(zc + 2 * H), __ldg(zc + 3 * H), cc);
          cs[(DEC_XR + r) * UPC + uu] = cc;
          a.h1[((size_t)(t & 1) * B + rb) * H + c * UPC + uu] = h;
          a.hout[((size_t)rb * N + t) * 2 * H + H + c * UPC + uu] = h;
        }
        __syncthreads();
      }
      DEC_MARK(6)
      grid.sync();
      DEC_MARK(7)
    }
  } else if (c < DEC_LSTM + DEC_PRE) {
    // =============================== prenet role ===============================
    const int q = c - DEC_LSTM;                        // columns 16q .. 16q+15 of p1 and p2
    constexpr int XP = H + 4;                          // 516: row pitch of the 512-wide staging buffer
    constexpr int WCB = (2 * H + 2 * H / 8) * 8;       // (Wo.W1) half-block: 1024 rows + one pad row per 8, x 8 columns
    constexpr int W2B = (vc::PRENET + NSLICE) * 8;
    float* xs = sm;                                    // [32][XP]: h1_{t-1} (EA), p1 (B), h0_t (D window)
    float* wcs = xs + 32 * XP;                         // [2][WCB]
    float* w2s = wcs + 2 * WCB;                        // [2][W2B]
    float* part = w2s + 2 * W2B;                       // [8][32][8]
    float* outv = part + 8 * 256;                      // [256]
    float* pp1 = outv + 256;                           // [DEC_NG][2][256] h0 half of p1's pre-activation, computed one phase early
    for (int hf = 0; hf < 2; ++hf) {
      pre_load_w8(wcs + (size_t)hf * WCB, a.wc + (size_t)q * 2 * H * 16, 2 * H, 8, 16, hf * 8);
      pre_load_w8(w2s + (size_t)hf * W2B, a.wp2 + (size_t)q * vc::PRENET * 16, vc::PRENET, 4, 16, hf * 8);
    }
    for (int e = tid; e < 32 * XP; e += SCAN_THREADS) xs[e] = 0.f;
    for (int e = tid; e < DEC_NG * 512; e += SCAN_THREADS) pp1[e] = 0.f;
    __syncthreads();
    const int orow = tid >> 3, ocol = tid & 7;         // output handled by this thread after a pre_gemm8 pass
    constexpr int H1OFF = (H + H / 8) * 8;             // first padded row of the h1 half inside a WCB block
    const int NG = (B + DEC_XR - 1) / DEC_XR;
    for (int t = 0; t < N; ++t) {
      // ---- phase EA: p1(t) = drop(relu(pp1 + h1_{t-1} . Wc[512:1024] + bc)) ----
      if (t > 0) {
        for (int rg = 0; rg < NG; ++rg) {
          const int r0 = rg * DEC_XR, nb = min(DEC_XR, B - r0);
          if (NG > 1) __syncthreads();
          dec_fetch(xs, XP, 0, a.h1 + ((size_t)((t - 1) & 1) * B + r0) * H, H, H, nb);
          __syncthreads();
#pragma unroll 1
          for (int hf = 0; hf < 2; ++hf) {
            pre_gemm8<8>(xs, XP, wcs + (size_t)hf * WCB + H1OFF, part, outv);
            if (orow < nb) {
              const int u = q * 16 + hf * 8 + ocol;
              const float v = fmaxf(outv[tid] + pp1[rg * 512 + hf * 256 + tid] + __ldg(a.bc + u), 0.f);
              a.p1[(size_t)(r0 + orow) * vc::PRENET + u] = v * keep_scale(a.mode, a.keep, a.seed, a.row_base + r0 + orow, t, N, 0, u);
            }
          }
        }
      } else {
        for (int e = tid; e < B * 16; e += SCAN_THREADS) a.p1[(size_t)(e >> 4) * vc::PRENET + q * 16 + (e & 15)] = 0.f;  // prenet(0) = 0
      }
      DEC_MARK(0)
      prenet_barrier(a.pre_bar, (unsigned)DEC_PRE * (unsigned)(t + 1), a.err);   // every column block of p1(t) is in L2
      DEC_MARK(1)
      // ---- phase B: p2(t) = drop(relu(p1 . W2)) ----
      for (int rg = 0; rg < NG; ++rg) {
        const int r0 = rg * DEC_XR, nb = min(DEC_XR, B - r0);
        if (NG > 1) __syncthreads();
        dec_fetch(xs, XP, 0, a.p1 + (size_t)r0 * vc::PRENET, vc::PRENET, vc::PRENET, nb);
        __syncthreads();
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          pre_gemm8<4>(xs, XP, w2s + (size_t)hf * W2B, part, outv);
          if (orow < nb) {
            const int u = q * 16 + hf * 8 + ocol;
            const float v = fmaxf(outv[tid], 0.f);
            a.p2[(size_t)(r0 + orow) * vc::PRENET + u] = v * keep_scale(a.mode, a.keep, a.seed, a.row_base + r0 + orow, t, N, 1, u);
          }
        }
      }
      DEC_MARK(2)
      grid.sync();
      DEC_MARK(3)
      grid.sync();
      DEC_MARK(5)
      // ---- D window: h0_t is final -> its half of p1(t+1)'s pre-activation ----
      for (int rg = 0; rg < NG; ++rg) {
        const int r0 = rg * DEC_XR, nb = min(DEC_XR, B - r0);
        __syncthreads();
        dec_fetch(xs, XP, 0, a.h0 + ((size_t)(t & 1) * B + r0) * H, H, H, nb);
        __syncthreads();
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          pre_gemm8<8>(xs, XP, wcs + (size_t)hf * WCB, part, outv);
          pp1[rg * 512 + hf * 256 + tid] = outv[tid];
        }
      }
      __syncthreads();
      DEC_MARK(6)
      grid.sync();
      DEC_MARK(7)
    }
  }
#undef DEC_MARK
  if (a.dbg && tid == 0)
    for (int i = 0; i < 8; ++i) a.dbg[(size_t)blockIdx.x * 16 + i] = tm[i];
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

// ---- teacher-forced decoder scan with zoneout (AcousticModel.__call__, model.py:146-169) ------------------
// The prenet and every input-side product are hoisted out of the recurrence (teacher forcing: the decoder input of
// frame t is the ground-truth frame t-1), so the scan only carries  h0s.W0[h0 rows],  h0n.W1[h0 rows]  and
// h1s.W1[h1 rows].  *n = the cores' new state (what the decoder outputs), *s = the state after zoneout
// (state = mask ? previous : new, model.py:157-159).  One grid barrier per frame: iteration s runs LSTM0 of frame s
// and LSTM1 of frame s-1 -- both only need values published in iteration s-1.
struct TfScanArgs {
  const float* zc0;      // [B][N][2048]  [cond, p2] . W0[0:768] + b0
  const float* zc1;      // [B][N][2048]  [cond, p2] . W1[0:768] + b1
  const float* w0r;      // [128][768][16]   rows 512..1279 of lstm0 ([p2, h0]); only the h0 rows are used here
  const float* w1r;      // [128][1280][16]  rows 512..1791 of lstm1 ([p2, h0, h1]); h0 and h1 rows used
  const uint8_t* zone;   // [B][N][4][512] (h0, c0, h1, c1; 1 = keep the previous state) or null
  uint64_t seed;
  int mode;              // VTTS_DROPOUT_OFF / MASK / SEED
  float* h0s;            // [2][B][512] zoned hidden state of layer 0, double buffered by frame parity
  float* h1s;            // [2][B][512]
  float* hout;           // [B][N][1024]  decoder outputs [h0n | h1n]
  int B, N;
  int row_base;
};

constexpr int TF_PITCH = 3 * vc::DEC_H + 4;    // [h0s | h0n | h1s] + pad

// true = keep the previous state.  SEED mode: Bernoulli(0.1) from the same threefry stream as the prenet masks,
// counter word 1 offset past the prenet's 2*256 entries.
__device__ __forceinline__ bool zone_keep(int mode, const uint8_t* zone, uint64_t seed, int b, int t, int N, int which, int unit) {
  if (mode == VTTS_DROPOUT_OFF) return false;
  if (mode == VTTS_DROPOUT_MASK) return zone[(((size_t)b * N + t) * 4 + which) * vc::DEC_H + unit] != 0;
  uint32_t o0, o1;
  (void)N;
  threefry2x32((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)t, ((uint32_t)b << 12) | (uint32_t)(2 * vc::PRENET + which * vc::DEC_H + unit), o0, o1);
  return o0 < 429496730u;   // 0.1 * 2^32
}

__global__ void __launch_bounds__(SCAN_THREADS, 1) decoder_tf_scan_kernel(const TfScanArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float sm[];
  constexpr int H = vc::DEC_H, K0 = vc::PRENET + H, K1 = vc::PRENET + 2 * H, SLH = H / NSLICE;   // 8
  const int c = blockIdx.x, tid = threadIdx.x;
  const int B = a.B, N = a.N;
  float* xs = sm;                                   // [DEC_XR][TF_PITCH]
  float* part = xs + DEC_XR * TF_PITCH;             // [8][DEC_XR][16]
  float* zs = part + 8 * DEC_XR * NCOL;             // [DEC_XR][16]
  float* cst = zs + DEC_XR * NCOL;                  // [2][DEC_XR][UPC] zoned cell states of the CTA's units
  const int ks = tid >> 2, cgp = tid & 3;
  float w0h[SLH][4], w1h0[SLH][4], w1h1[SLH][4];
  {
    auto ld = [&](float (&dst)[4], const float* base, int row) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)row * NCOL + cgp * 4));
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    };
    const float* g0 = a.w0r + (size_t)c * K0 * NCOL;
    const float* g1 = a.w1r + (size_t)c * K1 * NCOL;
#pragma unroll
    for (int i = 0; i < SLH; ++i) {
      ld(w0h[i], g0, vc::PRENET + ks * SLH + i);
      ld(w1h0[i], g1, vc::PRENET + ks * SLH + i);
      ld(w1h1[i], g1, vc::PRENET + H + ks * SLH + i);
    }
  }
  for (int e = tid; e < 2 * DEC_XR * UPC; e += SCAN_THREADS) cst[e] = 0.f;
  for (int e = tid; e < DEC_XR * TF_PITCH; e += SCAN_THREADS) xs[e] = 0.f;
  __syncthreads();
  const int ngroups = (B + RG - 1) / RG;
  for (int s = 0; s <= N; ++s) {
    // values published in iteration s-1: h0s_{s-1}, h0n_{s-1} (LSTM0 of frame s-1) and h1s_{s-2} (LSTM1 of frame s-2)
    // three copy groups in flight: LSTM0 of frame s only needs the first one, the operands of LSTM1 (frame s-1) keep
    // arriving while its product runs
    if (s >= 1) {
      dec_fetch_async(xs, TF_PITCH, 0, a.h0s + (size_t)((s - 1) & 1) * B * H, H, H, B);
      dec_fetch_async(xs, TF_PITCH, H, a.hout + (size_t)(s - 1) * 2 * H, H, N * 2 * H, B);
    } else {
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (s >= 2) dec_fetch_async(xs, TF_PITCH, 2 * H, a.h1s + (size_t)(s & 1) * B * H, H, H, B);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    cp_async_wait<2>();
    __syncthreads();
    if (s < N) {
      // ---- LSTM0 of frame s: z = zc0[s] + h0s_{s-1} . W0[h0 rows] ----
      dec_matmul<SLH, TF_PITCH>(xs, w0h, ngroups, part, zs, nullptr);
      if (tid < B * UPC) {
        const int r = tid / UPC, uu = tid % UPC, u = c * UPC + uu;
        const float* zc = a.zc0 + ((size_t)r * N + s) * (4 * H) + u;
        const float c_prev = cst[r * UPC + uu];
        float cc = c_prev;
        const float h = lstm_cell(zs, r, uu, __ldg(zc), __ldg(zc + H), __ldg(zc + 2 * H), __ldg(zc + 3 * H), cc);
        const bool kh = zone_keep(a.mode, a.zone, a.seed, a.row_base + r, s, N, 0, u);
        const bool kc = zone_keep(a.mode, a.zone, a.seed, a.row_base + r, s, N, 1, u);
        a.hout[((size_t)r * N + s) * 2 * H + u] = h;
        a.h0s[((size_t)(s & 1) * B + r) * H + u] = kh ? xs[(size_t)r * TF_PITCH + u] : h;
        cst[r * UPC + uu] = kc ? c_prev : cc;
      }
    }
    cp_async_wait<0>();
    if (s >= 1) {
      // ---- LSTM1 of frame s-1: z = zc1[s-1] + h0n_{s-1} . W1[h0 rows] + h1s_{s-2} . W1[h1 rows] ----
      const int t = s - 1;
      __syncthreads();    // zs is reused; every thread's copies of the LSTM1 operands have landed
      dec_matmul2<SLH, SLH, TF_PITCH>(xs + H, w1h0, xs + 2 * H, w1h1, ngroups, part, zs, nullptr);
      if (tid < B * UPC) {
        const int r = tid / UPC, uu = tid % UPC, u = c * UPC + uu;
        const float* zc = a.zc1 + ((size_t)r * N + t) * (4 * H) + u;
        const float c_prev = cst[(DEC_XR + r) * UPC + uu];
        float cc = c_prev;
        const float h = lstm_cell(zs, r, uu, __ldg(zc), __ldg(zc + H), __ldg(zc + 2 * H), __ldg(zc + 3 * H), cc);
        const bool kh = zone_keep(a.mode, a.zone, a.seed, a.row_base + r, t, N, 2, u);
        const bool kc = zone_keep(a.mode, a.zone, a.seed, a.row_base + r, t, N, 3, u);
        a.hout[((size_t)r * N + t) * 2 * H + H + u] = h;
        a.h1s[((size_t)(t & 1) * B + r) * H + u] = kh ? xs[(size_t)r * TF_PITCH + 2 * H + u] : h;
        cst[(DEC_XR + r) * UPC + uu] = kc ? c_prev : cc;
      }
    }
    __syncthreads();
    grid.sync();
  }
}

// out[row][c] (row stride out_ld) = relu(x[row][c]) * keep_scale  -- the two prenet dropouts applied to whole sequences
__global__ void prenet_act_kernel(const float* __restrict__ x, const uint8_t* __restrict__ keep, uint64_t seed, int mode, int layer,
                                  int B, int N, float* __restrict__ out, int out_ld) {
  const size_t total = (size_t)B * N * vc::PRENET;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int u = (int)(i % vc::PRENET);
    const size_t row = i / vc::PRENET;
    const int b = (int)(row / N), t = (int)(row % N);
    out[row * out_ld + u] = fmaxf(x[i], 0.f) * keep_scale(mode, keep, seed, b, t, N, layer, u);
  }
}

constexpr size_t tf_scan_smem() { return ((size_t)DEC_XR * TF_PITCH + 8 * DEC_XR * NCOL + DEC_XR * NCOL + 2 * DEC_XR * UPC) * 4; }

constexpr size_t enc_scan_smem() {
  return ((size_t)32 * (vc::ENC_D + 4) + 8 * DEC_XR * NCOL + DEC_XR * NCOL + MAX_ROWS * UPC) * 4;
}
constexpr size_t dec_scan_smem() {
  constexpr size_t lstm = (size_t)DEC_XR * DEC_KPAD + 8 * DEC_XR * NCOL + DEC_XR * NCOL + DEC_NG * (2 * DEC_XR * NCOL + 2 * DEC_XR * UPC);
  constexpr size_t pre = (size_t)32 * (vc::DEC_H + 4) + 2 * (2 * vc::DEC_H + 2 * vc::DEC_H / 8) * 8 + 2 * (vc::PRENET + NSLICE) * 8 + 8 * 256 + 256 + DEC_NG * 512;
  return (lstm > pre ? lstm : pre) * 4;
}

// derived-weight slots (ctx->ac_d)
enum {
  D_ENC_BNINV0 = 0, D_ENC_BNINV1, D_ENC_BNINV2,
  D_POST_BNINV0, D_POST_BNINV1, D_POST_BNINV2, D_POST_BNINV3,
  D_ENC_WHR,     // [2][64][256][16]
  D_DEC_W0R,     // [128][768][16]
  D_DEC_W1R,     // [128][1280][16]
  D_DEC_WC,      // [16][1024][16]  (Wo . W1) columns
  D_DEC_WCFULL,  // [1024][256] scratch
  D_DEC_BC,      // [256]
  D_DEC_WP2,     // [128][256][2]
  D_ZERO,        // [2048] zeros: bias of the bias-free prenet linears (model.py:88-89) on the generic conv path
  D_COUNT
};

// slots of ctx->ac_wpk_t (tensor-core packed weights)
enum { WP_ENC = 0, WP_ENCH = 3, WP_DECH = 11, WP_POST0 = 27, WP_POST1 = 29, WP_POST2 = 31, WP_POST3 = 33, WP_POST4 = 35, WP_PROJ = 36,
       // teacher-forced pass: [cond | p2] rows 0..767 of both decoder LSTMs (8 N=256 tiles each), the two prenet linears
       WP_TF_L0 = 37, WP_TF_L1 = 45, WP_PRE1 = 53, WP_PRE2 = 54, WP_COUNT = 55 };

}  // namespace

// TokenEncoder.__call__ (model.py:26-47, is_training=False): embed -> 3 x [conv k3, BN(eval), relu] -> BiLSTM.
// The acoustic model and the duration model instantiate it with the same dimensions (config.py:11-17), so one
// implementation serves both; only the weight pointers differ.
struct EncWeights {
  const float* embed;
  const float* conv_w[3];
  const float* conv_b[3];
  const float* bn_off[3];
  const float* bn_mean[3];
  const float* bn_inv[3];
  const float *lf_w, *lf_b, *lb_w, *lb_b;   // hk.LSTM linear of the forward / backward core, w[512][1024]
  const float* whr;                         // [2][64][256][16] recurrent rows, per direction / CTA
  void* const* wpk_conv;                    // 3 packed conv weights (tensor-core path)
  void* const* wpk_hoist;                   // 8 packed tiles of the two hoisted input projections
};

static int run_token_encoder(vtts_ctx* ctx, const EncWeights& w, const int32_t* tokens, const int32_t* lengths, int B, int L,
                             float* e0, float* e1, float* zx, float* enc, cudaStream_t st) {
  const size_t BL = (size_t)B * L;
  embed_kernel<<<(unsigned)((BL + 3) / 4), 256, 0, st>>>(tokens, w.embed, e0, (int)BL);
  ctx->launches++;
  VTTS_CUDA(cudaGetLastError());
  ConvLaunch Lc;
  float* cur = e0;
  float* nxt = e1;
  for (int i = 0; i < 3; ++i) {
    memset(&Lc, 0, sizeof(Lc));
    Lc.nprob = 1; Lc.Cin = 256; Lc.Cout = 256; Lc.B = B; Lc.T_rows = L; Lc.rows_out = L;
    Lc.len = lengths; Lc.len_mul = 1; Lc.pre_mode = 0; Lc.pre_slope = 1.f; Lc.post_act = 2;
    Lc.p[0] = ConvProb{cur, nullptr, nullptr, w.conv_w[i], w.conv_b[i], nullptr, w.bn_mean[i], w.bn_inv[i], w.bn_off[i], nxt, 3, 1, -1, 1, 0};
    int rc = vtts_conv_dispatch(ctx, Lc, w.wpk_conv + i, st);
    if (rc) return rc;
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
  // rows past len[b] of `cur` were never written: the scans mask them, but the hoisted GEMM reads them
  // -> harmless garbage confined to rows that are never consumed (k=1 GEMM has no row mixing).
  // ---- hoisted input projections of the two LSTMs: zx[dir] = x . W[0:256] + b ----
  memset(&Lc, 0, sizeof(Lc));
  Lc.nprob = 2; Lc.Cin = 256; Lc.Cout = 1024; Lc.B = 1; Lc.T_rows = (int)BL; Lc.rows_out = (int)BL;
  Lc.len = nullptr; Lc.len_mul = 1; Lc.pre_mode = 0; Lc.pre_slope = 1.f; Lc.post_act = 0;
  Lc.p[0] = ConvProb{cur, nullptr, nullptr, w.lf_w, w.lf_b, nullptr, nullptr, nullptr, nullptr, zx, 1, 1, 0, 1, 0};
  Lc.p[1] = ConvProb{cur, nullptr, nullptr, w.lb_w, w.lb_b, nullptr, nullptr, nullptr, nullptr, zx + BL * 1024, 1, 1, 0, 1, 0};
  int rc = vtts_conv_dispatch(ctx, Lc, w.wpk_hoist, st);
  if (rc) return rc;
  // ---- BiLSTM scan (forward core + ResetCore'd backward core) ----
  EncScanArgs ea;
  ea.zx = zx; ea.whr = w.whr; ea.lengths = lengths; ea.out = enc; ea.B = B; ea.L = L;
  void* args[] = {&ea};
  VTTS_CUDA(cudaLaunchCooperativeKernel((void*)enc_scan_kernel, dim3(SCAN_CTAS), dim3(SCAN_THREADS), args, enc_scan_smem(), st));
  ctx->launches++;
  return VTTS_OK;
}

int vtts_acoustic_prepare(vtts_ctx* ctx) {
  const size_t sizes[D_COUNT] = {256, 256, 256, 512, 512, 512, 512,
                                 (size_t)2 * 64 * 256 * 16, (size_t)128 * 768 * 16, (size_t)128 * 1280 * 16,
                                 (size_t)16 * 1024 * 16, (size_t)1024 * 256, 256, (size_t)16 * 256 * 16, 2048};
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
  repack_cols_kernel<<<256, 256>>>(ctx->ac_d[D_DEC_WCFULL], 256, 0, 1024, ctx->ac_d[D_DEC_WC], 16, 16, 16, 0);
  repack_cols_kernel<<<64, 256>>>(T[aci::PRE2_W], 256, 0, 256, ctx->ac_d[D_DEC_WP2], 16, 16, 16, 0);
  VTTS_CUDA(cudaMemset(ctx->ac_d[D_ZERO], 0, 2048 * sizeof(float)));
  VTTS_CUDA(cudaGetLastError());
  // ---- tensor-core packed weights of the convs and hoisted GEMMs ----
  {
    size_t bytes = 3 * vtts_tc_conv_packed_bytes(3, 256, 256) + 2 * vtts_tc_conv_packed_bytes(1, 256, 1024) +
                   2 * vtts_tc_conv_packed_bytes(1, 512, 2048) + vtts_tc_conv_packed_bytes(5, 80, 512) +
                   3 * vtts_tc_conv_packed_bytes(5, 512, 512) + vtts_tc_conv_packed_bytes(5, 512, 80) +
                   vtts_tc_conv_packed_bytes(1, 1024, 80) + 2 * vtts_tc_conv_packed_bytes(1, 768, 2048) +
                   vtts_tc_conv_packed_bytes(1, 80, 256) + vtts_tc_conv_packed_bytes(1, 256, 256);
    if (ctx->ac_wpk) cudaFree(ctx->ac_wpk);
    VTTS_CUDA(cudaMalloc(&ctx->ac_wpk, bytes));
    char* cur = (char*)ctx->ac_wpk;
    ctx->ac_wpk_t.clear();
    int rc = 0;
    for (int i = 0; i < 3 && !rc; ++i) rc = vtts_tc_pack_conv(ctx, T[aci::ENC_CONV(i, 0)], 3, 256, 256, cur, ctx->ac_wpk_t);
    if (!rc) rc = vtts_tc_pack_conv(ctx, T[aci::ENC_LSTM_F_W], 1, 256, 1024, cur, ctx->ac_wpk_t);
    if (!rc) rc = vtts_tc_pack_conv(ctx, T[aci::ENC_LSTM_B_W], 1, 256, 1024, cur, ctx->ac_wpk_t);
    if (!rc) rc = vtts_tc_pack_conv(ctx, T[aci::DEC_L0_W], 1, 512, 2048, cur, ctx->ac_wpk_t);
    if (!rc) rc = vtts_tc_pack_conv(ctx, T[aci::DEC_L1_W], 1, 512, 2048, cur, ctx->ac_wpk_t);
    if (!rc) rc = vtts_tc_pack_conv(ctx, T[aci::POST_CONV(0, 0)], 5, 80, 512, cur, ctx->ac_wpk_t);
    for (int i = 1; i < 4 && !rc; ++i) rc = vtts_tc_pack_conv(ctx, T[aci::POST_CONV(i, 0)], 5, 512, 512, cur, ctx->ac_wpk_t);
    if (!rc) rc = vtts_tc_pack_conv(ctx, T[aci::POST_CONV(4, 0)], 5, 512, 80, cur, ctx->ac_wpk_t);
    if (!rc) rc = vtts_tc_pack_conv(ctx, T[aci::PROJ_W], 1, 1024, 80, cur, ctx->ac_wpk_t);
    if (!rc) rc = vtts_tc_pack_conv(ctx, T[aci::DEC_L0_W], 1, 768, 2048, cur, ctx->ac_wpk_t);   // rows 0..767 = [cond | p2]
    if (!rc) rc = vtts_tc_pack_conv(ctx, T[aci::DEC_L1_W], 1, 768, 2048, cur, ctx->ac_wpk_t);
    if (!rc) rc = vtts_tc_pack_conv(ctx, T[aci::PRE1_W], 1, 80, 256, cur, ctx->ac_wpk_t);
    if (!rc) rc = vtts_tc_pack_conv(ctx, T[aci::PRE2_W], 1, 256, 256, cur, ctx->ac_wpk_t);
    if (rc) return rc;
    if ((int)ctx->ac_wpk_t.size() != WP_COUNT || (size_t)(cur - (char*)ctx->ac_wpk) > bytes)
      return ctx->fail(VTTS_ERR_BAD_ARG, "acoustic: packed weight table has %d entries", (int)ctx->ac_wpk_t.size());
  }
  VTTS_CUDA(cudaDeviceSynchronize());
  VTTS_CUDA(cudaFuncSetAttribute(enc_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc_scan_smem()));
  VTTS_CUDA(cudaFuncSetAttribute(decoder_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_scan_smem()));
  VTTS_CUDA(cudaFuncSetAttribute(decoder_tf_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tf_scan_smem()));
  return VTTS_OK;
}

// AcousticModel.postnet (model.py:113-121, is_training=False) + the residual add (:143-144 / :169):
// mel = melpre + conv5(tanh(bn(conv5(...))));  q0/q1 are [B*N][512] scratch
static int run_postnet(vtts_ctx* ctx, const float* melpre, const int32_t* n_frames, int B, int N, float* q0, float* q1, float* mel,
                       cudaStream_t st) {
  auto& T = ctx->ac_t;
  auto& D = ctx->ac_d;
  ConvLaunch Lc;
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
    int rc = vtts_conv_dispatch(ctx, Lc, &ctx->ac_wpk_t[i == 0 ? WP_POST0 : (i == 1 ? WP_POST1 : (i == 2 ? WP_POST2 : (i == 3 ? WP_POST3 : WP_POST4)))], st);
    if (rc) return rc;
    pin = pout;
    pout = (pout == q0) ? q1 : q0;
    cin = cout;
  }
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
    if (ctx->sm_count < DEC_CTAS) return ctx->fail(VTTS_ERR_NO_DEVICE, "scan kernels need %d SMs, device has %d", DEC_CTAS, ctx->sm_count);
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
  float* hout = ar.take<float>(BN * 1024);
  float* h0 = ar.take<float>((size_t)2 * MAX_ROWS * 512);
  float* h1 = ar.take<float>((size_t)2 * MAX_ROWS * 512);
  unsigned int* pre_bar = ar.take<unsigned int>(64);
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
  ctx->sub_mark(0, st);

  // ---- TokenEncoder (shared with the duration model, run_token_encoder above) ----
  {
    EncWeights ew;
    ew.embed = T[aci::EMBED];
    for (int i = 0; i < 3; ++i) {
      ew.conv_w[i] = T[aci::ENC_CONV(i, 0)]; ew.conv_b[i] = T[aci::ENC_CONV(i, 1)];
      ew.bn_off[i] = T[aci::ENC_CONV(i, 3)]; ew.bn_mean[i] = T[aci::ENC_CONV(i, 4)]; ew.bn_inv[i] = D[D_ENC_BNINV0 + i];
    }
    ew.lf_w = T[aci::ENC_LSTM_F_W]; ew.lf_b = T[aci::ENC_LSTM_F_B]; ew.lb_w = T[aci::ENC_LSTM_B_W]; ew.lb_b = T[aci::ENC_LSTM_B_B];
    ew.whr = D[D_ENC_WHR]; ew.wpk_conv = &ctx->ac_wpk_t[WP_ENC]; ew.wpk_hoist = &ctx->ac_wpk_t[WP_ENCH];
    int rc = run_token_encoder(ctx, ew, tokens, lengths, B, L, e0, e1, zx, enc, st);
    if (rc) return rc;
  }
  ctx->sub_mark(1, st);
  // ---- Gaussian upsampling ----
  {
    dim3 grid((N + UP_F - 1) / UP_F, B);
    size_t smem = (size_t)(L + UP_F * L) * sizeof(float);
    if (smem > 200 * 1024) return ctx->fail(VTTS_ERR_BAD_ARG, "acoustic: L=%d too long for the upsample kernel", L);
    if (smem > 48 * 1024) VTTS_CUDA(cudaFuncSetAttribute(upsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    upsample_kernel<<<grid, 256, smem, st>>>(enc, dur, lengths, n_frames, L, N, cond, vc::ENC_OUT);
    ctx->launches++;
    VTTS_CUDA(cudaGetLastError());
  }
  ctx->sub_mark(2, st);
  // ---- hoisted cond projections of the decoder LSTMs ----
  ConvLaunch Lc;
  memset(&Lc, 0, sizeof(Lc));
  Lc.nprob = 2; Lc.Cin = 512; Lc.Cout = 2048; Lc.B = 1; Lc.T_rows = (int)BN; Lc.rows_out = (int)BN;
  Lc.pre_mode = 0; Lc.pre_slope = 1.f; Lc.post_act = 0; Lc.len_mul = 1;
  Lc.p[0] = ConvProb{cond, nullptr, nullptr, T[aci::DEC_L0_W], T[aci::DEC_L0_B], nullptr, nullptr, nullptr, nullptr, zc0, 1, 1, 0, 1, 0};
  Lc.p[1] = ConvProb{cond, nullptr, nullptr, T[aci::DEC_L1_W], T[aci::DEC_L1_B], nullptr, nullptr, nullptr, nullptr, zc1, 1, 1, 0, 1, 0};
  int rc = vtts_conv_dispatch(ctx, Lc, &ctx->ac_wpk_t[WP_DECH], st);
  if (rc) return rc;
  ctx->sub_mark(3, st);
  // ---- autoregressive scan: ONE launch for up to DEC_NG * 32 rows (row groups share the grid barriers of a frame) ----
  for (int b0 = 0; b0 < B; b0 += DEC_NG * DEC_XR) {
    const int nb = B - b0 < DEC_NG * DEC_XR ? B - b0 : DEC_NG * DEC_XR;
    DecScanArgs da;
    memset(&da, 0, sizeof(da));
    da.zc0 = zc0 + (size_t)b0 * N * 2048; da.zc1 = zc1 + (size_t)b0 * N * 2048;
    da.w0r = D[D_DEC_W0R]; da.w1r = D[D_DEC_W1R]; da.wc = D[D_DEC_WC]; da.bc = D[D_DEC_BC]; da.wp2 = D[D_DEC_WP2];
    da.keep = keep; da.seed = seed; da.mode = mode;
    da.p1 = p1; da.p2 = p2; da.h0 = h0; da.h1 = h1; da.hout = hout + (size_t)b0 * N * 1024;
    da.pre_bar = pre_bar; da.err = ctx->d_err;
    VTTS_CUDA(cudaMemsetAsync(pre_bar, 0, sizeof(unsigned int), st));
    da.B = nb; da.N = N; da.row_base = b0; da.dbg = ctx->tc_dbg_on ? ctx->d_tc_dbg : nullptr;
    void* args[] = {&da};
    VTTS_CUDA(cudaLaunchCooperativeKernel((void*)decoder_scan_kernel, dim3(DEC_CTAS), dim3(SCAN_THREADS), args, dec_scan_smem(), st));
    ctx->launches++;
  }
  ctx->sub_mark(4, st);
  // ---- output projection of every frame in one GEMM: mel_pre = [h0 | h1] . Wo + bo (model.py:135);
  //      rows past n_frames[b] stay 0 ----
  memset(&Lc, 0, sizeof(Lc));
  Lc.nprob = 1; Lc.Cin = 1024; Lc.Cout = 80; Lc.B = B; Lc.T_rows = N; Lc.rows_out = N;
  Lc.len = n_frames; Lc.len_mul = 1; Lc.pre_mode = 0; Lc.pre_slope = 1.f; Lc.post_act = 0;
  Lc.p[0] = ConvProb{hout, nullptr, nullptr, T[aci::PROJ_W], T[aci::PROJ_B], nullptr, nullptr, nullptr, nullptr, melpre, 1, 1, 0, 1, 0};
  rc = vtts_conv_dispatch(ctx, Lc, &ctx->ac_wpk_t[WP_PROJ], st);
  if (rc) return rc;
  ctx->sub_mark(5, st);
  rc = run_postnet(ctx, melpre, n_frames, B, N, q0, q1, mel, st);
  if (rc) return rc;
  ctx->sub_mark(6, st);
  return VTTS_OK;
}

// AcousticModel.__call__ (model.py:146-169) with is_training=False, as gta.py:24-25 (`val_net`) runs it:
// encoder -> upsample to the mel length -> prenet of the SHIFTED ground-truth mel (dropout live) -> zoneout decoder
// scan -> projection -> postnet.  mel1 = projection output (may be null), mel2 = mel1 + postnet(mel1).
int vtts_acoustic_teacher_run(vtts_ctx* ctx, const int32_t* tokens, const int32_t* lengths, const float* dur,
                              const int32_t* n_frames, const float* mels_in, const uint8_t* keep, const uint8_t* zone, int mode,
                              uint64_t seed, int B, int L, int N, float* mel1, float* mel2, cudaStream_t st, void* ws_base,
                              size_t ws_cap, size_t* ws_need) {
  const bool measure = ws_need != nullptr;
  if (!measure) {
    if (!ctx->ac_loaded) return ctx->fail(VTTS_ERR_NOT_LOADED, "acoustic weights not loaded");
    if (B < 1 || L < 1 || N < 1 || B > MAX_ROWS)
      return ctx->fail(VTTS_ERR_BAD_ARG, "acoustic (teacher forced): B=%d L=%d N=%d (1 <= B <= %d rows per call)", B, L, N, MAX_ROWS);
    if (mode < 0 || mode > 2) return ctx->fail(VTTS_ERR_BAD_ARG, "acoustic (teacher forced): dropout_mode %d", mode);
    if (mode == VTTS_DROPOUT_MASK && (!keep || !zone))
      return ctx->fail(VTTS_ERR_BAD_ARG, "acoustic (teacher forced): dropout_mode MASK needs keep_mask and zone_mask");
    if (ctx->sm_count < SCAN_CTAS) return ctx->fail(VTTS_ERR_NO_DEVICE, "scan kernels need %d SMs, device has %d", SCAN_CTAS, ctx->sm_count);
  }
  Arena ar(ws_base, ws_cap, measure);
  const size_t BL = (size_t)B * L, BN = (size_t)B * N;
  constexpr int XW = vc::ENC_OUT + vc::PRENET;   // 768: decoder input [cond | prenet(mel)]
  float* e0 = ar.take<float>(BL * 256);
  float* e1 = ar.take<float>(BL * 256);
  float* zx = ar.take<float>(2 * BL * 1024);
  float* enc = ar.take<float>(BL * 512);
  float* xin = ar.take<float>(BN * XW);
  float* pa = ar.take<float>(BN * 256);
  float* pb = ar.take<float>(BN * 256);
  float* zc0 = ar.take<float>(BN * 2048);
  float* zc1 = ar.take<float>(BN * 2048);
  float* hout = ar.take<float>(BN * 1024);
  float* melpre = ar.take<float>(BN * 80);
  float* q0 = ar.take<float>(BN * 512);
  float* q1 = ar.take<float>(BN * 512);
  float* h0s = ar.take<float>((size_t)2 * DEC_XR * 512);
  float* h1s = ar.take<float>((size_t)2 * DEC_XR * 512);
  if (measure) {
    *ws_need = ar.off + 256;
    return VTTS_OK;
  }
  auto& T = ctx->ac_t;
  auto& D = ctx->ac_d;
  ctx->tap_enc = enc; ctx->tap_enc_n = BL * 512;
  ctx->tap_cond = nullptr; ctx->tap_cond_n = 0;
  ctx->tap_melpre = melpre; ctx->tap_melpre_n = BN * 80;
  VTTS_CUDA(cudaMemsetAsync(mel2, 0, BN * 80 * sizeof(float), st));
  if (mel1) VTTS_CUDA(cudaMemsetAsync(mel1, 0, BN * 80 * sizeof(float), st));
  VTTS_CUDA(cudaMemsetAsync(xin, 0, BN * XW * sizeof(float), st));
  VTTS_CUDA(cudaMemsetAsync(melpre, 0, BN * 80 * sizeof(float), st));
  ctx->sub_mark(16, st);
  {
    EncWeights ew;
    ew.embed = T[aci::EMBED];
    for (int i = 0; i < 3; ++i) {
      ew.conv_w[i] = T[aci::ENC_CONV(i, 0)]; ew.conv_b[i] = T[aci::ENC_CONV(i, 1)];
      ew.bn_off[i] = T[aci::ENC_CONV(i, 3)]; ew.bn_mean[i] = T[aci::ENC_CONV(i, 4)]; ew.bn_inv[i] = D[D_ENC_BNINV0 + i];
    }
    ew.lf_w = T[aci::ENC_LSTM_F_W]; ew.lf_b = T[aci::ENC_LSTM_F_B]; ew.lb_w = T[aci::ENC_LSTM_B_W]; ew.lb_b = T[aci::ENC_LSTM_B_B];
    ew.whr = D[D_ENC_WHR]; ew.wpk_conv = &ctx->ac_wpk_t[WP_ENC]; ew.wpk_hoist = &ctx->ac_wpk_t[WP_ENCH];
    int rc = run_token_encoder(ctx, ew, tokens, lengths, B, L, e0, e1, zx, enc, st);
    if (rc) return rc;
  }
  {  // cond -> columns 0..511 of the decoder input
    dim3 grid((N + UP_F - 1) / UP_F, B);
    size_t smem = (size_t)(L + UP_F * L) * sizeof(float);
    if (smem > 200 * 1024) return ctx->fail(VTTS_ERR_BAD_ARG, "acoustic: L=%d too long for the upsample kernel", L);
    if (smem > 48 * 1024) VTTS_CUDA(cudaFuncSetAttribute(upsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    upsample_kernel<<<grid, 256, smem, st>>>(enc, dur, lengths, n_frames, L, N, xin, XW);
    ctx->launches++;
    VTTS_CUDA(cudaGetLastError());
  }
  ctx->sub_mark(17, st);
  // ---- prenet over the whole sequence (model.py:95-100,149): two bias-free linears, relu, dropout 0.5 each ----
  ConvLaunch Lc;
  auto gemm = [&](const float* x, const float* w, const float* bias, float* out, int cin, int cout, int wp) {
    memset(&Lc, 0, sizeof(Lc));
    Lc.nprob = 1; Lc.Cin = cin; Lc.Cout = cout; Lc.B = 1; Lc.T_rows = (int)BN; Lc.rows_out = (int)BN;
    Lc.pre_mode = 0; Lc.pre_slope = 1.f; Lc.post_act = 0; Lc.len_mul = 1;
    Lc.p[0] = ConvProb{x, nullptr, nullptr, w, bias, nullptr, nullptr, nullptr, nullptr, out, 1, 1, 0, 1, 0};
    return vtts_conv_dispatch(ctx, Lc, &ctx->ac_wpk_t[wp], st);     // tensor-core path in BF16X3 mode, FMA path in FP32 mode
  };
  const size_t act_blocks = (BN * 256 + 255) / 256;
  const unsigned act_grid = (unsigned)(act_blocks < 148 * 16 ? act_blocks : 148 * 16);
  int rc = gemm(mels_in, T[aci::PRE1_W], D[D_ZERO], pa, 80, 256, WP_PRE1);
  if (rc) return rc;
  prenet_act_kernel<<<act_grid, 256, 0, st>>>(pa, keep, seed, mode, 0, B, N, pb, 256);
  ctx->launches++;
  rc = gemm(pb, T[aci::PRE2_W], D[D_ZERO], pa, 256, 256, WP_PRE2);
  if (rc) return rc;
  prenet_act_kernel<<<act_grid, 256, 0, st>>>(pa, keep, seed, mode, 1, B, N, xin + vc::ENC_OUT, XW);
  ctx->launches++;
  VTTS_CUDA(cudaGetLastError());
  // ---- every input-side product of both LSTMs in one launch: zc = [cond | p2] . W[0:768] + b ----
  memset(&Lc, 0, sizeof(Lc));
  Lc.nprob = 2; Lc.Cin = XW; Lc.Cout = 2048; Lc.B = 1; Lc.T_rows = (int)BN; Lc.rows_out = (int)BN;
  Lc.pre_mode = 0; Lc.pre_slope = 1.f; Lc.post_act = 0; Lc.len_mul = 1;
  Lc.p[0] = ConvProb{xin, nullptr, nullptr, T[aci::DEC_L0_W], T[aci::DEC_L0_B], nullptr, nullptr, nullptr, nullptr, zc0, 1, 1, 0, 1, 0};
  Lc.p[1] = ConvProb{xin, nullptr, nullptr, T[aci::DEC_L1_W], T[aci::DEC_L1_B], nullptr, nullptr, nullptr, nullptr, zc1, 1, 1, 0, 1, 0};
  rc = vtts_conv_dispatch(ctx, Lc, &ctx->ac_wpk_t[WP_TF_L0], st);      // tiles of problem 0 then problem 1: WP_TF_L0 .. WP_TF_L1+7
  if (rc) return rc;
  ctx->sub_mark(18, st);
  // ---- zoneout scan, <= 32 rows per launch ----
  for (int b0 = 0; b0 < B; b0 += DEC_XR) {
    const int nb = B - b0 < DEC_XR ? B - b0 : DEC_XR;
    TfScanArgs ta;
    memset(&ta, 0, sizeof(ta));
    ta.zc0 = zc0 + (size_t)b0 * N * 2048; ta.zc1 = zc1 + (size_t)b0 * N * 2048;
    ta.w0r = D[D_DEC_W0R]; ta.w1r = D[D_DEC_W1R];
    ta.zone = zone; ta.seed = seed; ta.mode = mode;
    ta.h0s = h0s; ta.h1s = h1s; ta.hout = hout + (size_t)b0 * N * 1024;
    ta.B = nb; ta.N = N; ta.row_base = b0;
    void* args[] = {&ta};
    VTTS_CUDA(cudaLaunchCooperativeKernel((void*)decoder_tf_scan_kernel, dim3(SCAN_CTAS), dim3(SCAN_THREADS), args, tf_scan_smem(), st));
    ctx->launches++;
  }
  ctx->sub_mark(19, st);
  // ---- projection over all frames, then the postnet ----
  memset(&Lc, 0, sizeof(Lc));
  Lc.nprob = 1; Lc.Cin = 1024; Lc.Cout = 80; Lc.B = B; Lc.T_rows = N; Lc.rows_out = N;
  Lc.len = n_frames; Lc.len_mul = 1; Lc.pre_mode = 0; Lc.pre_slope = 1.f; Lc.post_act = 0;      // rows past n_frames[b] stay 0
  Lc.p[0] = ConvProb{hout, nullptr, nullptr, T[aci::PROJ_W], T[aci::PROJ_B], nullptr, nullptr, nullptr, nullptr, melpre, 1, 1, 0, 1, 0};
  rc = vtts_conv_dispatch(ctx, Lc, &ctx->ac_wpk_t[WP_PROJ], st);
  if (rc) return rc;
  if (mel1) VTTS_CUDA(cudaMemcpyAsync(mel1, melpre, BN * 80 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  rc = run_postnet(ctx, melpre, n_frames, B, N, q0, q1, mel2, st);
  ctx->sub_mark(20, st);
  return rc;
}

// =====================================================================================================
// DurationModel (vietTTS/nat/model.py:49-70): TokenEncoder -> Linear(512->256) -> gelu -> Linear(256->1) -> softplus.
// The encoder is run_token_encoder above with the duration checkpoint's weights; the first Linear is a k=1
// contraction through the shared conv dispatch; the head below finishes gelu / dot / softplus per token.
// =====================================================================================================
namespace {

// jax.nn.gelu default (approximate=True): 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))
__device__ __forceinline__ float gelu_tanh(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.f + tanhf(u));
}
// jax.nn.softplus = logaddexp(x, 0)
__device__ __forceinline__ float softplus(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

// one warp per token: y [BL][256] (pre-activation of the first Linear, bias included) -> dur[BL] seconds
__global__ void __launch_bounds__(256) duration_head_kernel(const float* __restrict__ y, const float* __restrict__ w2,
                                                            const float* __restrict__ b2, const int32_t* __restrict__ lengths,
                                                            int B, int L, float* __restrict__ dur) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tok = blockIdx.x * 8 + warp;
  if (tok >= B * L) return;
  const int b = tok / L, l = tok - b * L;
  if (lengths && l >= lengths[b]) {
    if (lane == 0) dur[tok] = 0.f;
    return;
  }
  const float* yr = y + (size_t)tok * 256 + lane * 8;
  const float4 a0 = __ldg(reinterpret_cast<const float4*>(yr)), a1 = __ldg(reinterpret_cast<const float4*>(yr + 4));
  const float4 w0 = __ldg(reinterpret_cast<const float4*>(w2 + lane * 8)), w1 = __ldg(reinterpret_cast<const float4*>(w2 + lane * 8 + 4));
  float s = gelu_tanh(a0.x) * w0.x;
  s = fmaf(gelu_tanh(a0.y), w0.y, s);
  s = fmaf(gelu_tanh(a0.z), w0.z, s);
  s = fmaf(gelu_tanh(a0.w), w0.w, s);
  s = fmaf(gelu_tanh(a1.x), w1.x, s);
  s = fmaf(gelu_tanh(a1.y), w1.y, s);
  s = fmaf(gelu_tanh(a1.z), w1.z, s);
  s = fmaf(gelu_tanh(a1.w), w1.w, s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) dur[tok] = softplus(s + __ldg(b2));
}

enum { DU_BNINV0 = 0, DU_BNINV1, DU_BNINV2, DU_WHR, DU_COUNT };
enum { DWP_ENC = 0, DWP_ENCH = 3, DWP_FC1 = 11, DWP_COUNT = 12 };

}  // namespace

int vtts_duration_prepare(vtts_ctx* ctx) {
  const size_t sizes[DU_COUNT] = {256, 256, 256, (size_t)2 * 64 * 256 * 16};
  size_t total = 0;
  std::vector<size_t> offs(DU_COUNT);
  for (int i = 0; i < DU_COUNT; ++i) {
    offs[i] = total;
    total += (sizes[i] + 63) & ~size_t(63);
  }
  if (ctx->du_derived) cudaFree(ctx->du_derived);
  VTTS_CUDA(cudaMalloc(&ctx->du_derived, total * sizeof(float)));
  ctx->du_d.resize(DU_COUNT);
  for (int i = 0; i < DU_COUNT; ++i) ctx->du_d[i] = ctx->du_derived + offs[i];
  auto& T = ctx->du_t;
  for (int i = 0; i < 3; ++i) bn_inv_kernel<<<1, 256>>>(T[aci::ENC_CONV(i, 2)], T[aci::ENC_CONV(i, 5)], ctx->du_d[DU_BNINV0 + i], 256);
  repack_cols_kernel<<<256, 256>>>(T[aci::ENC_LSTM_F_W], 1024, 256, 256, ctx->du_d[DU_WHR], 64, 16, UPC, 256);
  repack_cols_kernel<<<256, 256>>>(T[aci::ENC_LSTM_B_W], 1024, 256, 256, ctx->du_d[DU_WHR] + (size_t)64 * 256 * 16, 64, 16, UPC, 256);
  VTTS_CUDA(cudaGetLastError());
  {
    const size_t bytes = 3 * vtts_tc_conv_packed_bytes(3, 256, 256) + 2 * vtts_tc_conv_packed_bytes(1, 256, 1024) +
                         vtts_tc_conv_packed_bytes(1, 512, 256);
    if (ctx->du_wpk) cudaFree(ctx->du_wpk);
    VTTS_CUDA(cudaMalloc(&ctx->du_wpk, bytes));
    char* cur = (char*)ctx->du_wpk;
    ctx->du_wpk_t.clear();
    int rc = 0;
    for (int i = 0; i < 3 && !rc; ++i) rc = vtts_tc_pack_conv(ctx, T[aci::ENC_CONV(i, 0)], 3, 256, 256, cur, ctx->du_wpk_t);
    if (!rc) rc = vtts_tc_pack_conv(ctx, T[aci::ENC_LSTM_F_W], 1, 256, 1024, cur, ctx->du_wpk_t);
    if (!rc) rc = vtts_tc_pack_conv(ctx, T[aci::ENC_LSTM_B_W], 1, 256, 1024, cur, ctx->du_wpk_t);
    if (!rc) rc = vtts_tc_pack_conv(ctx, T[dui::FC1_W], 1, 512, 256, cur, ctx->du_wpk_t);
    if (rc) return rc;
    if ((int)ctx->du_wpk_t.size() != DWP_COUNT || (size_t)(cur - (char*)ctx->du_wpk) > bytes)
      return ctx->fail(VTTS_ERR_BAD_ARG, "duration: packed weight table has %d entries", (int)ctx->du_wpk_t.size());
  }
  VTTS_CUDA(cudaDeviceSynchronize());
  VTTS_CUDA(cudaFuncSetAttribute(enc_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc_scan_smem()));
  return VTTS_OK;
}

int vtts_duration_run(vtts_ctx* ctx, const int32_t* tokens, const int32_t* lengths, int B, int L, float* dur_sec,
                      cudaStream_t st, void* ws_base, size_t ws_cap, size_t* ws_need) {
  const bool measure = ws_need != nullptr;
  if (!measure) {
    if (!ctx->du_loaded) return ctx->fail(VTTS_ERR_NOT_LOADED, "duration weights not loaded");
    if (B < 1 || L < 1 || B > MAX_ROWS)
      return ctx->fail(VTTS_ERR_BAD_ARG, "duration: B=%d L=%d (1 <= B <= %d rows per call; the host layer chunks larger batches)", B, L, MAX_ROWS);
    if (ctx->sm_count < SCAN_CTAS) return ctx->fail(VTTS_ERR_NO_DEVICE, "scan kernels need %d SMs, device has %d", SCAN_CTAS, ctx->sm_count);
  }
  Arena ar(ws_base, ws_cap, measure);
  const size_t BL = (size_t)B * L;
  float* e0 = ar.take<float>(BL * 256);
  float* e1 = ar.take<float>(BL * 256);
  float* zx = ar.take<float>(2 * BL * 1024);
  float* enc = ar.take<float>(BL * 512);
  float* y = ar.take<float>(BL * 256);
  if (measure) {
    *ws_need = ar.off + 256;
    return VTTS_OK;
  }
  auto& T = ctx->du_t;
  auto& D = ctx->du_d;
  EncWeights ew;
  ew.embed = T[aci::EMBED];
  for (int i = 0; i < 3; ++i) {
    ew.conv_w[i] = T[aci::ENC_CONV(i, 0)]; ew.conv_b[i] = T[aci::ENC_CONV(i, 1)];
    ew.bn_off[i] = T[aci::ENC_CONV(i, 3)]; ew.bn_mean[i] = T[aci::ENC_CONV(i, 4)]; ew.bn_inv[i] = D[DU_BNINV0 + i];
  }
  ew.lf_w = T[aci::ENC_LSTM_F_W]; ew.lf_b = T[aci::ENC_LSTM_F_B]; ew.lb_w = T[aci::ENC_LSTM_B_W]; ew.lb_b = T[aci::ENC_LSTM_B_B];
  ew.whr = D[DU_WHR]; ew.wpk_conv = &ctx->du_wpk_t[DWP_ENC]; ew.wpk_hoist = &ctx->du_wpk_t[DWP_ENCH];
  // padded encoder rows are never written by the scan: clear them so the projection reads zeros, not stale workspace
  VTTS_CUDA(cudaMemsetAsync(enc, 0, BL * 512 * sizeof(float), st));
  int rc = run_token_encoder(ctx, ew, tokens, lengths, B, L, e0, e1, zx, enc, st);
  if (rc) return rc;
  ctx->tap_enc = enc; ctx->tap_enc_n = BL * 512;
  // ---- projection head ----
  ConvLaunch Lc;
  memset(&Lc, 0, sizeof(Lc));
  Lc.nprob = 1; Lc.Cin = 512; Lc.Cout = 256; Lc.B = 1; Lc.T_rows = (int)BL; Lc.rows_out = (int)BL;
  Lc.len = nullptr; Lc.len_mul = 1; Lc.pre_mode = 0; Lc.pre_slope = 1.f; Lc.post_act = 0;
  Lc.p[0] = ConvProb{enc, nullptr, nullptr, T[dui::FC1_W], T[dui::FC1_B], nullptr, nullptr, nullptr, nullptr, y, 1, 1, 0, 1, 0};
  rc = vtts_conv_dispatch(ctx, Lc, &ctx->du_wpk_t[DWP_FC1], st);
  if (rc) return rc;
  duration_head_kernel<<<(unsigned)((BL + 7) / 8), 256, 0, st>>>(y, T[dui::FC2_W], T[dui::FC2_B], lengths, B, L, dur_sec);
  ctx->launches++;
  VTTS_CUDA(cudaGetLastError());
  return VTTS_OK;
}
