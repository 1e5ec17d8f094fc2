// STFT + log-mel -- restates MelFilter.__call__ (vietTTS/nat/dsp.py:115-128):
//   reflect pad 384 | frame 1024 / hop 256 | periodic Hann | FFT-1024, bins 0..512 |
//   sqrt(re^2+im^2+1e-9) | 80x513 Slaney filterbank | log(max(.,1e-5))
// One CTA per PAIR of frames: the two real frames are packed into one complex FFT-1024
// (frame A real part, frame B imaginary part) and separated afterwards, so the 4x overlap of
// neighbouring frames is served from L1/L2 and every sample is read from HBM once.
// FFT = 5 radix-4 Stockham passes in shared memory, 256 threads = 256 butterflies per pass.
#include <math.h>

#include "vtts_internal.cuh"

namespace {

constexpr int NF = vc::NFFT;      // 1024
constexpr int NB = vc::NBINS;     // 513

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__global__ void __launch_bounds__(256) melspec_kernel(const float* __restrict__ wav, int S, int F,
                                                      const float* __restrict__ hann, const float2* __restrict__ tw,
                                                      const float* __restrict__ fb, const int* __restrict__ lo,
                                                      const int* __restrict__ hi, float* __restrict__ mel) {
  __shared__ float2 buf0[NF];
  __shared__ float2 buf1[NF];
  __shared__ float magA[NB + 3];
  __shared__ float magB[NB + 3];
  const int b = blockIdx.y, fa = blockIdx.x * 2, fbn = fa + 1, tid = threadIdx.x;
  const bool hasB = fbn < F;
  const float* y = wav + (size_t)b * S;
  // ---- load + reflect pad (p = 384) + window ----
  for (int i = tid; i < NF; i += 256) {
    const float h = hann[i];
    int ia = fa * vc::HOP - 384 + i;
    ia = ia < 0 ? -ia : (ia >= S ? 2 * (S - 1) - ia : ia);
    float va = __ldg(y + ia) * h, vb = 0.f;
    if (hasB) {
      int ib = fbn * vc::HOP - 384 + i;
      ib = ib < 0 ? -ib : (ib >= S ? 2 * (S - 1) - ib : ib);
      vb = __ldg(y + ib) * h;
    }
    buf0[i] = make_float2(va, vb);
  }
  __syncthreads();
  // ---- radix-4 Stockham, Ns = 1,4,16,64,256 ----
  float2* in = buf0;
  float2* out = buf1;
#pragma unroll
  for (int Ns = 1; Ns < NF; Ns *= 4) {
    const int j = tid;
    const int kk = j % Ns;
    const int tstep = (NF / 4) / Ns;  // twiddle index = kk * tstep * r  (angle -2 pi kk r / (4 Ns))
    float2 v0 = in[j], v1 = in[j + NF / 4], v2 = in[j + NF / 2], v3 = in[j + 3 * NF / 4];
    if (Ns > 1) {
      v1 = cmul(v1, tw[kk * tstep]);
      v2 = cmul(v2, tw[2 * kk * tstep]);
      v3 = cmul(v3, tw[3 * kk * tstep]);
    }
    const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y);
    const float2 a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
    const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
    const float2 d = make_float2(v1.x - v3.x, v1.y - v3.y);
    const float2 a3 = make_float2(d.y, -d.x);  // -i * (v1 - v3)
    const int j0 = (j / Ns) * Ns * 4 + kk;
    out[j0] = make_float2(a0.x + a2.x, a0.y + a2.y);
    out[j0 + Ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
    out[j0 + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
    out[j0 + 3 * Ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
    __syncthreads();
    float2* t = in; in = out; out = t;
  }
  // ---- separate the two real transforms, magnitude ----
  // Z = FFT(a + i b):  A[k] = (Z[k] + conj(Z[N-k]))/2,  B[k] = (Z[k] - conj(Z[N-k]))/(2i)
  for (int k = tid; k < NB; k += 256) {
    const float2 z = in[k];
    const float2 zc = in[(NF - k) & (NF - 1)];
    const float ar = 0.5f * (z.x + zc.x), ai = 0.5f * (z.y - zc.y);
    const float br = 0.5f * (z.y + zc.y), bi = -0.5f * (z.x - zc.x);
    magA[k] = sqrtf(ar * ar + ai * ai + 1e-9f);
    magB[k] = sqrtf(br * br + bi * bi + 1e-9f);
  }
  __syncthreads();
  // ---- filterbank (only the non-zero span of each triangle) + log ----
  if (tid < 2 * vc::MEL) {
    const int m = tid % vc::MEL, which = tid / vc::MEL;
    if (which == 0 || hasB) {
      const float* mg = which == 0 ? magA : magB;
      const float* frow = fb + (size_t)m * NB;
      float s = 0.f;
      for (int k = lo[m]; k < hi[m]; ++k) s = fmaf(__ldg(frow + k), mg[k], s);
      const int f = which == 0 ? fa : fbn;
      mel[((size_t)b * F + f) * vc::MEL + m] = logf(fmaxf(s, 1e-5f));
    }
  }
}

__global__ void mel_span_kernel(const float* __restrict__ fb, int* lo, int* hi) {
  const int m = threadIdx.x;
  if (m >= vc::MEL) return;
  int l = NB, h = 0;
  for (int k = 0; k < NB; ++k)
    if (fb[(size_t)m * NB + k] != 0.f) {
      if (k < l) l = k;
      h = k + 1;
    }
  if (h == 0) l = 0;
  lo[m] = l;
  hi[m] = h;
}

}  // namespace

int vtts_melspec_prepare(vtts_ctx* ctx) {
  // twiddles exp(-2 pi i k / 1024) and the periodic Hann window, computed in double on the host
  std::vector<float> tw(2 * NF), hn(NF);
  for (int k = 0; k < NF; ++k) {
    const double a = -2.0 * M_PI * (double)k / (double)NF;
    tw[2 * k] = (float)cos(a);
    tw[2 * k + 1] = (float)sin(a);
    hn[k] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)k / (double)NF));  // np.hanning(1025)[:-1], dsp.py:81
  }
  if (!ctx->fft_tw) VTTS_CUDA(cudaMalloc(&ctx->fft_tw, tw.size() * sizeof(float)));
  if (!ctx->hann) VTTS_CUDA(cudaMalloc(&ctx->hann, hn.size() * sizeof(float)));
  if (!ctx->mel_lo) VTTS_CUDA(cudaMalloc(&ctx->mel_lo, vc::MEL * sizeof(int)));
  if (!ctx->mel_hi) VTTS_CUDA(cudaMalloc(&ctx->mel_hi, vc::MEL * sizeof(int)));
  VTTS_CUDA(cudaMemcpy(ctx->fft_tw, tw.data(), tw.size() * sizeof(float), cudaMemcpyHostToDevice));
  VTTS_CUDA(cudaMemcpy(ctx->hann, hn.data(), hn.size() * sizeof(float), cudaMemcpyHostToDevice));
  mel_span_kernel<<<1, 128>>>(ctx->mel_fb, ctx->mel_lo, ctx->mel_hi);
  VTTS_CUDA(cudaGetLastError());
  VTTS_CUDA(cudaDeviceSynchronize());
  return VTTS_OK;
}

int vtts_melspec_run(vtts_ctx* ctx, const float* wav, int B, int S, float* mel, cudaStream_t st) {
  if (!ctx->mel_loaded) return ctx->fail(VTTS_ERR_NOT_LOADED, "mel filterbank not loaded");
  if (B < 1 || B > 65535 || S < 512 || S % vc::HOP != 0)
    return ctx->fail(VTTS_ERR_BAD_ARG, "melspec: B=%d S=%d (need S %% 256 == 0, S >= 512)", B, S);
  const int F = S / vc::HOP;
  dim3 grid((F + 1) / 2, B);
  melspec_kernel<<<grid, 256, 0, st>>>(wav, S, F, ctx->hann, reinterpret_cast<const float2*>(ctx->fft_tw), ctx->mel_fb,
                                       ctx->mel_lo, ctx->mel_hi, mel);
  ctx->launches++;
  VTTS_CUDA(cudaGetLastError());
  return VTTS_OK;
}
