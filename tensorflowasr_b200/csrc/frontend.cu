// Mel/dB frontend for sm_100a: wav -> |STFT|^2 -> dB (per-utterance max, -80 dB floor) -> mel projection.
//
// Replaces asr/models/layers/time_frequency.py:100-122,173-189 + backend_keras.py:5-23 of the reference, which
// evaluates the STFT as two dense 1024-tap strided convolutions (2.1 GFLOP / 10 s utterance).  Here each CTA runs
// one 1024-point complex radix-4 Stockham FFT in shared memory that carries TWO real frames (frame 2i in the
// real lane, frame 2i+1 in the imaginary lane) and untangles them with the Hermitian symmetry, i.e. ~26 MFLOP per
// utterance.  HBM-bound by design: reads the waveform once (frames overlap in L1/L2), writes the power
// spectrogram once; the dB + mel kernel re-reads it once after the per-utterance max is known.
#include "kernels.cuh"
#include "stft_warp.cuh"

#include <cstdlib>

namespace b200asr {

constexpr int kNfft = 1024;
constexpr int kBins = kNfft / 2 + 1;  // 513

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// grid (ceil(T/2), B), block 256.  power: [B, T, ps] (ps >= 513 row stride), pmax: [B] float bits (>= 0).
__global__ void __launch_bounds__(256) stft_power_kernel(const float* __restrict__ wav, const float* __restrict__ window,
                                                         const float2* __restrict__ twiddle, float* __restrict__ power,
                                                         unsigned int* __restrict__ pmax, int L, int T, int pad_left,
                                                         int hop, int ps) {
  __shared__ float2 buf0[kNfft];
  __shared__ float2 buf1[kNfft];
  __shared__ float2 tw[kNfft];
  __shared__ float red[8];

  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * 2;
  const bool has_second = (t0 + 1) < T;
  const float* w = wav + (size_t)b * L;

  float2 v[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int n = tid + r * 256;
    tw[n] = twiddle[n];
    const float wn = window[n];
    const int s0 = t0 * hop - pad_left + n;
    const int s1 = s0 + hop;
    float x0 = (s0 >= 0 && s0 < L) ? w[s0] : 0.0f;
    float x1 = (has_second && s1 >= 0 && s1 < L) ? w[s1] : 0.0f;
    v[r] = make_float2(x0 * wn, x1 * wn);
  }
  __syncthreads();

  float2* src = buf0;
  float2* dst = buf1;
#pragma unroll
  for (int pass = 0; pass < 5; ++pass) {
    const int Ns = 1 << (2 * pass);
    const int k = tid & (Ns - 1);
    if (pass > 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r) v[r] = src[tid + r * 256];
      const int step = 256 / Ns;  // twiddle exponent = r * k * step  (< 1024)
      v[1] = cmul(v[1], tw[k * step]);
      v[2] = cmul(v[2], tw[2 * k * step]);
      v[3] = cmul(v[3], tw[3 * k * step]);
    }
    // radix-4 butterfly (forward DFT)
    float2 a0 = make_float2(v[0].x + v[2].x, v[0].y + v[2].y);
    float2 a1 = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
    float2 a2 = make_float2(v[1].x + v[3].x, v[1].y + v[3].y);
    float2 d = make_float2(v[1].x - v[3].x, v[1].y - v[3].y);
    float2 a3 = make_float2(d.y, -d.x);  // * (-i)
    const int j0 = ((tid - k) << 2) + k;
    dst[j0] = make_float2(a0.x + a2.x, a0.y + a2.y);
    dst[j0 + Ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
    dst[j0 + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
    dst[j0 + 3 * Ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
    __syncthreads();
    float2* tmp = src;
    src = dst;
    dst = tmp;
  }
  // src holds Z[0..1023]; frame A = Re lane, frame B = Im lane.
  float* pa = power + ((size_t)b * T + t0) * ps;
  float* pb = pa + ps;
  float vmax = 0.0f;
  for (int k = tid; k < kBins; k += 256) {
    const float2 z = src[k];
    const float2 zc = src[(kNfft - k) & (kNfft - 1)];
    const float ar = z.x + zc.x, ai = z.y - zc.y;
    const float br = z.y + zc.y, bi = z.x - zc.x;
    const float p0 = 0.25f * (ar * ar + ai * ai);
    pa[k] = p0;
    vmax = fmaxf(vmax, p0);
    if (has_second) {
      const float p1 = 0.25f * (br * br + bi * bi);
      pb[k] = p1;
      vmax = fmaxf(vmax, p1);
    }
  }
  vmax = warp_max(vmax);
  if ((tid & 31) == 0) red[tid >> 5] = vmax;
  __syncthreads();
  if (tid < 8) {
    float m = red[tid];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
    if (tid == 0) atomicMax(pmax + b, __float_as_uint(m));  // non-negative floats order like unsigned ints
  }
}

// Warp-per-FFT variant (stft_warp.cuh): each warp transforms kSwPairs frame pairs with a register-resident 32 x 32 four-step
// FFT; the only shared-memory traffic is one transposed exchange and the natural-order spectrum, both warp-private and
// bank-conflict free, and there is no block barrier inside the transform.  grid (ceil(pairs / (kSwWarps*kSwPairs)), B).
constexpr int kSwWarps = 8, kSwPairs = 2;
__global__ void __launch_bounds__(kSwWarps * 32) stft_power_warp_kernel(const float* __restrict__ wav, const float* __restrict__ window,
                                                                        const float2* __restrict__ twiddle, float* __restrict__ power,
                                                                        unsigned int* __restrict__ pmax, int L, int T, int pad_left, int hop,
                                                                        int ps) {
  extern __shared__ __align__(16) unsigned char sw_smem[];
  float2* tw = reinterpret_cast<float2*>(sw_smem);            // [1024]
  float* win = reinterpret_cast<float*>(tw + kNfft);          // [1024]
  __shared__ float red[kSwWarps];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float2* sb = reinterpret_cast<float2*>(win + kNfft) + warp * kSwTile;
  pdl_trigger();
  for (int i = tid; i < kNfft; i += kSwWarps * 32) {          // constant tables: staged while the previous kernel may still run
    tw[i] = twiddle[i];
    win[i] = window[i];
  }
  __syncthreads();
  pdl_wait();
  const int b = blockIdx.y;
  const float* w = wav + (size_t)b * L;
  float vmax = 0.0f;
#pragma unroll 1
  for (int q = 0; q < kSwPairs; ++q) {
    const int pair = (blockIdx.x * kSwPairs + q) * kSwWarps + warp;   // a CTA's warps take consecutive frame pairs (they share samples in L1)
    const int t0 = 2 * pair;
    if (t0 < T) {                                                     // warp-uniform
      const bool has_second = (t0 + 1) < T;
      stft_pass_a(lane, w, L, t0 * hop - pad_left, hop, has_second, win, tw, sb);
      __syncwarp();
      float2 v[32];
      stft_pass_b_load(lane, sb, v);
      __syncwarp();
      stft_pass_b_store(lane, v, sb);
      __syncwarp();
      float* pa = power + ((size_t)b * T + t0) * ps;
      vmax = fmaxf(vmax, stft_untangle(lane, sb, pa, pa + ps, has_second));
      __syncwarp();
    }
  }
  vmax = warp_max(vmax);
  if (lane == 0) red[warp] = vmax;
  __syncthreads();
  if (tid < 32) {
    float m = (tid < kSwWarps) ? red[tid] : 0.0f;
    m = warp_max(m);
    if (tid == 0) atomicMax(pmax + b, __float_as_uint(m));    // non-negative floats order like unsigned ints
  }
}

// 10*log(max(p,1e-10))/ln(10) exactly as backend_keras.py:15 writes it in fp32.
__device__ __forceinline__ float to_db(float p) { return 10.0f * logf(fmaxf(p, 1e-10f)) / 2.302585093f; }

// grid (ceil(T/FR), B), block 256; mel: [B, T, n_mels].  mode 0: offline ('same' -> per-utterance max, -80 floor);
// mode 1: chunk ('valid' -> log10(max(p,1e-10)) only, backend_keras.py:25-37).
template <int FR>
__global__ void __launch_bounds__(256) db_mel_kernel(const float* __restrict__ power, const unsigned int* __restrict__ pmax,
                                                     const float* __restrict__ melw, const int* __restrict__ mel_lo,
                                                     const int* __restrict__ mel_hi, float* __restrict__ mel, int T, int ps,
                                                     int n_mels, int mode) {
  __shared__ float db[FR][kBins + 3];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * FR;
  const int tid = threadIdx.x;
  const float maxdb = (mode == 0) ? to_db(__uint_as_float(pmax[b])) : 0.0f;
  for (int i = tid; i < FR * kBins; i += 256) {
    const int f = i / kBins, k = i - f * kBins;
    const int t = t0 + f;
    float vdb = 0.0f;
    if (t < T) {
      const float p = power[((size_t)b * T + t) * ps + k];
      if (mode == 0) {
        vdb = fmaxf(to_db(p) - maxdb, -80.0f);
      } else {
        vdb = logf(fmaxf(p, 1e-10f)) / 2.302585093f;
      }
    }
    db[f][k] = vdb;
  }
  __syncthreads();
  for (int o = tid; o < FR * n_mels; o += 256) {
    const int f = o / n_mels, m = o - f * n_mels;
    const int t = t0 + f;
    if (t >= T) continue;
    float acc = 0.0f;
    const int lo = mel_lo[m], hi = mel_hi[m];
    for (int k = lo; k < hi; ++k) acc = fmaf(db[f][k], melw[k * n_mels + m], acc);
    mel[((size_t)b * T + t) * n_mels + m] = acc;
  }
}


// Second-generation dB + mel kernel.  FR frames per CTA; every thread issues all of its power loads before the first use
// (the first version exposed one DRAM round trip per element), dB through lg2.approx (10 log10 p = 3.0103 log2 p; the
// approximation error is < 1e-5 dB against the 5e-3 dB test tolerance), and the sparse triangular mel filters are staged in
// shared memory as compact band weights before griddepcontrol.wait (they are constants).
constexpr int kMelMaxNnz = 1536, kMelMaxFilters = 128;
// dB differences are formed in the log2 domain, (log2 p - log2 pmax) * 3.0103: the subtraction of equal values is exactly 0
// (a product-then-subtract form lets the compiler contract one side into an FMA and breaks "the loudest bin is 0 dB").
__device__ __forceinline__ float log2_floor(float p) { return __log2f(fmaxf(p, 1e-10f)); }

template <int FR>
__global__ void __launch_bounds__(256) db_mel_fast_kernel(const float* __restrict__ power, const unsigned int* __restrict__ pmax,
                                                          const int* __restrict__ mel_lo, const int* __restrict__ mel_off,
                                                          const float* __restrict__ mel_wc, int nnz, float* __restrict__ mel, int T, int ps,
                                                          int n_mels, int mode) {
  __shared__ float db[FR][kBins + 3];
  __shared__ float wc_s[kMelMaxNnz];
  __shared__ int lo_s[kMelMaxFilters], off_s[kMelMaxFilters + 1];
  const int tid = threadIdx.x;
  pdl_trigger();
  for (int i = tid; i < nnz; i += 256) wc_s[i] = mel_wc[i];
  for (int i = tid; i < n_mels; i += 256) lo_s[i] = mel_lo[i];
  for (int i = tid; i <= n_mels; i += 256) off_s[i] = mel_off[i];
  pdl_wait();
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * FR;
  const float maxlg = (mode == 0) ? log2_floor(__uint_as_float(pmax[b])) : 0.0f;
  // frame-major / bin-minor: bin k = tid + 256 j (j = 0, 1, 2) of every frame, so no index needs a division (the first version
  // spent most of its 34 M warp instructions on i / 513 and 64-bit address arithmetic); all FR * 3 loads are issued before use
  constexpr int NJ = (kBins + 255) / 256;
  float pv[FR][NJ];
  const float* prow = power + ((size_t)b * T + t0) * ps + tid;
#pragma unroll
  for (int f = 0; f < FR; ++f)
#pragma unroll
    for (int j = 0; j < NJ; ++j)
      pv[f][j] = (tid + 256 * j < kBins && t0 + f < T) ? __ldg(prow + (size_t)f * ps + 256 * j) : 1.0f;
#pragma unroll
  for (int f = 0; f < FR; ++f)
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int k = tid + 256 * j;
      if (k < kBins) {
        float vdb;
        if (mode == 0) vdb = fmaxf(3.0102999566398120f * (log2_floor(pv[f][j]) - maxlg), -80.0f);
        else vdb = 0.30102999566398120f * log2_floor(pv[f][j]);
        db[f][k] = vdb;
      }
    }
  __syncthreads();
  for (int o = tid; o < FR * n_mels; o += 256) {
    const int f = o / n_mels, m = o - f * n_mels;
    const int t = t0 + f;
    if (t >= T) continue;
    const int lo = lo_s[m], w0 = off_s[m], n = off_s[m + 1] - w0;
    float acc = 0.0f;
#pragma unroll 4
    for (int k = 0; k < n; ++k) acc = fmaf(db[f][lo + k], wc_s[w0 + k], acc);   // (same summation order as the reference's dot product)
    mel[((size_t)b * T + t) * n_mels + m] = acc;
  }
}

int launch_frontend(const FrontendParams& p, cudaStream_t stream) {
  B200_CUDA_OK(cudaMemsetAsync(p.pmax, 0, sizeof(unsigned int) * p.B, stream));
  static int legacy = -1;
  if (legacy < 0) {
    const char* e = getenv("B200ASR_STFT_LEGACY");
    legacy = (e && e[0] == '1') ? 1 : 0;
  }
  dim3 g1(ceil_div(p.T, 2), p.B);
  if (!legacy) {
    const size_t smem = sizeof(float2) * kNfft + sizeof(float) * kNfft + sizeof(float2) * kSwTile * kSwWarps;
    static PerDeviceSmem configured;
    if (configured.need(smem)) B200_CUDA_OK(cudaFuncSetAttribute(stft_power_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 gw(ceil_div(ceil_div(p.T, 2), kSwWarps * kSwPairs), p.B);
    B200_CUDA_OK(launch_k(stft_power_warp_kernel, gw, dim3(kSwWarps * 32), smem, stream, p.wav, p.window, p.twiddle, p.power, p.pmax, p.L, p.T,
                          p.pad_left, p.hop, p.power_stride));
  } else {
    B200_CUDA_OK(launch_k(stft_power_kernel, g1, dim3(256), 0, stream, p.wav, p.window, p.twiddle, p.power, p.pmax, p.L, p.T, p.pad_left,
                          p.hop, p.power_stride));
  }
  if (!legacy && p.mel_wc != nullptr && p.mel_nnz <= kMelMaxNnz && p.n_mels <= kMelMaxFilters) {
    constexpr int FR = 8;
    dim3 g2(ceil_div(p.T, FR), p.B);
    B200_CUDA_OK(launch_k(db_mel_fast_kernel<FR>, g2, dim3(256), 0, stream, (const float*)p.power, (const unsigned int*)p.pmax, p.mel_lo,
                          p.mel_off, p.mel_wc, p.mel_nnz, p.mel, p.T, p.power_stride, p.n_mels, p.mode));
  } else {
    constexpr int FR = 4;
    dim3 g2(ceil_div(p.T, FR), p.B);
    B200_CUDA_OK(launch_k(db_mel_kernel<FR>, g2, dim3(256), 0, stream, (const float*)p.power, (const unsigned int*)p.pmax, p.melw, p.mel_lo,
                          p.mel_hi, p.mel, p.T, p.power_stride, p.n_mels, p.mode));
  }
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace b200asr
