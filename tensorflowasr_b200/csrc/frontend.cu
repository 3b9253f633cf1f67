// Mel/dB frontend for sm_100a: wav -> |STFT|^2 -> dB (per-utterance max, -80 dB floor) -> mel projection.
//
// Replaces asr/models/layers/time_frequency.py:100-122,173-189 + backend_keras.py:5-23 of the reference, which
// evaluates the STFT as two dense 1024-tap strided convolutions (2.1 GFLOP / 10 s utterance).  Here each CTA runs
// one 1024-point complex radix-4 Stockham FFT in shared memory that carries TWO real frames (frame 2i in the
// real lane, frame 2i+1 in the imaginary lane) and untangles them with the Hermitian symmetry, i.e. ~26 MFLOP per
// utterance.  HBM-bound by design: reads the waveform once (frames overlap in L1/L2), writes the power
// spectrogram once; the dB + mel kernel re-reads it once after the per-utterance max is known.
#include "kernels.cuh"

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

int launch_frontend(const FrontendParams& p, cudaStream_t stream) {
  B200_CUDA_OK(cudaMemsetAsync(p.pmax, 0, sizeof(unsigned int) * p.B, stream));
  dim3 g1(ceil_div(p.T, 2), p.B);
  stft_power_kernel<<<g1, 256, 0, stream>>>(p.wav, p.window, p.twiddle, p.power, p.pmax, p.L, p.T, p.pad_left, p.hop,
                                             p.power_stride);
  constexpr int FR = 4;
  dim3 g2(ceil_div(p.T, FR), p.B);
  db_mel_kernel<FR><<<g2, 256, 0, stream>>>(p.power, p.pmax, p.melw, p.mel_lo, p.mel_hi, p.mel, p.T, p.power_stride,
                                            p.n_mels, p.mode);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace b200asr
