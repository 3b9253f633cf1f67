// Warp-private 1024-point FFT carrying two real STFT frames (frame 2i in the real lane, frame 2i+1 in the imaginary lane).
//
// 1024 = 32 x 32 (four-step FFT): with n = 32*n1 + n2 and k = k1 + 32*k2,
//   X[k1 + 32 k2] = sum_{n2} W_1024^{n2 k1} W_32^{n2 k2} [ sum_{n1} x[32 n1 + n2] W_32^{n1 k1} ]
// pass A: lane = n2 holds x[32 n1 + lane] for all n1 in registers, one 32-point FFT (fft32_gen.cuh), twiddle by
//         W_1024^{lane k1}, transposed through a warp-private shared-memory tile (row stride 33 float2: conflict-free);
// pass B: lane = k1 holds the 32 values over n2, second 32-point FFT -> X[lane + 32 k2];
// untangle: Z = FFT(a + i b)  =>  A[k] = (Z[k] + conj Z[N-k]) / 2,  B[k] = (Z[k] - conj Z[N-k]) / (2i); |.|^2 of both.
// No block-wide barrier is needed: the exchange buffer belongs to one warp (__syncwarp only).
// The phases are plain functions of (lane, buffers) so that a host harness (tests/test_host.py -> scripts/stft_warp_host.cu)
// can run the very same code lane by lane against numpy.fft.
#pragma once
#include <cuda_runtime.h>

#include "fft32_gen.cuh"

namespace b200asr {

constexpr int kSwStride = 33;                      // float2 per row of the exchange tile
constexpr int kSwTile = 32 * kSwStride;            // float2 per warp (>= 1024: reused as the linear spectrum Z[0..1023])

// pass A for one lane.  w: this utterance's samples [L]; s0: sample index of n = 0 of frame A (may be negative: 'same' padding);
// frame B starts `hop` samples later (absent when !has_second).  win: Hann window [1024]; tw: exp(-2 pi i m / 1024) [1024].
B200_HD void stft_pass_a(int lane, const float* __restrict__ w, int L, int s0, int hop, bool has_second, const float* win,
                         const float2* tw, float2* sbuf) {
  float2 v[32];
#pragma unroll
  for (int n1 = 0; n1 < 32; ++n1) {
    const int n = 32 * n1 + lane;
    const int sa = s0 + n, sb = sa + hop;
    const float wn = win[n];
    const float xa = (sa >= 0 && sa < L) ? w[sa] : 0.0f;
    const float xb = (has_second && sb >= 0 && sb < L) ? w[sb] : 0.0f;
    v[n1] = make_float2(xa * wn, xb * wn);
  }
  fft32_dif(v);
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) {
    float2 y = v[B200_BITREV32(k1)];
    if (k1 > 0) {
      const float2 t = tw[(lane * k1) & 1023];
      y = make_float2(y.x * t.x - y.y * t.y, y.x * t.y + y.y * t.x);
    }
    sbuf[k1 * kSwStride + lane] = y;
  }
}

// pass B, part 1: lane = k1 gathers its row and transforms it; v[bitrev(k2)] = Z[lane + 32 k2]
B200_HD void stft_pass_b_load(int lane, const float2* sbuf, float2 (&v)[32]) {
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) v[n2] = sbuf[lane * kSwStride + n2];
  fft32_dif(v);
}
// pass B, part 2 (after every lane has finished part 1): the spectrum in natural order, Z[k] at zbuf[k]
B200_HD void stft_pass_b_store(int lane, const float2 (&v)[32], float2* zbuf) {
#pragma unroll
  for (int k2 = 0; k2 < 32; ++k2) zbuf[lane + 32 * k2] = v[B200_BITREV32(k2)];
}

// untangle + power for bins k = lane + 32 j (j = 0..16, k <= 512); returns the largest power this lane produced
B200_HD float stft_untangle(int lane, const float2* zbuf, float* pa, float* pb, bool has_second) {
  float vmax = 0.0f;
#pragma unroll
  for (int j = 0; j <= 16; ++j) {
    const int k = lane + 32 * j;
    if (k <= 512) {
      const float2 z = zbuf[k];
      const float2 zc = zbuf[(1024 - k) & 1023];
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
  }
  return vmax;
}

}  // namespace b200asr
