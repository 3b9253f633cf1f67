// Host harness for csrc/stft_warp.cuh: runs the warp-FFT phases lane by lane on the CPU (same source the kernel compiles)
// and compares both frames' power spectra with a direct double-precision DFT.  Prints "max_rel_err <e>"; exit 0 when < 2e-5
// (relative to the largest bin of the frame).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../tensorflowasr_b200/csrc/stft_warp.cuh"

using namespace b200asr;

int main() {
  const int L = 5000, hop = 160;
  std::vector<float> wav(L), win(1024);
  std::vector<float2> tw(1024);
  unsigned s = 12345u;
  for (int i = 0; i < L; ++i) {
    s = s * 1664525u + 1013904223u;
    wav[i] = ((s >> 8) & 0xffff) / 65536.0f - 0.5f + 0.3f * sinf(0.05f * i);
  }
  for (int n = 0; n < 1024; ++n) {
    win[n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * n / 1024.0));
    tw[n] = make_float2((float)cos(-2.0 * M_PI * n / 1024.0), (float)sin(-2.0 * M_PI * n / 1024.0));
  }
  double worst = 0.0;
  const int starts[3] = {-432, 700, 4100};   // left padding, interior, right padding
  for (int c = 0; c < 3; ++c) {
    for (int second = 1; second >= 0; --second) {
      const int s0 = starts[c];
      std::vector<float2> sbuf(kSwTile);
      std::vector<float> pa(513, -1.f), pb(513, -1.f);
      for (int lane = 0; lane < 32; ++lane) stft_pass_a(lane, wav.data(), L, s0, hop, second != 0, win.data(), tw.data(), sbuf.data());
      float2 v[32][32];
      for (int lane = 0; lane < 32; ++lane) stft_pass_b_load(lane, sbuf.data(), v[lane]);
      for (int lane = 0; lane < 32; ++lane) stft_pass_b_store(lane, v[lane], sbuf.data());
      float vmax = 0.f;
      for (int lane = 0; lane < 32; ++lane) vmax = fmaxf(vmax, stft_untangle(lane, sbuf.data(), pa.data(), pb.data(), second != 0));
      double rmax_all = 0.0;
      for (int f = 0; f <= second; ++f) {
        std::vector<double> ref(513);
        double rmax = 0.0;
        for (int k = 0; k <= 512; ++k) {
          double re = 0, im = 0;
          for (int n = 0; n < 1024; ++n) {
            const int si = s0 + f * hop + n;
            const double x = (si >= 0 && si < L) ? (double)wav[si] * (double)win[n] : 0.0;
            re += x * cos(2.0 * M_PI * k * n / 1024.0);
            im -= x * sin(2.0 * M_PI * k * n / 1024.0);
          }
          ref[k] = re * re + im * im;
          rmax = fmax(rmax, ref[k]);
        }
        const std::vector<float>& got = f ? pb : pa;
        for (int k = 0; k <= 512; ++k) worst = fmax(worst, fabs((double)got[k] - ref[k]) / rmax);
        rmax_all = fmax(rmax_all, rmax);
      }
      if (fabs((double)vmax - rmax_all) / rmax_all > 1e-4) { printf("vmax mismatch %g %g\n", vmax, rmax_all); return 2; }
    }
  }
  printf("max_rel_err %.3e\n", worst);
  return worst < 2e-5 ? 0 : 1;
}
