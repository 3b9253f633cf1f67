// Shared helpers for the B200 (sm_100a) Conformer-CTC kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

namespace b200asr {

constexpr int kWarp = 32;

#define B200_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      snprintf(g_errbuf, sizeof(g_errbuf), "%s:%d %s -> %s", __FILE__, __LINE__, #expr,      \
               cudaGetErrorString(_e));                                                      \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

extern thread_local char g_errbuf[512];

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// TensorFlow "SAME" padding (== ONNX SAME_UPPER): out = ceil(in/stride), pad_before = total/2.
struct SamePad {
  int out, before, after;
};
__host__ __device__ inline SamePad same_pad(int n_in, int k, int stride) {
  SamePad p;
  p.out = (n_in + stride - 1) / stride;
  int total = (p.out - 1) * stride + k - n_in;
  if (total < 0) total = 0;
  p.before = total / 2;
  p.after = total - p.before;
  return p;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// x * sigmoid(x) and sigmoid with full-precision expf (parity with the reference's fp32 graph matters more than
// the last few % of SFU throughput here; these sit in GEMM epilogues that are not SFU bound).
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float swishf_(float x) { return x / (1.0f + expf(-x)); }

}  // namespace b200asr
