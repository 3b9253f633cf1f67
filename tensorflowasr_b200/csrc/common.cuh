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

// ------------------------------------------------------------------------------------------------ programmatic dependent launch
// Every kernel of the schedule is launched with cudaLaunchAttributeProgrammaticStreamSerialization (also inside the captured
// CUDA graph): a kernel calls pdl_trigger() once its prologue resources are taken (TMEM allocated), which lets the NEXT
// kernel's CTAs start and run THEIR prologue (barrier init, TMEM alloc, tensor-map prefetch, parameter caches) under this
// kernel's main body; pdl_wait() then blocks until the previous grid has completed and its writes are visible.  Rule: no
// global-memory access that depends on (or could disturb) an earlier kernel before pdl_wait(); trigger only AFTER tcgen05.alloc
// (a dependent CTA must never hold TMEM that a not-yet-allocated CTA of the primary still needs).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

extern bool g_pdl_enabled;   // engine.cu; B200ASR_NO_PDL=1 in the environment turns the launch attribute off

// cudaFuncSetAttribute (dynamic shared memory size) is per DEVICE: a kernel's launcher keeps one of these per instantiation and
// sets the attribute once for every device an engine is created on (several engines on different GPUs may live in one process)
struct PerDeviceSmem {
  size_t set[64] = {};
  // true when the attribute must be (re)applied for `bytes` on the current device
  bool need(size_t bytes) {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (set[d] >= bytes && bytes > 0) return false;
    set[d] = bytes;
    return true;
  }
};

// cluster_x > 1: thread-block cluster of that many CTAs along x (grid.x must be a multiple of it)
template <class... KArgs, class... Args>
inline cudaError_t launch_k_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                                    Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (g_pdl_enabled) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  return launch_k_cluster(kern, grid, block, smem, stream, 1, static_cast<Args&&>(args)...);
}

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

// round-to-nearest (ties away) to the 10-bit tf32 mantissa on the integer ALU.  The tcgen05 kind::tf32 datapath TRUNCATES whatever
// fp32 bits it is handed (a -2^-11 relative bias per operand); every tensor that is only ever read as a tensor-core operand is
// therefore stored already rounded by the kernel that produces it (scripts/tf32_error_study.py: logits error 0.109 -> 0.014).
__device__ __forceinline__ float tf32_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

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
