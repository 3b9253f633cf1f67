// tcgen05.mma kind::tf32 pacing: latency and back-to-back throughput of M=128 x N x K=8 instructions issued by one thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o scripts/ubench_mma scripts/ubench_mma.cu
// For each (N, operand form, accumulator pattern, commit pattern) the issuing thread reads clock64, issues n MMAs, commits to an
// mbarrier, waits for it and reads clock64 again: cycles(n) = latency + n * per-MMA cost.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace b200asr { thread_local char g_errbuf[512]; bool g_pdl_enabled = false; }
#include "../tensorflowasr_b200/csrc/tc_common.cuh"

using namespace b200asr;
using namespace b200asr::tc;

struct MParams {
  int n_mma, a_tmem, alt_d, commit_every, N;
  long long* out;
};

template <int N>
__device__ __forceinline__ void run_case(const MParams& p, uint32_t tmem_base, uint8_t* xs, uint8_t* bs, uint64_t* done_bar, uint64_t* junk_bar,
                                         uint32_t parity) {
  constexpr uint32_t idesc = make_idesc(128, N);
  const uint64_t da = make_smem_desc(smem_u32(xs));
  const uint64_t db = make_smem_desc(smem_u32(bs));
  const long long t0 = clock64();
  for (int i = 0; i < p.n_mma; ++i) {
    const uint32_t dd = (p.alt_d && (i & 1)) ? tmem_base + 64 : tmem_base + 256;   // accumulators at columns [256, 256+N) / [64, 64+N)
    if (p.a_tmem) umma_tf32_ts(dd, tmem_base + (uint32_t)(8 * (i & 7)), db + (uint64_t)(2 * (i & 3)), idesc, 1u);   // A: columns [0, 64)
    else umma_tf32(dd, da + (uint64_t)(2 * (i & 3)), db + (uint64_t)(2 * (i & 3)), idesc, 1u);
    if (p.commit_every > 0 && (i % p.commit_every) == p.commit_every - 1) tcgen05_commit(junk_bar);
  }
  tcgen05_commit(done_bar);
  mbar_wait(done_bar, parity);
  const long long t1 = clock64();
  p.out[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(64, 1) ubench_mma_kernel(const MParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* xs = smem;                    // A: 128 rows x 128 B
  uint8_t* bs = xs + 128 * 128;          // B: 256 rows x 128 B
  uint64_t* done_bar = reinterpret_cast<uint64_t*>(bs + 256 * 128);
  uint64_t* junk_bar = done_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(junk_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (128 * 128 + 256 * 128) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) {
    mbar_init(done_bar, 1);
    mbar_init(junk_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 1 && lane == 0) {
    // untimed warm-up (first-use effects), then the timed run
    MParams w = p;
    w.n_mma = 4;
    switch (p.N) {
      case 64: run_case<64>(w, tmem_base, xs, bs, done_bar, junk_bar, 0); run_case<64>(p, tmem_base, xs, bs, done_bar, junk_bar, 1); break;
      case 96: run_case<96>(w, tmem_base, xs, bs, done_bar, junk_bar, 0); run_case<96>(p, tmem_base, xs, bs, done_bar, junk_bar, 1); break;
      case 144: run_case<144>(w, tmem_base, xs, bs, done_bar, junk_bar, 0); run_case<144>(p, tmem_base, xs, bs, done_bar, junk_bar, 1); break;
      default: run_case<256>(w, tmem_base, xs, bs, done_bar, junk_bar, 0); run_case<256>(p, tmem_base, xs, bs, done_bar, junk_bar, 1); break;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
  long long* out;
  CK(cudaMalloc(&out, sizeof(long long) * 148));
  const size_t smem = 128 * 128 + 256 * 128 + 1024 + 64;
  CK(cudaFuncSetAttribute(ubench_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int Ns[4] = {64, 96, 144, 256};
  const int ns[6] = {1, 4, 16, 64, 256, 1024};
  printf("%-6s %-4s %-6s %-8s |", "N", "A", "D", "commit");
  for (int n : ns) printf(" n=%-7d", n);
  printf(" | cyc/MMA (n=1024 vs 64)\n");
  for (int N : Ns)
    for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
      for (int alt = 0; alt < 2; ++alt)
        for (int ce : {0, 4}) {
          if (N == 256 && alt) continue;                 // two 256-column accumulators + the TS operand do not fit side by side
          printf("%-6d %-4s %-6s %-8s |", N, a_tmem ? "tmem" : "smem", alt ? "alt2" : "same", ce ? "every4" : "end");
          double c64 = 0, c1024 = 0;
          for (int n : ns) {
            MParams p{n, a_tmem, alt, ce, N, out};
            ubench_mma_kernel<<<1, 64, smem>>>(p);
            CK(cudaDeviceSynchronize());
            long long h;
            CK(cudaMemcpy(&h, out, sizeof(h), cudaMemcpyDeviceToHost));
            printf(" %-9lld", h);
            if (n == 64) c64 = (double)h;
            if (n == 1024) c1024 = (double)h;
          }
          printf(" | %.1f\n", (c1024 - c64) / (1024 - 64));
        }
  return 0;
}
