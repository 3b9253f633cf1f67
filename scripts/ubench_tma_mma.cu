// Micro-benchmarks behind DESIGN.md's kernel analysis: what paces the weight-streaming tcgen05 kernels (gemm_chain / gemm_tc)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o scripts/ubench_tma_mma scripts/ubench_tma_mma.cu
// Each test runs G CTAs (one per SM) in lock step on the SAME weight matrix (as the real kernels do) and reports cycles per
// 18 KB weight slab (= 144 rows x 32 fp32, the unit gemm_chain streams) seen by CTA 0.
//   tma2d   : ring of S stages filled by 2-D tensor-map loads (box {32, 144}: 144 rows of 128 B, SWIZZLE_128B), consumer frees at once
//   bulk1d  : same ring filled by ONE contiguous 18432-byte cp.async.bulk per slab (weights pre-tiled in global memory)
//   mma_ss  : 4 x tcgen05.mma kind::tf32 M=128 N=144 K=8 per "slab", A and B from shared memory, no loads at all
//   mma_ts  : same with A from tensor memory
//   chain_* : loads + MMAs together (the gemm_chain inner loop without epilogues)
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

constexpr int kRows = 144, kSlabBytes = kRows * 128, kMaxStages = 6;

struct UParams {
  int mode;       // bit0: use loads; bit1: loads are 1-D bulk; bit2: issue MMAs; bit3: A from TMEM
  int stages, nslabs, n_mma;
  const float* packed;   // pre-tiled weights (1-D mode)
  long long* out;        // [grid] cycles
};

__global__ void __launch_bounds__(128, 1) ubench_kernel(const __grid_constant__ CUtensorMap map_w, const UParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* xs = smem;                          // 5 slabs of A (128 rows x 128 B)
  uint8_t* ring = xs + 5 * 128 * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + kMaxStages * kSlabBytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* done_bar = empty_bar + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (5 * 128 * 128 + kMaxStages * kSlabBytes) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(done_bar, 1);
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
  const bool use_loads = p.mode & 1, bulk = p.mode & 2, use_mma = p.mode & 4, a_tmem = p.mode & 8;
  const long long t0 = clock64();
  if (warp == 0 && lane == 0 && use_loads) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < p.nslabs; ++i) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      mbar_expect_tx(&full_bar[stage], kSlabBytes);
      const int chunk = (i / 5) & 3, kb = i % 5;
      if (!bulk) {
        tma_load_2d(&map_w, &full_bar[stage], ring + stage * kSlabBytes, kb * 32, chunk * kRows);
      } else {
        const float* src = p.packed + (size_t)(chunk * 5 + kb) * (kSlabBytes / 4);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + stage * kSlabBytes)),
                     "l"(src), "r"(kSlabBytes), "r"(smem_u32(&full_bar[stage]))
                     : "memory");
      }
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(128, kRows);
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < p.nslabs; ++i) {
      if (use_loads) { mbar_wait(&full_bar[stage], phase); tcgen05_fence_after(); }
      if (lane == 0) {
        if (use_mma) {
          const uint64_t da = make_smem_desc(smem_u32(xs + (i % 5) * 128 * 128));
          const uint64_t db = make_smem_desc(smem_u32(ring + stage * kSlabBytes));
          for (int k = 0; k < p.n_mma; ++k) {
            if (a_tmem) umma_tf32_ts(tmem_base + 288, tmem_base + (uint32_t)(8 * ((i * 4 + k) % 18)), db + (uint64_t)(2 * (k & 3)), idesc, 1u);
            else umma_tf32(tmem_base + 288, da + (uint64_t)(2 * (k & 3)), db + (uint64_t)(2 * (k & 3)), idesc, 1u);
          }
          if (use_loads) tcgen05_commit(&empty_bar[stage]);
        } else if (use_loads) {
          mbar_arrive(&empty_bar[stage]);
        }
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    if (lane == 0 && use_mma) tcgen05_commit(done_bar);
    if (use_mma) { mbar_wait(done_bar, 0); tcgen05_fence_after(); }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) p.out[blockIdx.x] = t1 - t0;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
  const int N1 = 576, K1 = 144;
  float *W, *packed;
  CK(cudaMalloc(&W, sizeof(float) * N1 * K1));
  CK(cudaMalloc(&packed, 20 * kSlabBytes));
  CK(cudaMemset(W, 0, sizeof(float) * N1 * K1));
  CK(cudaMemset(packed, 0, 20 * kSlabBytes));
  long long* out;
  CK(cudaMalloc(&out, sizeof(long long) * 148));
  TcContext ctx;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  ctx.encode_tiled = fn;
  CUtensorMap mw;
  const cuuint64_t dims[2] = {(cuuint64_t)K1, (cuuint64_t)N1};
  const cuuint64_t strides[1] = {(cuuint64_t)K1 * 4};
  const cuuint32_t box[2] = {32, kRows}, ones[2] = {1, 1};
  if (encode_map(ctx, &mw, W, 2, dims, strides, box, ones)) { printf("encode failed: %s\n", g_errbuf); return 1; }
  const size_t smem = 5 * 128 * 128 + kMaxStages * kSlabBytes + 1024 + 256;
  CK(cudaFuncSetAttribute(ubench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  struct Case { const char* name; int mode, stages, n_mma; };
  const Case cases[] = {
      {"tma2d   S=1", 1, 1, 0},       {"tma2d   S=2", 1, 2, 0},       {"tma2d   S=6", 1, 6, 0},
      {"bulk1d  S=1", 3, 1, 0},       {"bulk1d  S=2", 3, 2, 0},       {"bulk1d  S=6", 3, 6, 0},
      {"mma_ss  4/slab", 4, 6, 4},    {"mma_ts  4/slab", 12, 6, 4},
      {"chain_ss tma2d S=6", 5, 6, 4}, {"chain_ts tma2d S=6", 13, 6, 4},
      {"chain_ss bulk1d S=6", 7, 6, 4}, {"chain_ts bulk1d S=6", 15, 6, 4},
      {"chain_ts tma2d S=3", 13, 3, 4}, {"chain_ts bulk1d S=3", 15, 3, 4},
  };
  const int grids[3] = {1, 63, 148};
  const int nslabs = 200;
  printf("%-24s %8s %14s %14s\n", "case", "grid", "cyc/slab(cta0)", "cyc/slab(max)");
  for (const Case& c : cases) {
    for (int g : grids) {
      UParams p{c.mode, c.stages, nslabs, c.n_mma, packed, out};
      for (int rep = 0; rep < 2; ++rep) {
        ubench_kernel<<<g, 128, smem>>>(mw, p);
        CK(cudaDeviceSynchronize());
      }
      std::vector<long long> h(g);
      CK(cudaMemcpy(h.data(), out, sizeof(long long) * g, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (long long v : h) mx = v > mx ? v : mx;
      printf("%-24s %8d %14.1f %14.1f\n", c.name, g, (double)h[0] / nslabs, (double)mx / nslabs);
    }
  }
  return 0;
}
