// Shared device/host helpers of the tcgen05 kernels (gemm_tc.cu, gemm_chain.cu): PTX wrappers for mbarrier / TMA /
// tcgen05 (mma, ld, st, commit, fences), UMMA descriptors, fused epilogues, tensor-map encoding.
#pragma once
#include "gemm_tc.cuh"
#include <cuda_fp16.h>

#include <cuda.h>
#include <cudaTypedefs.h>

#include <type_traits>

namespace b200asr {
namespace tc {

constexpr int BLOCK_K = 32;        // fp32 elements = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 8;          // tf32
constexpr int kStages = 4;
constexpr int kThreads = 192;      // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kTmemCols = 512;
constexpr unsigned kSpinLimit = 1u << 28;

struct TcParams {
  const float* bias;
  const float* resid;
  float* C;
  int M, N, K, ldc;
  float alpha;
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  float* C2;
  float ln_eps;
  int num_m_tiles, num_n_tiles, num_k_blocks;
  // conv2 implicit GEMM
  int a_mode, T2, F2, D, bt, kc, pad_t, pad_f, tiles_per_b;
  int round_out;   // plain epilogues: store C rounded to nearest tf32 (it is only read as a tensor-core operand again)
  int f16;         // A and B operands hold IEEE fp16 (kind::f16, 64 columns per 128-byte swizzle row, K = 16 per instruction)
  int out_f16;     // plain epilogues: C holds IEEE fp16 [M, ldc] (it is only read as an fp16 tensor-core operand again)
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  unsigned spins = 0;
  while (true) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (++spins > kSpinLimit) {  // a protocol bug must never hang the GPU: fail the launch instead
      printf("b200asr gemm_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T   (both K-major, tf32)
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// the same with fp16 operands (kind::f16: K = 16 per instruction = the same 32 bytes of a swizzle row)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address  bits [0,14)
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major) bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset = 1024 B  bits [32,46)
  d |= (uint64_t)1 << 46;                    // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3 at [17,23), M>>4 at [24,29) (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D=f32, A=B=f16 (format code 0)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// SFU-approximate activations for the tensor-core path (ex2.approx + rcp.approx, ~1e-6 relative: far below tf32 input rounding)
// branch-free: ex2.approx + rcp.approx (IEEE __frcp_rn carries a per-element slow-path branch + call that serialises the
// epilogue: measured 168 cycles/element in the chained kernel)
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// sigmoid(x) = 0.5 + 0.5 tanh(x/2): ONE MUFU op per element (tanh.approx, ~2^-11 relative: at the tf32 input-rounding level)
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
__device__ __forceinline__ float swish_fast(float x) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(h), h);
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
          taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// round-to-nearest (ties away) to the 10-bit tf32 mantissa on the integer ALU (cvt.rna.tf32 sits on a slow conversion pipe)
__device__ __forceinline__ uint32_t tf32_rn_bits(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
// D[tmem] (+)= A[tmem] . B[smem]^T   (A: 128 lanes x K columns of tf32, B K-major)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ------------------------------------------------------------------------------------------------ epilogue
// In the TMEM accumulator layout a thread owns one output ROW, so naive global loads/stores touch 32 different cache lines
// per warp instruction (measured: the LSU wavefront rate, not HBM or math, bounded every epilogue).  All global traffic of
// the epilogues therefore goes through a warp-private shared-memory transpose: a 32-row x W-column tile is exchanged so that
// each warp instruction moves whole 128-byte (W=32) / 64-byte (W=16) row segments.
constexpr int kWsmLd = 36;                       // floats per row of the per-warp scratch tile (16-byte aligned, conflict-free)
constexpr int kWsmFloats = 32 * kWsmLd;          // per warp

// v[0..W) = W consecutive columns of this lane's row  ->  global rows grow0.. (coalesced).  nrows/ncols clip the tile.
template <int W>
__device__ __forceinline__ void warp_tile_store(float* wsm, const float* v, float* gtile, size_t ld, int nrows, int ncols, int lane) {
#pragma unroll
  for (int q = 0; q < W / 4; ++q) *reinterpret_cast<float4*>(wsm + lane * kWsmLd + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  __syncwarp();
  constexpr int CH = W / 4, RPI = 32 / CH;
  const int q = lane % CH;
#pragma unroll
  for (int i = 0; i < 32 / RPI; ++i) {
    const int r = i * RPI + lane / CH;
    if (r < nrows && 4 * q < ncols) *reinterpret_cast<float4*>(gtile + (size_t)r * ld + 4 * q) = *reinterpret_cast<const float4*>(wsm + r * kWsmLd + 4 * q);
  }
  __syncwarp();
}
// the same with an fp16 destination (row segments of 2 * W bytes)
template <int W>
__device__ __forceinline__ void warp_tile_store_h(float* wsm, const float* v, __half* gtile, size_t ld, int nrows, int ncols, int lane) {
#pragma unroll
  for (int q = 0; q < W / 4; ++q) *reinterpret_cast<float4*>(wsm + lane * kWsmLd + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  __syncwarp();
  constexpr int CH = W / 4, RPI = 32 / CH;
  const int q = lane % CH;
#pragma unroll
  for (int i = 0; i < 32 / RPI; ++i) {
    const int r = i * RPI + lane / CH;
    if (r < nrows && 4 * q < ncols) {
      const float4 t = *reinterpret_cast<const float4*>(wsm + r * kWsmLd + 4 * q);
      const __half2 h01 = __floats2half2_rn(t.x, t.y), h23 = __floats2half2_rn(t.z, t.w);
      uint2 pk;
      pk.x = *reinterpret_cast<const unsigned int*>(&h01);
      pk.y = *reinterpret_cast<const unsigned int*>(&h23);
      *reinterpret_cast<uint2*>(gtile + (size_t)r * ld + 4 * q) = pk;
    }
  }
  __syncwarp();
}
template <int W>
__device__ __forceinline__ void warp_tile_load(float* wsm, float* v, const float* gtile, size_t ld, int nrows, int ncols, int lane) {
  constexpr int CH = W / 4, RPI = 32 / CH;
  const int q = lane % CH;
#pragma unroll
  for (int i = 0; i < 32 / RPI; ++i) {
    const int r = i * RPI + lane / CH;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nrows && 4 * q < ncols) t = *reinterpret_cast<const float4*>(gtile + (size_t)r * ld + 4 * q);
    *reinterpret_cast<float4*>(wsm + r * kWsmLd + 4 * q) = t;
  }
  __syncwarp();
#pragma unroll
  for (int qq = 0; qq < W / 4; ++qq) {
    const float4 t = *reinterpret_cast<const float4*>(wsm + lane * kWsmLd + 4 * qq);
    v[4 * qq] = t.x; v[4 * qq + 1] = t.y; v[4 * qq + 2] = t.z; v[4 * qq + 3] = t.w;
  }
  __syncwarp();
}

// Plain epilogues (bias / ReLU / swish / GLU / residual / none) over W accumulator columns starting at tile column c.
//   taddr: TMEM address of this warp's lanes, column 0 of the accumulator; gcol0: global column of tile column 0
//   grow0: global output row of lane 0; nrows: valid rows of this warp (0..32)
template <int EPI, int W>
__device__ __forceinline__ void epilogue_plain_group(const TcParams& p, uint32_t taddr, float* wsm, size_t grow0, int nrows, int gcol,
                                                     int c, int lane) {
  uint32_t raw[W];
#pragma unroll
  for (int j = 0; j < W / 16; ++j) tmem_ld16_nowait(taddr + (uint32_t)(c + 16 * j), raw + 16 * j);   // warp-collective
  float r[W];
  if (EPI == EPI_RESID) warp_tile_load<W>(wsm, r, p.resid + grow0 * p.ldc + gcol, p.ldc, nrows, p.N - gcol, lane);
  tmem_ld_wait();
  float v[W];
#pragma unroll
  for (int q = 0; q < W / 4; ++q) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (EPI != EPI_NONE && p.bias != nullptr && gcol + 4 * q < p.N) b = __ldg(reinterpret_cast<const float4*>(p.bias + gcol + 4 * q));
    v[4 * q + 0] = __uint_as_float(raw[4 * q + 0]) + b.x; v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + b.y;
    v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + b.z; v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + b.w;
  }
  if constexpr (EPI == EPI_GLU) {
    float o[W / 2];
#pragma unroll
    for (int i = 0; i < W / 2; ++i) o[i] = v[2 * i] * sigmoid_fast(v[2 * i + 1]);
    warp_tile_store<W / 2>(wsm, o, p.C + grow0 * p.ldc + (gcol >> 1), p.ldc, nrows, (p.N - gcol) >> 1, lane);
  } else {
#pragma unroll
    for (int i = 0; i < W; ++i) {
      if (EPI == EPI_BIAS_RELU) v[i] = fmaxf(v[i], 0.f);
      else if (EPI == EPI_BIAS_SWISH) v[i] = swish_fast(v[i]);
      else if (EPI == EPI_RESID) v[i] = r[i] + p.alpha * v[i];
      if (p.round_out) v[i] = __uint_as_float(tf32_rn_bits(v[i]));
    }
    if (p.out_f16) warp_tile_store_h<W>(wsm, v, reinterpret_cast<__half*>(p.C) + grow0 * p.ldc + gcol, p.ldc, nrows, p.N - gcol, lane);   // warp-uniform
    else warp_tile_store<W>(wsm, v, p.C + grow0 * p.ldc + gcol, p.ldc, nrows, p.N - gcol, lane);
  }
}

template <int EPI, int BLOCK_N>
__device__ __forceinline__ void epilogue_plain(const TcParams& p, uint32_t taddr, float* wsm, size_t grow0, int nrows, int gcol0, int lane) {
  constexpr int FULL = (BLOCK_N / 32) * 32;
#pragma unroll 1
  for (int c = 0; c < FULL; c += 32) {
    if (gcol0 + c < p.N) epilogue_plain_group<EPI, 32>(p, taddr, wsm, grow0, nrows, gcol0 + c, c, lane);   // warp-uniform
  }
  if (BLOCK_N % 32 != 0) {
    if (gcol0 + FULL < p.N) epilogue_plain_group<EPI, 16>(p, taddr, wsm, grow0, nrows, gcol0 + FULL, FULL, lane);
  }
}

// CTC head + per-frame argmax: the thread (== output row) keeps the running first maximum of acc + bias over this tile's valid
// columns and writes one (max, argmax) pair per (row, N tile); the logits never leave the SM.
template <int BLOCK_N>
__device__ __forceinline__ void epilogue_argmax(const TcParams& p, uint32_t taddr, size_t grow0, int nrows, int gcol0, int nt, int lane) {
  float best = -INFINITY;
  int bidx = 0;
#pragma unroll 1
  for (int c = 0; c < BLOCK_N; c += 16) {
    if (gcol0 + c >= p.N) break;                     // warp-uniform
    uint32_t raw[16];
    tmem_ld16_nowait(taddr + (uint32_t)c, raw);      // warp-collective
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int col = gcol0 + c + 4 * q;
      if (col < p.N) {                               // N % 4 == 0: the four columns are valid together
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        const float v0 = __uint_as_float(raw[4 * q + 0]) + b.x, v1 = __uint_as_float(raw[4 * q + 1]) + b.y;
        const float v2 = __uint_as_float(raw[4 * q + 2]) + b.z, v3 = __uint_as_float(raw[4 * q + 3]) + b.w;
        if (v0 > best) { best = v0; bidx = col; }
        if (v1 > best) { best = v1; bidx = col + 1; }
        if (v2 > best) { best = v2; bidx = col + 2; }
        if (v3 > best) { best = v3; bidx = col + 3; }
      }
    }
  }
  if (lane < nrows) reinterpret_cast<float2*>(p.C)[(grow0 + lane) * (size_t)p.num_n_tiles + nt] = make_float2(best, __int_as_float(bidx));
}

// ------------------------------------------------------------------------------------------------ LayerNorm epilogue through a TMA-staged tile
// The residual tile is fetched by TMA into shared memory while the MMAs run; the epilogue works on it in place (thread ==
// row, SWIZZLE_128B slabs of 32 columns: conflict-free 16-byte accesses) and the results leave through TMA stores.  No
// global-memory latency is exposed to the epilogue warps and every global transaction is a full line.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0),
               "r"(c1)
               : "memory");
}
// ---- thread-block-cluster helpers (gemm_chain_pair.cu)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t raddr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// bulk copy of `bytes` (multiple of 16) from this CTA's shared memory into another CTA of the cluster; completion (complete_tx)
// is signalled on an mbarrier of the DESTINATION CTA.  dst / bar are shared::cluster addresses (map_to_cta).
__device__ __forceinline__ void bulk_copy_to_cta(uint32_t dst_cluster, const void* src_local, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(smem_u32(src_local)), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope (remote writers)
  const uint32_t addr = smem_u32(bar);
  unsigned spins = 0;
  while (true) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (++spins > kSpinLimit) {
      printf("b200asr chain_pair: cluster mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

template <int NT>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// HALVES == 2: two warps share each TMEM lane quadrant and split the row's columns (16-column units [u0, u1)); the LayerNorm
// partial sums are exchanged through `statbuf` ([ROWS][2][2] floats).  With column-split statistics the variance is the
// plain E[x^2] - mean^2 (no common shift is available); fp32 is ample for |x| <~ 1e3 over <= 256 columns.
// NT = threads taking part (named barrier 1).  xchg (optional): a second partial accumulator tile, same layout as `stile`,
// added to the TMEM accumulator before the bias (gemm_chain_pair.cu: the peer CTA's half of the hidden dimension).
template <int EPI, int BLOCK_N, int ROWS, int HALVES = 1, int NT = 128 * HALVES>
__device__ __forceinline__ void epilogue_ln_tma(const TcParams& p, uint32_t taddr, uint8_t* stile, const CUtensorMap* map_c,
                                                const CUtensorMap* map_c2, int row0, int trow, bool issuer, int half = 0,
                                                float* statbuf = nullptr, uint8_t* stile2 = nullptr, uint8_t* xchg = nullptr) {
  constexpr bool has_resid = (EPI == EPI_RESID_LN || EPI == EPI_RESID_LN2);
  constexpr int NSLAB = (BLOCK_N + 31) / 32;
  constexpr int NUNIT = BLOCK_N / 16;                    // 16-column units
  constexpr int SPLIT = (HALVES == 2) ? (NUNIT + 1) / 2 : NUNIT;
  const int u0 = (HALVES == 2 && half == 1) ? SPLIT : 0;
  const int u1 = (HALVES == 2 && half == 0) ? SPLIT : NUNIT;
  const bool active = trow >= 0;
  auto chunk_ptr_in = [&](uint8_t* base, int u, int q) -> float4* {   // 16-byte chunk q (0..3) of unit u in this thread's row
    const int s = u >> 1, qq = ((u & 1) << 2) | q;
    return reinterpret_cast<float4*>(base + (size_t)s * ROWS * 128 + (size_t)(active ? trow : 0) * 128 + (((qq ^ (trow & 7)) & 7) << 4));
  };
  auto chunk_ptr = [&](int u, int q) -> float4* { return chunk_ptr_in(stile, u, q); };
  uint8_t* out2 = stile2 ? stile2 : stile;              // where the second output is staged
  // TMA-store the staged tile.  wait_read: block until the store has finished READING shared memory (tile reusable).
  auto store_tile = [&](const CUtensorMap* map, uint8_t* base, bool wait_read) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    epi_bar_sync<NT>();
    if (issuer) {
#pragma unroll
      for (int s = 0; s < NSLAB; ++s) tma_store_2d(map, base + (size_t)s * ROWS * 128, 32 * s, row0);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (wait_read) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    if (wait_read) epi_bar_sync<NT>();   // the tile(s) may be overwritten again
  };
  // combine per-half partial sums (sum, sum of squares) of this row
  auto combine = [&](float& a, float& b, int slot) {
    if (HALVES == 1) return;
    float* mine = statbuf + ((size_t)(active ? trow : 0) * 2 + half) * 4 + slot * 2;
    if (active) { mine[0] = a; mine[1] = b; }
    epi_bar_sync<NT>();
    const float* other = statbuf + ((size_t)(active ? trow : 0) * 2 + (half ^ 1)) * 4 + slot * 2;
    a += other[0];
    b += other[1];
  };
  // x = acc + bias (+ resid from smem) for the 16 columns of unit u
  auto load_x = [&](int u, float* v) {
    uint32_t raw[16];
    tmem_ld16_nowait(taddr + (uint32_t)(16 * u), raw);   // warp-collective
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = *reinterpret_cast<const float4*>(p.bias + 16 * u + 4 * q);   // (may point to shared memory)
      float a0 = __uint_as_float(raw[4 * q + 0]), a1 = __uint_as_float(raw[4 * q + 1]);
      float a2 = __uint_as_float(raw[4 * q + 2]), a3 = __uint_as_float(raw[4 * q + 3]);
      if (xchg != nullptr) {
        // the two partial sums are added FIRST (a commutative step), the bias after: a row's result must not depend on which CTA
        // of the pair finishes it, i.e. on the utterance's position in the batch
        const float4 o = *chunk_ptr_in(xchg, u, q);
        a0 += o.x; a1 += o.y; a2 += o.z; a3 += o.w;
      }
      a0 += b.x; a1 += b.y; a2 += b.z; a3 += b.w;
      if (has_resid) {
        const float4 r = *chunk_ptr(u, q);
        a0 = r.x + p.alpha * a0; a1 = r.y + p.alpha * a1; a2 = r.z + p.alpha * a2; a3 = r.w + p.alpha * a3;
      }
      v[4 * q + 0] = a0; v[4 * q + 1] = a1; v[4 * q + 2] = a2; v[4 * q + 3] = a3;
    }
  };
  auto put_in = [&](uint8_t* base, int u, const float* v) {
    if (!active) return;
#pragma unroll
    for (int q = 0; q < 4; ++q) *chunk_ptr_in(base, u, q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  };
  auto put = [&](int u, const float* v) { put_in(stile, u, v); };
  auto get = [&](int u, float* v) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 t = *chunk_ptr(u, q);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
  };
  auto affine = [&](float* v, float mean, float rstd, const float* g, const float* be, int u) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 gg = *reinterpret_cast<const float4*>(g + 16 * u + 4 * q);
      const float4 bb = *reinterpret_cast<const float4*>(be + 16 * u + 4 * q);
      v[4 * q + 0] = (v[4 * q + 0] - mean) * rstd * gg.x + bb.x; v[4 * q + 1] = (v[4 * q + 1] - mean) * rstd * gg.y + bb.y;
      v[4 * q + 2] = (v[4 * q + 2] - mean) * rstd * gg.z + bb.z; v[4 * q + 3] = (v[4 * q + 3] - mean) * rstd * gg.w + bb.w;
    }
  };
  // C2 (the LayerNorm-ed copy) is only ever read as the A operand of the next module's GEMMs: store it rounded to nearest tf32
  // (the tensor core would otherwise truncate it)
  auto round16 = [&](float* v) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(tf32_rn_bits(v[i]));
  };
  const float invn = 1.0f / (float)BLOCK_N;

  // (A register-resident variant -- every TMEM load of the row segment issued before the first use, x / y kept in registers between the
  // statistics and the normalisation pass -- was measured in round 1 and dropped: 1.877 -> 1.986 ms per step, DESIGN.md "tried and dropped".)
  // the row is swept from TMEM / the staging tile once per pass
  // sweep 1: x (kept in the tile when C holds the un-normalised stream) + statistics
  float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
  for (int u = u0; u < u1; ++u) {
    float v[16];
    load_x(u, v);
    if (HALVES == 1 && u == 0) shift = v[0];              // shifted one-pass variance when one thread sees the whole row
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float d = v[i] - shift;
      s1 += d;
      s2 = fmaf(d, d, s2);
    }
    if (EPI != EPI_RESID_LN2) put(u, v);
  }
  combine(s1, s2, 0);
  const float m1 = s1 * invn;
  const float mean1 = shift + m1;
  const float rstd1 = rsqrtf(fmaxf(s2 * invn - m1 * m1, 0.f) + p.ln_eps);

  if (EPI != EPI_RESID_LN2) {
    store_tile(map_c, stile, stile2 == nullptr);         // C = x   (no wait when C2 is staged elsewhere)
#pragma unroll 1
    for (int u = u0; u < u1; ++u) {
      float v[16];
      get(u, v);
      affine(v, mean1, rstd1, p.ln1_g, p.ln1_b, u);
      round16(v);
      put_in(out2, u, v);
    }
    store_tile(map_c2, out2, true);                      // C2 = LN(x; ln1)
    return;
  }
  float shift2 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll 1
  for (int u = u0; u < u1; ++u) {
    float v[16];
    load_x(u, v);
    affine(v, mean1, rstd1, p.ln1_g, p.ln1_b, u);
    if (HALVES == 1 && u == 0) shift2 = v[0];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float d = v[i] - shift2;
      t1 += d;
      t2 = fmaf(d, d, t2);
    }
    if (p.round_out) round16(v);   // (set only when no second LayerNorm follows: C is then read as a tensor-core operand only)
    put(u, v);
  }
  if (p.ln2_g == nullptr) {
    store_tile(map_c, stile, true);
    return;
  }
  store_tile(map_c, stile, stile2 == nullptr);           // C = y = LN(x; ln1)
  combine(t1, t2, 1);
  const float m2 = t1 * invn;
  const float mean2 = shift2 + m2;
  const float rstd2 = rsqrtf(fmaxf(t2 * invn - m2 * m2, 0.f) + p.ln_eps);
#pragma unroll 1
  for (int u = u0; u < u1; ++u) {
    float v[16];
    get(u, v);
    affine(v, mean2, rstd2, p.ln2_g, p.ln2_b, u);
    round16(v);
    put_in(out2, u, v);
  }
  store_tile(map_c2, out2, true);                        // C2 = LN(y; ln2)
}

// ------------------------------------------------------------------------------------------------ host side
using EncodeTiledFn = PFN_cuTensorMapEncodeTiled_v12000;

inline int encode_map(TcContext& ctx, CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
               const cuuint32_t* box, const cuuint32_t* estr, bool f16 = false) {
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx.encode_tiled);
  CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_errbuf, sizeof(g_errbuf), "cuTensorMapEncodeTiled failed (%d) rank=%d dims=%llu,%llu box=%u,%u", (int)r, rank,
             (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return 1;
  }
  return 0;
}

}  // namespace tc
}  // namespace b200asr
