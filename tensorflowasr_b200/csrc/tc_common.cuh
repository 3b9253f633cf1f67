// Shared device/host helpers of the tcgen05 kernels (gemm_tc.cu, gemm_chain.cu): PTX wrappers for mbarrier / TMA /
// tcgen05 (mma, ld, st, commit, fences), UMMA descriptors, fused epilogues, tensor-map encoding.
#pragma once
#include "gemm_tc.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

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
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float swish_fast(float x) { return x * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }

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
__device__ __forceinline__ uint32_t tf32_rn_bits(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
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
template <int EPI>
__device__ __forceinline__ void epilogue_store16(const TcParams& p, float* v, size_t row_off, int n) {
  // v: 16 consecutive accumulator columns starting at output column n of the row at C + row_off
  if (EPI != EPI_NONE && p.bias != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (n + 4 * q < p.N) {
        const float4 b = *reinterpret_cast<const float4*>(p.bias + n + 4 * q);
        v[4 * q + 0] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
      }
    }
  }
  if (EPI == EPI_GLU) {
    float* dst = p.C + row_off + (n >> 1);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (n + 8 * q < p.N) {
        float4 o;
        o.x = v[8 * q + 0] * sigmoid_fast(v[8 * q + 1]);
        o.y = v[8 * q + 2] * sigmoid_fast(v[8 * q + 3]);
        o.z = v[8 * q + 4] * sigmoid_fast(v[8 * q + 5]);
        o.w = v[8 * q + 6] * sigmoid_fast(v[8 * q + 7]);
        *reinterpret_cast<float4*>(dst + 4 * q) = o;
      }
    }
    return;
  }
  float* dst = p.C + row_off + n;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (n + 4 * q < p.N) {
      float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      if (EPI == EPI_BIAS_RELU) {
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
      } else if (EPI == EPI_BIAS_SWISH) {
        o.x = swish_fast(o.x); o.y = swish_fast(o.y); o.z = swish_fast(o.z); o.w = swish_fast(o.w);
      } else if (EPI == EPI_RESID) {
        const float4 r = *reinterpret_cast<const float4*>(p.resid + row_off + n + 4 * q);
        o.x = r.x + p.alpha * o.x; o.y = r.y + p.alpha * o.y; o.z = r.z + p.alpha * o.z; o.w = r.w + p.alpha * o.w;
      }
      *reinterpret_cast<float4*>(dst + 4 * q) = o;
    }
  }
}


// ------------------------------------------------------------------------------------------------ fused LayerNorm epilogues
// thread == output row and BLOCK_N == N, so a row's statistics never leave the thread: the accumulator row is swept from
// TMEM two (three) times -- statistics (shifted one-pass variance), then normalise -- instead of being parked in registers.
template <int EPI, int BLOCK_N>
__device__ __forceinline__ void epilogue_ln(const TcParams& p, uint32_t taddr, bool row_ok, size_t row_off) {
  constexpr bool has_resid = (EPI == EPI_RESID_LN || EPI == EPI_RESID_LN2);
  constexpr int G = (BLOCK_N % 48 == 0) ? 48 : ((BLOCK_N % 32 == 0) ? 32 : 16);   // columns per batch of loads
  // x[c0 .. c0+G) of this thread's row: accumulator (TMEM) + bias (+ residual).  All G/16 tcgen05.ld and all residual
  // loads are issued before the first use so their latencies overlap (one exposed L2 round trip per batch, not per chunk).
  auto load_x = [&](int c0, float* v) {
    uint32_t raw[G];
#pragma unroll
    for (int j = 0; j < G / 16; ++j) tmem_ld16_nowait(taddr + (uint32_t)(c0 + 16 * j), raw + 16 * j);   // warp-collective
    float4 rr[G / 4];
    if (has_resid && row_ok) {
#pragma unroll
      for (int q = 0; q < G / 4; ++q) rr[q] = *reinterpret_cast<const float4*>(p.resid + row_off + c0 + 4 * q);
    }
    tmem_ld_wait();
    if (!row_ok) return;
#pragma unroll
    for (int q = 0; q < G / 4; ++q) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + 4 * q));
      float a0 = __uint_as_float(raw[4 * q + 0]) + b.x, a1 = __uint_as_float(raw[4 * q + 1]) + b.y;
      float a2 = __uint_as_float(raw[4 * q + 2]) + b.z, a3 = __uint_as_float(raw[4 * q + 3]) + b.w;
      if (has_resid) {
        a0 = rr[q].x + p.alpha * a0; a1 = rr[q].y + p.alpha * a1; a2 = rr[q].z + p.alpha * a2; a3 = rr[q].w + p.alpha * a3;
      }
      v[4 * q + 0] = a0; v[4 * q + 1] = a1; v[4 * q + 2] = a2; v[4 * q + 3] = a3;
    }
  };
  // re-read this thread's own row from C (written earlier by this same thread): needed because C may alias resid
  auto load_c = [&](int c0, float* v) {
#pragma unroll
    for (int q = 0; q < G / 4; ++q) {
      const float4 r = *reinterpret_cast<const float4*>(p.C + row_off + c0 + 4 * q);
      v[4 * q + 0] = r.x; v[4 * q + 1] = r.y; v[4 * q + 2] = r.z; v[4 * q + 3] = r.w;
    }
  };
  auto store_g = [&](float* dst, const float* v) {
#pragma unroll
    for (int q = 0; q < G / 4; ++q) *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  };
  auto affine_g = [&](float* v, float mean, float rstd, const float* g, const float* be, int c0) {
#pragma unroll
    for (int q = 0; q < G / 4; ++q) {
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c0 + 4 * q));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(be + c0 + 4 * q));
      v[4 * q + 0] = (v[4 * q + 0] - mean) * rstd * gg.x + bb.x; v[4 * q + 1] = (v[4 * q + 1] - mean) * rstd * gg.y + bb.y;
      v[4 * q + 2] = (v[4 * q + 2] - mean) * rstd * gg.z + bb.z; v[4 * q + 3] = (v[4 * q + 3] - mean) * rstd * gg.w + bb.w;
    }
  };
  const float invn = 1.0f / (float)BLOCK_N;
  // sweep 1: x (stored for *_LN modes where C holds the un-normalised stream) + statistics (shifted one-pass variance)
  float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
  for (int c = 0; c < BLOCK_N; c += G) {
    float v[G];
    load_x(c, v);
    if (row_ok) {
      if (c == 0) shift = v[0];
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const float d = v[i] - shift;
        s1 += d;
        s2 = fmaf(d, d, s2);
      }
      if (EPI != EPI_RESID_LN2) store_g(p.C + row_off + c, v);
    }
  }
  const float m1 = s1 * invn;
  const float mean1 = shift + m1;
  const float rstd1 = 1.0f / sqrtf(fmaxf(s2 * invn - m1 * m1, 0.f) + p.ln_eps);
  if (EPI != EPI_RESID_LN2) {
    // sweep 2: LN(x; ln1) -> C2   (x read back from C: the residual operand may have been overwritten in place)
    if (!row_ok) return;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += G) {
      float v[G];
      load_c(c, v);
      affine_g(v, mean1, rstd1, p.ln1_g, p.ln1_b, c);
      store_g(p.C2 + row_off + c, v);
    }
    return;
  }
  // EPI_RESID_LN2: sweep 2: y = LN(x; ln1) -> C, statistics of y; sweep 3: LN(y; ln2) -> C2
  float shift2 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll 1
  for (int c = 0; c < BLOCK_N; c += G) {
    float v[G];
    load_x(c, v);
    if (row_ok) {
      affine_g(v, mean1, rstd1, p.ln1_g, p.ln1_b, c);
      if (c == 0) shift2 = v[0];
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const float d = v[i] - shift2;
        t1 += d;
        t2 = fmaf(d, d, t2);
      }
      store_g(p.C + row_off + c, v);
    }
  }
  if (p.ln2_g == nullptr || !row_ok) return;
  const float m2 = t1 * invn;
  const float mean2 = shift2 + m2;
  const float rstd2 = 1.0f / sqrtf(fmaxf(t2 * invn - m2 * m2, 0.f) + p.ln_eps);
#pragma unroll 1
  for (int c = 0; c < BLOCK_N; c += G) {
    float v[G];
    load_c(c, v);                     // y, as stored in sweep 2
    affine_g(v, mean2, rstd2, p.ln2_g, p.ln2_b, c);
    store_g(p.C2 + row_off + c, v);
  }
}


// ------------------------------------------------------------------------------------------------ host side
using EncodeTiledFn = PFN_cuTensorMapEncodeTiled_v12000;

inline int encode_map(TcContext& ctx, CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
               const cuuint32_t* box, const cuuint32_t* estr) {
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx.encode_tiled);
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
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
