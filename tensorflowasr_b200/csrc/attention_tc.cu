// Multi-head self-attention on the 5th-gen tensor cores (tcgen05, tf32 inputs, fp32 accumulate in TMEM).
//
// Reference semantics: MultiHeadAttention.call (asr/models/layers/multihead_attention.py:151-188): softmax(Q K^T) V per head,
// no mask, no positional term, 1/sqrt(d) already folded into Wq.  One CTA = (batch, head, 128 queries):
//   * Q [128 x d], K [256 x d] and V^T [d x 256] tiles are written by the CTA's threads straight into the canonical
//     K-major SWIZZLE_128B shared-memory layout (rounded to nearest tf32, zero padded to 64 columns / masked rows), so no
//     padded copy of the QKV activations is ever needed in HBM;
//   * S = Q K^T: 8 x tcgen05.mma (M=128, N=256, K=8) into TMEM columns [0,256);
//   * softmax: thread == query row; the row is swept from TMEM twice (max, then exp / sum); P is written back IN PLACE
//     over S with tcgen05.st (tf32-truncated, and the row sum is taken over the truncated values so the truncation
//     cancels in the normalisation);
//   * O_blk = P V: 32 x tcgen05.mma with the A operand read from TMEM (TS form), B = V^T from shared memory, into TMEM
//     columns [256,320); the running output is kept in registers with the usual online-softmax rescale, so longer
//     sequences simply loop over 256-key blocks.
// An optional band (chunk_conformer_blocks.py:158-176) restricts the visible keys per query.
#include "kernels.cuh"

#include <cstdlib>

namespace b200asr {

namespace {

constexpr int kQT = 128;     // queries per CTA
constexpr int kKT = 256;     // keys per block
constexpr int kDP = 64;      // head dim padded to two 32-float swizzle slabs
constexpr int kThreadsA = 384;     // warps 0-7: softmax (two per TMEM lane quadrant, 128 key columns each), warp 8: MMA issue +
                                   // TMEM, all 12 warps: tile staging
constexpr int kMmaWarp = 8;
constexpr unsigned kSpin = 1u << 28;

__device__ __forceinline__ uint32_t smem_u32a(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// round-to-nearest (ties away) to tf32 on the integer ALU
__device__ __forceinline__ float to_tf32_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void mbar_init_a(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32a(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait_a(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32a(bar);
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
    if (++spins > kSpin) {
      printf("b200asr attention_tc: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void commit_a(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32a(bar)) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}

// cp.async (LDGSTS): the copy engine moves the bytes straight into shared memory, dozens in flight per thread, nothing blocks on the
// L2 round trip; src_bytes = 0 zero-fills the destination (rows beyond T, padding columns).  Used when the producer of QKV already
// rounded it to tf32 (the engine's schedule), so that staging does not have to.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// byte offset of element (row, k) inside a K-major SWIZZLE_128B tile made of 32-float slabs of `rows` rows each
__device__ __forceinline__ uint32_t sw128_off(int row, int k, int rows) {
  const int slab = k >> 5, kk = k & 31;
  return (uint32_t)(slab * rows * 128 + row * 128 + ((((kk >> 2) ^ (row & 7)) << 4) | ((kk & 3) << 2)));
}

__global__ void __launch_bounds__(kThreadsA, 1) attention_tc_kernel(const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment by POINTER arithmetic on the shared array (an integer round trip would strip the address space and turn every
  // access through a derived pointer into a generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  uint8_t* Qs = smem;                                  // 2 slabs x 128 rows x 128 B = 32 KB
  uint8_t* Ks = Qs + 2 * kQT * 128;                    // 2 slabs x 256 rows x 128 B = 64 KB
  uint8_t* Vt = Ks + 2 * kKT * 128;                    // 8 slabs x  64 rows x 128 B = 64 KB   (V transposed: rows = head dim)
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(Vt + 8 * kDP * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
  float* red = reinterpret_cast<float*>(tmem_slot + 2);      // [2 kinds][128 rows][2 halves] partial max / sum exchange

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  int ev = 0;
  auto stamp = [&]() {
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0 && ev < 64) p.dbg[ev++] = clock64();
  };
  stamp();
  const int dh = p.dh, ld = 3 * p.H * dh;
  const float* base = p.qkv + (size_t)b * p.T * ld;
  const float* qbase = base + h * dh;
  const float* kbase = base + p.H * dh + h * dh;
  const float* vbase = base + 2 * p.H * dh + h * dh;

  if (tid == 0) {
    mbar_init_a(mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32a(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  pdl_trigger();   // TMEM is allocated: the next kernel's CTAs may start their prologue
  pdl_wait();      // QKV (written by the previous kernel) is complete and visible
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;            // columns [0, 256)
  const uint32_t tmem_O = tmem_base + kKT;      // columns [256, 320)
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t phase = 0;
  const bool single_block = (p.T <= kKT);       // K / V^T staged once and reused by every query tile of this CTA
  bool kv_loaded = false;

  // staging helpers: kSU independent 16-byte loads in flight per thread before the first shared-memory store (measured: 8 in
  // flight made the kernel slower, 20.7 vs 18.6 us; pre-zeroed padding with fewer stores did not help either -- the loops are
  // bound by the latency of the strided global loads)
  constexpr int kSU = 4;
  // (t0, nt): the staging threads are t0 .. t0 + nt - 1 of the CTA (all of it, or the four warps that idle during the softmax)
  auto stage_rows = [&](uint8_t* dst, const float* src, int row0, int rows, int t0 = 0, int nt = kThreadsA) {   // row-major tile -> K-major SW128 (Q, K)
    for (int i0 = tid - t0; i0 < rows * (kDP / 4); i0 += kSU * nt) {
      float4 v[kSU];
#pragma unroll
      for (int u = 0; u < kSU; ++u) {
        const int i = i0 + u * nt;
        const int r = i >> 4, c4 = i & 15;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < rows * (kDP / 4) && row0 + r < p.T && 4 * c4 < dh)
          v[u] = *reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * ld + 4 * c4);
      }
#pragma unroll
      for (int u = 0; u < kSU; ++u) {
        const int i = i0 + u * nt;
        if (i < rows * (kDP / 4)) {
          const int r = i >> 4, c4 = i & 15;
          float4 w = v[u];
          w.x = to_tf32_rn(w.x); w.y = to_tf32_rn(w.y); w.z = to_tf32_rn(w.z); w.w = to_tf32_rn(w.w);
          *reinterpret_cast<float4*>(dst + sw128_off(r, 4 * c4, rows)) = w;
        }
      }
    }
  };
  auto stage_vt = [&](int k0) {                                                 // V [key][d] -> V^T [d][key] K-major SW128
    for (int i0 = tid; i0 < kKT * (kDP / 4); i0 += kSU * kThreadsA) {
      float4 v[kSU];
#pragma unroll
      for (int u = 0; u < kSU; ++u) {
        const int i = i0 + u * kThreadsA;
        const int key = i & (kKT - 1), c4 = i >> 8;   // key fastest: conflict-free transposed stores
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < kKT * (kDP / 4) && k0 + key < p.T && 4 * c4 < dh)
          v[u] = *reinterpret_cast<const float4*>(vbase + (size_t)(k0 + key) * ld + 4 * c4);
      }
#pragma unroll
      for (int u = 0; u < kSU; ++u) {
        const int i = i0 + u * kThreadsA;
        if (i < kKT * (kDP / 4)) {
          const int key = i & (kKT - 1), c4 = i >> 8;
          *reinterpret_cast<float*>(Vt + sw128_off(4 * c4 + 0, key, kDP)) = to_tf32_rn(v[u].x);
          *reinterpret_cast<float*>(Vt + sw128_off(4 * c4 + 1, key, kDP)) = to_tf32_rn(v[u].y);
          *reinterpret_cast<float*>(Vt + sw128_off(4 * c4 + 2, key, kDP)) = to_tf32_rn(v[u].z);
          *reinterpret_cast<float*>(Vt + sw128_off(4 * c4 + 3, key, kDP)) = to_tf32_rn(v[u].w);
        }
      }
    }
  };

  // asynchronous variants (p.async_stage): same destinations, no rounding (the data is tf32 already)
  const int nch_real = dh >> 2, nch_total = ((dh + 7) >> 3) << 1;   // 16-byte chunks per row holding data / read by the issued k-steps
  auto stage_rows_async = [&](uint8_t* dst, const float* src, int row0, int rows, int t0 = 0, int nt = kThreadsA) {
    const uint32_t d0 = smem_u32a(dst);
    for (int i = tid - t0; i < rows * (kDP / 4); i += nt) {
      const int r = i >> 4, c4 = i & 15;
      if (c4 >= nch_total) continue;
      const bool real = (row0 + r < p.T) && (c4 < nch_real);
      cp_async16(d0 + sw128_off(r, 4 * c4, rows), real ? (const void*)(src + (size_t)(row0 + r) * ld + 4 * c4) : (const void*)src, real ? 16 : 0);
    }
  };
  auto stage_vt_async = [&](int k0) {
    const uint32_t d0 = smem_u32a(Vt);
    for (int i = tid; i < kKT * dh; i += kThreadsA) {
      const int key = i & (kKT - 1), d = i >> 8;      // key fastest: conflict-free transposed writes
      const bool real = (k0 + key < p.T);
      cp_async4(d0 + sw128_off(d, key, kDP), real ? (const void*)(vbase + (size_t)(k0 + key) * ld + d) : (const void*)vbase, real ? 4 : 0);
    }
  };
  const bool async_stage = p.async_stage != 0;

  bool q_staged = false;                          // the next query tile was staged under the previous tile's softmax
  for (int q0 = blockIdx.x * kQT; q0 < p.T; q0 += gridDim.x * kQT) {
  stamp();
  if (!q_staged) {                                 // rows beyond T and columns beyond dh are zero
    if (async_stage) stage_rows_async(Qs, qbase, q0, kQT);
    else stage_rows(Qs, qbase, q0, kQT);
  }
  q_staged = false;
  stamp();

  // per-row state: warps 0..7, row = (warp & 3) * 32 + lane; `half` selects this thread's 128 key columns of a block and
  // its 32 output columns
  const int half = (warp >> 2) & 1;
  const int row = (warp & 3) * 32 + lane;
  const int qi = q0 + row;
  float m_run = -INFINITY, l_run = 0.f;
  float o[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) o[d] = 0.f;
  int lo = 0, hi = p.T - 1;
  const bool band = p.win_front >= 0;
  if (band) {
    lo = max(min(max(qi - p.win_front, 0), p.T - p.win_back), 0);
    hi = min(max(min(qi + p.win_back, p.T), p.win_back), p.T - 1);
  }

  for (int k0 = 0; k0 < p.T; k0 += kKT) {
    // ---- stage K [256 x 64] and V^T [64 x 256] for this key block
    if (!(single_block && kv_loaded)) {
      if (async_stage) {
        stage_rows_async(Ks, kbase, k0, kKT);
        stage_vt_async(k0);
      } else {
        stage_rows(Ks, kbase, k0, kKT);
        stage_vt(k0);
      }
      kv_loaded = true;
    }
    if (async_stage) cp_async_wait_all();          // (also covers the Q tile issued above)
    stamp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
    fence_before();
    __syncthreads();
    // ---- S = Q K^T
    if (warp == kMmaWarp) {
      fence_after();
      if (lane == 0) {
        const uint32_t qa = smem_u32a(Qs), ka = smem_u32a(Ks);
        const int ksteps = (dh + 7) / 8;               // only k-steps that hold real data (rest is zero padding)
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t da = smem_desc_sw128(qa + (ks >> 2) * (kQT * 128) + (ks & 3) * 32);
          const uint64_t db = smem_desc_sw128(ka + (ks >> 2) * (kKT * 128) + (ks & 3) * 32);
          mma_ss(tmem_S, da, db, idesc_tf32(kQT, kKT), ks > 0 ? 1u : 0u);
        }
        commit_a(mma_bar);
      }
      __syncwarp();
    }
    mbar_wait_a(mma_bar, phase);
    stamp();
    phase ^= 1;
    fence_after();
    // ---- the Q tile is dead once the last S = Q K^T of this query tile has completed: the four warps that do not take part in
    // the softmax stage the NEXT query tile over it now (visible to the tensor core through the proxy fence + the barriers below)
    const int q_next = q0 + gridDim.x * kQT;
    if (k0 + kKT >= p.T && q_next < p.T) {
      if (warp >= 8) {
        if (async_stage) {
          stage_rows_async(Qs, qbase, q_next, kQT, 256, kThreadsA - 256);
          cp_async_wait_all();
        } else {
          stage_rows(Qs, qbase, q_next, kQT, 256, kThreadsA - 256);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      q_staged = true;
    }
    // ---- softmax over this block's keys (8 warps: each thread sweeps its 128 columns twice), P written back in place
    float corr = 1.f;
    const int nk = min(kKT, p.T - k0);
    if (warp < 8) {
      const int cbeg = half * (kKT / 2), cend = cbeg + kKT / 2;
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = cbeg; c < cend; c += 32) {
        if (c >= nk) break;                            // warp-uniform: whole chunk beyond the sequence
        uint32_t r[32];
        tmem_ld32(tmem_S + lane_addr + (uint32_t)c, r);
        if (!band && c + 32 <= nk) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int key = k0 + c + j;
            if (c + j < nk && key >= lo && key <= hi) mx = fmaxf(mx, __uint_as_float(r[j]));
          }
        }
      }
      red[row * 2 + half] = mx;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      mx = fmaxf(red[row * 2], red[row * 2 + 1]);
      const float m_new = fmaxf(m_run, mx);
      const bool any = (m_new != -INFINITY);
      corr = (m_run == -INFINITY) ? 0.f : __expf(m_run - m_new);
      if (!any) corr = 1.f;
      const float mscaled = m_new * 1.4426950408889634f;
      float sum = 0.f;
#pragma unroll 1
      for (int c = cbeg; c < cend; c += 32) {
        if (c >= ((nk + 7) & ~7)) break;               // P columns at or beyond ceil8(nk) are never read by the P.V MMAs
        uint32_t r[32];
        tmem_ld32(tmem_S + lane_addr + (uint32_t)c, r);
        if (!band && c + 32 <= nk) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(__uint_as_float(r[j]), 1.4426950408889634f, -mscaled)));
            const uint32_t pt = __float_as_uint(e) & 0xFFFFE000u;   // what the tf32 datapath will see
            sum += __uint_as_float(pt);
            r[j] = pt;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int key = k0 + c + j;
            float pv = 0.f;
            if (any && c + j < nk && key >= lo && key <= hi) pv = __expf(__uint_as_float(r[j]) - m_new);
            const uint32_t pt = __float_as_uint(pv) & 0xFFFFE000u;
            sum += __uint_as_float(pt);
            r[j] = pt;
          }
        }
        tmem_st32(tmem_S + lane_addr + (uint32_t)c, r);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      red[256 + row * 2 + half] = sum;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      l_run = l_run * corr + red[256 + row * 2] + red[256 + row * 2 + 1];
      m_run = m_new;
    }
    stamp();
    fence_before();
    __syncthreads();
    // ---- O_blk = P V   (A from TMEM)
    if (warp == kMmaWarp) {
      fence_after();
      if (lane == 0) {
        const uint32_t va = smem_u32a(Vt);
        const int ksteps = (nk + 7) / 8;               // keys beyond nk have P == 0 and V^T == 0
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t db = smem_desc_sw128(va + (ks >> 2) * (kDP * 128) + (ks & 3) * 32);
          mma_ts(tmem_O, tmem_S + (uint32_t)(8 * ks), db, idesc_tf32(kQT, kDP), ks > 0 ? 1u : 0u);
        }
        commit_a(mma_bar);
      }
      __syncwarp();
    }
    mbar_wait_a(mma_bar, phase);
    stamp();
    phase ^= 1;
    fence_after();
    if (warp < 8) {
      uint32_t r[32];
      tmem_ld32(tmem_O + lane_addr + (uint32_t)(32 * half), r);
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = o[j] * corr + __uint_as_float(r[j]);
    }
    fence_before();
    __syncthreads();   // S/P, O_blk and the Q/K/V tiles may be overwritten next
    fence_after();
  }

  if (warp < 8 && qi < p.T) {
    const float inv = 1.0f / l_run;
    float* orow = p.out + ((size_t)b * p.T + qi) * (p.H * dh) + h * dh + 32 * half;
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      if (32 * half + 4 * c4 < dh)
      {
        float4 w4 = make_float4(o[4 * c4] * inv, o[4 * c4 + 1] * inv, o[4 * c4 + 2] * inv, o[4 * c4 + 3] * inv);
        if (p.round_tf32) { w4.x = to_tf32_rn(w4.x); w4.y = to_tf32_rn(w4.y); w4.z = to_tf32_rn(w4.z); w4.w = to_tf32_rn(w4.w); }
        *reinterpret_cast<float4*>(orow + 4 * c4) = w4;
      }
    }
  }
  stamp();
  }  // query tiles
  fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace

bool attention_tc_supported(const AttnParams& p) {
  return p.dh % 4 == 0 && p.dh <= kDP && p.dh >= 4 && ((p.H * p.dh) % 4) == 0 && p.T > 0;
}

int launch_attention_tc(const AttnParams& p, cudaStream_t stream) {
  if (p.B == 0 || p.T == 0) return 0;
  const size_t smem = (size_t)2 * kQT * 128 + 2 * kKT * 128 + 8 * kDP * 128 + 1024 + 64 + 2 * 128 * 2 * 4;
  static PerDeviceSmem configured;
  if (configured.need(smem)) B200_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // one CTA per (batch, head) when the whole sequence is a single key block (K / V^T staged once for all query tiles)
  const int qtiles = ceil_div(p.T, kQT);
  dim3 grid(p.T <= kKT ? 1 : qtiles, p.H, p.B);
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    const char* e = getenv("B200ASR_ATTN_DBG");
    dbg_on = (e && e[0] == '1') ? 1 : 0;
    if (dbg_on) cudaMalloc(&dbg, sizeof(long long) * 64);
  }
  AttnParams pp = p;
  if (dbg_on) {
    pp.dbg = dbg;
    cudaMemset(dbg, 0, sizeof(long long) * 64);
  }
  B200_CUDA_OK(launch_k(attention_tc_kernel, grid, dim3(kThreadsA), smem, stream, pp));
  B200_CUDA_OK(cudaGetLastError());
  if (dbg_on) {
    long long hb[64];
    cudaDeviceSynchronize();
    cudaMemcpy(hb, dbg, sizeof(hb), cudaMemcpyDeviceToHost);
    fprintf(stderr, "attn-dbg:");
    for (int i = 0; i < 64 && hb[i]; ++i) fprintf(stderr, " %lld", hb[i] - hb[0]);
    fprintf(stderr, "\n");
  }
  return 0;
}

}  // namespace b200asr
