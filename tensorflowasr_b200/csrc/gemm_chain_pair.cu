// Chained FFN / conv-tail GEMMs (see gemm_chain.cu) with the HIDDEN dimension split across a cluster of two CTAs.
//
// Why: a tcgen05.mma kind::tf32 instruction (M=128, K=8) costs ~132 cycles whatever N <= 144 is (measured, scripts/ubench_mma.cu;
// the A operand alone is 128 rows x 32 B), so one 128-row tile of the FFN (72 + 72 instructions) holds an SM for ~19 k cycles while
// the 8000-row activations of the benchmark batch only make 63 tiles for 148 SMs.  Here both CTAs of a cluster work on the SAME 128
// rows: CTA r streams the W1 rows / W2 columns of hidden chunks [r*n/2, (r+1)*n/2) only, i.e. half of the MMA instructions, half
// of the weight bytes through its L2 port and half of the activation work.  The two partial [128 x N2] accumulators are
// combined through distributed shared memory: each CTA stages the 64 rows it does not own in its (idle) weight ring, swizzled like
// the staging tiles, and one warp sends them as five 8 KB cp.async.bulk.shared::cluster copies into the peer's dead X slabs
// (per-lane st.shared::cluster of row fragments measured 4.7 k cycles for the same bytes); each CTA then finishes residual +
// LayerNorm(s) + TMA stores for its own 64 rows.
//
//   warp 0      TMA producer (first weight slabs before griddepcontrol.wait; X; weight ring; the 64-row residual tile)
//   warp 1      MMA issuer + TMEM owner (acc1 double buffer [0,2CH), partial acc2 [2CH, 2CH+N2))
//   warps 2-9   activation of every local chunk; then quadrants of the peer's rows ship, quadrants of the own rows finish
// DIRECT variant (attention out-projection): no hidden layer -- acc2 = X . W2^T with the K dimension (144 = 18 k-steps) split 9 / 9
// across the pair, then the same exchange and residual + LayerNorm epilogue (64 rows per CTA on 8 warps instead of the plain
// GEMM kernel's 64-row tiles whose epilogue runs on 4 half-empty warps).
// Cross-CTA protocol (mbarriers, one tile per CTA, so every parity is 0):
//   xchg_ready  (in the WRITER's smem, arrived remotely by the destination's producer as soon as its first-GEMM MMAs have
//               retired): "my X slabs are dead, you may write into them"
//   xchg_full   (in the DESTINATION's smem: expect_tx by the destination, complete_tx by the writer's bulk copy): "your peer's
//               partial has landed"
#include "tc_common.cuh"

#include <cstdlib>

namespace b200asr {

using namespace tc;

namespace {

constexpr int kPairThreads = 320;
constexpr int CH = 144, N2 = 144, STAGES = 6, BM = 128, HR = 64;   // HR: rows finished per CTA
constexpr uint32_t kXSlab = BM * 128;                              // 16 KB per 32-column slab of X
constexpr uint32_t kHSlab = HR * 128;                              // 8 KB per slab of a 64-row tile
constexpr uint32_t kRing = CH * 128;                               // one weight slab
constexpr int KB2 = (CH + 31) / 32;
constexpr int KSTEPS2 = CH / 8;
constexpr int kSlabs = (N2 + 31) / 32;                             // 5

struct PairParams {
  TcParams ep;
  const float* bias1;
  int nl;            // hidden chunks per CTA
  int kb1;           // ceil(K1 / 32) slabs of X (== kSlabs)
  int ks1;           // ceil(K1 / 8) k-steps of the first GEMM
  int num_m_tiles;
  long long* dbg;    // optional timeline of cluster 0 / CTA 0 (B200ASR_PAIR_DBG=1): [role][event] clock64 stamps
};

template <int EPI, bool DIRECT>
__global__ void __launch_bounds__(kPairThreads, 1)
gemm_chain_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                       const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_r,
                       const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_c2, const PairParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment by POINTER arithmetic on the shared array (an integer round trip would strip the address space and turn every
  // access through a derived pointer into a generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  uint8_t* xs = smem;                                    // X: kSlabs x 16 KB; later [0,40K) residual/output tile, [40K,80K) peer partial
  uint8_t* xchg = xs + kSlabs * kHSlab;                  // 64-row tile written by the peer CTA
  uint8_t* ring = xs + (size_t)kSlabs * kXSlab;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)STAGES * kRing);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* x_full = empty_bar + STAGES;
  uint64_t* x_empty = x_full + 1;
  uint64_t* acc1_full = x_empty + 1;     // [2]
  uint64_t* act_done = acc1_full + 2;    // [2]
  uint64_t* acc2_full = act_done + 2;
  uint64_t* r_full = acc2_full + 1;
  uint64_t* xchg_ready = r_full + 1;
  uint64_t* xchg_full = xchg_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xchg_full + 1);
  float* statbuf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + (((2 * STAGES + 10) * 8 + 4 + 15) / 16) * 16);   // [64][2][4] (16-byte aligned: bars sit on a 1024-byte boundary)
  float* pcache = statbuf + HR * 2 * 4;                  // bias1 (local chunks) | bias2 | ln1_g | ln1_b | ln2_g | ln2_b

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t peer = rank ^ 1u;
  const int tile = blockIdx.x >> 1;
  const int nl = p.nl;
  const int j0 = (int)rank * nl;                          // first hidden chunk of this CTA
  int ev = 0;
  auto stamp = [&](int role) {
    if (p.dbg && blockIdx.x == 0 && lane == 0 && ev < 32) p.dbg[role * 32 + ev++] = clock64();
  };
  if (warp == 0) stamp(0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_r) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c2) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc1_full[a], 1);
      mbar_init(&act_done[a], 8);
    }
    mbar_init(acc2_full, 1);
    mbar_init(r_full, 1);
    mbar_init(xchg_ready, 1);
    mbar_init(xchg_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers exist before anything is signalled across the pair
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc2 = tmem_base + 2 * CH;
  pdl_trigger();               // TMEM is allocated: the next kernel's CTAs may start their prologue

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto ring_load = [&](const CUtensorMap* map, int c0, int c1, uint32_t bytes) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], bytes);
        tma_load_2d(map, &full_bar[stage], ring + (size_t)stage * kRing, c0, c1);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      };
      auto g1 = [&](int j, int kb0) { for (int kb = kb0; kb < p.kb1; ++kb) ring_load(&map_w1, kb * 32, j * CH, CH * 128); };
      auto g2 = [&](int j) { for (int kb = 0; kb < KB2; ++kb) ring_load(&map_w2, j * CH + kb * 32, 0, N2 * 128); };
      // weights are constants: the first slabs are requested before griddepcontrol.wait (under the previous kernel's tail)
      const int npre = p.kb1 < STAGES ? p.kb1 : STAGES;
      if (DIRECT) {
        // this CTA's half of K: k-steps [9 rank, 9 rank + 9) live in weight slabs {0,1,2} (rank 0) / {2,3,4} (rank 1)
        for (int i = 0; i < 3; ++i) ring_load(&map_w2, (2 * (int)rank + i) * 32, 0, N2 * 128);
      } else {
        for (int kb = 0; kb < npre; ++kb) ring_load(&map_w1, kb * 32, j0 * CH, CH * 128);
      }
      stamp(0);
      pdl_wait();
      stamp(0);
      mbar_expect_tx(x_full, (uint32_t)p.kb1 * kXSlab);
      for (int kb = 0; kb < p.kb1; ++kb) tma_load_2d(&map_x, x_full, xs + (size_t)kb * kXSlab, kb * 32, tile * BM);
      if (!DIRECT) {
        g1(j0, npre);
        for (int jj = 1; jj < nl; ++jj) { g1(j0 + jj, 0); g2(j0 + jj - 1); }
        g2(j0 + nl - 1);
      }
      stamp(0);
      // the X slabs are dead once every first-GEMM MMA has retired: their first half takes this CTA's 64 residual rows
      mbar_wait(x_empty, 0);
      mbar_expect_tx(xchg_full, (uint32_t)kSlabs * kHSlab);                // the peer's partial: this many bytes will land in [40K, 80K)
      mbar_arrive_remote(map_to_cta(smem_u32(xchg_ready), peer));          // ... and from now on the peer may send them
      mbar_expect_tx(r_full, (uint32_t)kSlabs * kHSlab);
      for (int kb = 0; kb < kSlabs; ++kb) tma_load_2d(&map_r, r_full, xs + (size_t)kb * kHSlab, kb * 32, tile * BM + (int)rank * HR);
      stamp(0);
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    constexpr uint32_t idesc1 = make_idesc(BM, CH);
    constexpr uint32_t idesc2 = make_idesc(BM, N2);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t act_cnt[2] = {0, 0};
    stamp(1);
    mbar_wait(x_full, 0);
    stamp(1);
    tcgen05_fence_after();
    auto g1 = [&](int jj) {   // acc1[jj&1] = X . W1_chunk^T
      const uint32_t d = tmem_base + (uint32_t)((jj & 1) * CH);
      for (int kb = 0; kb < p.kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint64_t da = make_smem_desc(smem_u32(xs + (size_t)kb * kXSlab));
          const uint64_t db = make_smem_desc(smem_u32(ring + (size_t)stage * kRing));
#pragma unroll
          for (int k = 0; k < 4; ++k)   // (the last slab of K1 = 144 holds 16 columns: two k-steps, not four)
            if (kb * 4 + k < p.ks1) umma_tf32(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc1, (kb > 0 || k > 0) ? 1u : 0u);
          tcgen05_commit(&empty_bar[stage]);
          if (kb == p.kb1 - 1) {
            tcgen05_commit(&acc1_full[jj & 1]);
            if (jj == nl - 1) tcgen05_commit(x_empty);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    };
    auto g2 = [&](int jj) {   // acc2 += act(acc1[jj&1]) . W2_chunk^T   (A operand from TMEM)
      mbar_wait(&act_done[jj & 1], act_cnt[jj & 1] & 1);
      act_cnt[jj & 1]++;
      tcgen05_fence_after();
      const uint32_t a = tmem_base + (uint32_t)((jj & 1) * CH);
      for (int kb = 0; kb < KB2; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint64_t db = make_smem_desc(smem_u32(ring + (size_t)stage * kRing));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int step = kb * 4 + k;
            if (step < KSTEPS2) umma_tf32_ts(tmem_acc2, a + (uint32_t)(8 * step), db + (uint64_t)(2 * k), idesc2, (jj > 0 || step > 0) ? 1u : 0u);
          }
          tcgen05_commit(&empty_bar[stage]);
          if (jj == nl - 1 && kb == KB2 - 1) tcgen05_commit(acc2_full);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    };
    if constexpr (DIRECT) {
      // acc2 = X[:, my K half] . W2[:, my K half]^T : global k-steps [9 rank, 9 rank + 9), four per 32-column slab
      const int ks0 = 9 * (int)rank, ks1 = ks0 + 9;
      bool first = true;
      for (int i = 0; i < 3; ++i) {
        const int kb = 2 * (int)rank + i;                    // slab of X and of the weights
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint64_t da = make_smem_desc(smem_u32(xs + (size_t)kb * kXSlab));
          const uint64_t db = make_smem_desc(smem_u32(ring + (size_t)stage * kRing));
          for (int k = 0; k < 4; ++k) {
            const int step = 4 * kb + k;
            if (step >= ks0 && step < ks1) {
              umma_tf32(tmem_acc2, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2, first ? 0u : 1u);
              first = false;
            }
          }
          tcgen05_commit(&empty_bar[stage]);
          if (i == 2) {
            tcgen05_commit(acc2_full);
            tcgen05_commit(x_empty);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      stamp(1);
    } else {
      g1(0);
      stamp(1);
      for (int jj = 1; jj < nl; ++jj) { g1(jj); stamp(1); g2(jj - 1); stamp(1); }
      g2(nl - 1);
      stamp(1);
    }
  } else {
    // ===================================================================== activation, partial exchange, final epilogue (warps 2..9)
    const int quad = warp & 3;                       // TMEM lane quadrant (hardware: warp id % 4)
    const int half = (warp - 2) >> 2;                // which half of the columns this warp handles
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    constexpr int kUnits = CH / 16;
    constexpr int kSplit = (kUnits + 1) / 2;
    const int cu0 = half ? kSplit : 0, cu1 = half ? kUnits : kSplit;
    uint32_t full_cnt[2] = {0, 0};
    const int n1 = nl * CH;
    const int et = threadIdx.x - 64;                   // 0..255 among these warps
    if (!DIRECT)
      for (int i = et; i < n1; i += 256) pcache[i] = p.bias1[j0 * CH + i];
    for (int i = et; i < N2; i += 256) {
      pcache[n1 + i] = p.ep.bias[i];
      pcache[n1 + N2 + i] = p.ep.ln1_g[i];
      pcache[n1 + 2 * N2 + i] = p.ep.ln1_b[i];
      pcache[n1 + 3 * N2 + i] = p.ep.ln2_g ? p.ep.ln2_g[i] : 0.f;
      pcache[n1 + 4 * N2 + i] = p.ep.ln2_g ? p.ep.ln2_b[i] : 0.f;
    }
    const int role = (warp == 2) ? 2 : (warp == 4 ? 3 : -1);
    if (role >= 0) stamp(role);
    pdl_wait();   // (the parameter cache above only reads weights)
    epi_bar_sync<256>();
    if (role >= 0) stamp(role);
    TcParams ep = p.ep;
    ep.bias = pcache + n1;
    ep.ln1_g = pcache + n1 + N2;
    ep.ln1_b = pcache + n1 + 2 * N2;
    if (p.ep.ln2_g) {
      ep.ln2_g = pcache + n1 + 3 * N2;
      ep.ln2_b = pcache + n1 + 4 * N2;
    }
    const float* bias1 = pcache;
    for (int jj = 0; jj < nl; ++jj) {
      const int a = jj & 1;
      mbar_wait(&acc1_full[a], full_cnt[a] & 1);
      full_cnt[a]++;
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + lane_addr + (uint32_t)(a * CH);
      uint32_t raw[kSplit][16];
#pragma unroll
      for (int i = 0; i < kSplit; ++i)
        if (cu0 + i < cu1) tmem_ld16_nowait(taddr + (uint32_t)(16 * (cu0 + i)), raw[i]);   // warp-uniform predicate
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < kSplit; ++i) {
        if (cu0 + i < cu1) {
          const int u = cu0 + i;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b = *reinterpret_cast<const float4*>(bias1 + jj * CH + 16 * u + 4 * q);
            raw[i][4 * q + 0] = tf32_rn_bits(swish_fast(__uint_as_float(raw[i][4 * q + 0]) + b.x));
            raw[i][4 * q + 1] = tf32_rn_bits(swish_fast(__uint_as_float(raw[i][4 * q + 1]) + b.y));
            raw[i][4 * q + 2] = tf32_rn_bits(swish_fast(__uint_as_float(raw[i][4 * q + 2]) + b.z));
            raw[i][4 * q + 3] = tf32_rn_bits(swish_fast(__uint_as_float(raw[i][4 * q + 3]) + b.w));
          }
          tmem_st16(taddr + (uint32_t)(16 * u), raw[i]);
        }
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&act_done[a]);
      if (role >= 0) stamp(role);
    }
    // ---- this CTA's partial accumulator is complete (so every MMA that read the X slabs has retired)
    mbar_wait(acc2_full, 0);
    if (role >= 0) stamp(role);
    tcgen05_fence_after();
    const bool owner = ((uint32_t)(quad >> 1) == rank);     // quadrants 0,1 = rows 0..63 (rank 0), quadrants 2,3 = rows 64..127 (rank 1)
    const int trow = (quad & 1) * 32 + lane;                // row inside the 64-row half
    if (!owner) {
      // ---- stage the peer's 64 rows of my partial accumulator in the idle weight ring (every slab has been consumed: acc2_full),
      // then ONE thread sends the 40 KB tile into the peer's exchange area with a bulk DSMEM copy
      uint8_t* ship = ring + (size_t)kSlabs * kHSlab;        // ring [40K, 80K): the first 40 KB stage the second output later
#pragma unroll 1
      for (int u = cu0; u < cu1; ++u) {
        uint32_t raw[16];
        tmem_ld16_nowait(tmem_acc2 + lane_addr + (uint32_t)(16 * u), raw);
        tmem_ld_wait();
        const int s = u >> 1;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int qq = ((u & 1) << 2) | q;
          *reinterpret_cast<float4*>(ship + (size_t)s * kHSlab + (size_t)trow * 128 + (((qq ^ (trow & 7)) & 7) << 4)) =
              make_float4(__uint_as_float(raw[4 * q + 0]), __uint_as_float(raw[4 * q + 1]), __uint_as_float(raw[4 * q + 2]),
                          __uint_as_float(raw[4 * q + 3]));
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk copy
      asm volatile("bar.sync 2, 128;" ::: "memory");                 // the four shipping warps
      if ((quad & 1) == 0 && half == 0) {                            // one warp: five 8 KB slab copies in flight at once
        if (lane == 0) mbar_wait_cluster(xchg_ready, 0);
        __syncwarp();
        if (role >= 0) stamp(role);
        if (lane < kSlabs)
          bulk_copy_to_cta(map_to_cta(smem_u32(xchg + (size_t)lane * kHSlab), peer), ship + (size_t)lane * kHSlab, kHSlab,
                           map_to_cta(smem_u32(xchg_full), peer));
      }
      if (role >= 0) stamp(role);
    } else {
      // ---- finish my 64 rows: x = resid + alpha * (mine + peer's + bias2), LayerNorm(s), TMA stores
      mbar_wait(r_full, 0);
      if (role >= 0) stamp(role);
      mbar_wait(xchg_full, 0);
      if (role >= 0) stamp(role);
      epilogue_ln_tma<EPI, N2, HR, 2, 128>(ep, tmem_acc2 + lane_addr, xs, &map_c, &map_c2, tile * BM + (int)rank * HR, trow,
                                           (quad & 1) == 0 && half == 0 && lane == 0, half, statbuf, ring, xchg);
      if (role >= 0) stamp(role);
    }
    tcgen05_fence_before();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
  // the bulk copy reads THIS CTA's ring until the peer has received it (the peer's owners waited on xchg_full before they got
  // here): neither CTA may retire before both are done
  cluster_sync_all();
}

size_t pair_smem() {
  return (size_t)kSlabs * kXSlab + (size_t)STAGES * kRing + 1024 + 256 + HR * 2 * 4 * 4 + (2 * CH + 5 * N2) * 4 + 64;
}

template <int EPI, bool DIRECT>
int launch_pair_t(TcContext& ctx, const CUtensorMap& mx, const CUtensorMap& m1, const CUtensorMap& m2, const CUtensorMap& mr,
                  const CUtensorMap& mc, const CUtensorMap& mc2, const PairParams& pp, cudaStream_t stream) {
  auto kern = gemm_chain_pair_kernel<EPI, DIRECT>;
  const size_t smem = pair_smem();
  static PerDeviceSmem configured;
  if (configured.need(smem)) B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  (void)ctx;
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    const char* e = getenv("B200ASR_PAIR_DBG");
    dbg_on = (e && e[0] == '1') ? 1 : 0;
    if (dbg_on) cudaMalloc(&dbg, sizeof(long long) * 128);
  }
  PairParams p2 = pp;
  p2.dbg = dbg_on ? dbg : nullptr;
  if (dbg_on) cudaMemset(dbg, 0, sizeof(long long) * 128);
  B200_CUDA_OK(launch_k_cluster(kern, dim3(2 * pp.num_m_tiles), dim3(kPairThreads), smem, stream, 2, mx, m1, m2, mr, mc, mc2, p2));
  if (dbg_on) {
    long long hbuf[128];
    cudaDeviceSynchronize();
    cudaMemcpy(hbuf, dbg, sizeof(hbuf), cudaMemcpyDeviceToHost);
    long long t0 = 0;
    for (int i = 0; i < 128; ++i) if (hbuf[i] && (!t0 || hbuf[i] < t0)) t0 = hbuf[i];
    const char* names[4] = {"producer", "mma", "ship(w2)", "own(w4)"};
    for (int r = 0; r < 4; ++r) {
      fprintf(stderr, "pair-dbg %s:", names[r]);
      for (int i = 0; i < 32 && hbuf[r * 32 + i]; ++i) fprintf(stderr, " %lld", hbuf[r * 32 + i] - t0);
      fprintf(stderr, "\n");
    }
  }
  return 0;
}

}  // namespace

bool tc_chain_pair_supported(const ChainGemmParams& p, int epilogue) {
  if (!tc_chain_supported(p, epilogue)) return false;
  if (p.N2 != N2 || p.K1 <= 128 || p.K1 > 160) return false;
  const int nch = p.N1 / CH;
  return p.N1 % CH == 0 && nch >= 2 && nch % 2 == 0 && nch / 2 <= 2;   // pcache holds at most 2 local chunks of bias1
}

// DIRECT mode: ChainGemmParams with N1 == 0 and W1 == bias1 == null means C/C2 = LN-epilogue(resid + alpha * (X . W2^T + bias2)),
// X [M, 144], W2 [144, 144] (the attention out-projection with its residual + LayerNorm)
bool tc_pair_direct_supported(const ChainGemmParams& p, int epilogue) {
  if (epilogue != EPI_RESID_LN && epilogue != EPI_RESID_LN2) return false;
  if (p.N1 != 0 || p.N2 != N2 || p.K1 != N2 || p.ldx != p.K1 || p.M <= 0) return false;
  if (!p.bias2 || !p.resid || !p.C || !p.C2 || !p.ln1_g || !p.X || !p.W2) return false;
  return ((reinterpret_cast<uintptr_t>(p.X) | reinterpret_cast<uintptr_t>(p.W2)) & 15) == 0;
}

int launch_gemm_chain_pair(TcContext& ctx, const ChainGemmParams& p, int epilogue, cudaStream_t stream) {
  if (!ctx.ready) {
    snprintf(g_errbuf, sizeof(g_errbuf), "gemm_chain_pair: tensor-map encoder not initialised");
    return 1;
  }
  PairParams pp{};
  pp.ep.bias = p.bias2; pp.ep.resid = p.resid; pp.ep.C = p.C; pp.ep.C2 = p.C2; pp.ep.M = p.M; pp.ep.N = p.N2; pp.ep.K = p.N1;
  pp.ep.ldc = p.N2; pp.ep.alpha = p.alpha; pp.ep.ln1_g = p.ln1_g; pp.ep.ln1_b = p.ln1_b; pp.ep.ln2_g = p.ln2_g; pp.ep.ln2_b = p.ln2_b;
  pp.ep.ln_eps = p.ln_eps;
  pp.ep.round_out = (p.round_c && p.ln2_g == nullptr) ? 1 : 0;
  pp.bias1 = p.bias1;
  const bool direct = (p.N1 == 0);
  pp.nl = direct ? 0 : p.N1 / CH / 2;
  pp.kb1 = ceil_div(p.K1, 32);
  pp.ks1 = ceil_div(p.K1, 8);
  pp.num_m_tiles = ceil_div(p.M, BM);
  const cuuint32_t ones[2] = {1, 1};
  CUtensorMap mx, m1, m2, mr, mc, mc2;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)p.K1, (cuuint64_t)p.M};
    const cuuint64_t strides[1] = {(cuuint64_t)p.ldx * 4};
    const cuuint32_t box[2] = {32, BM};
    if (encode_map(ctx, &mx, p.X, 2, dims, strides, box, ones)) return 1;
  }
  {
    const int kw2 = direct ? p.K1 : p.N1;                    // K extent of the second operand
    const cuuint64_t dims[2] = {(cuuint64_t)kw2, (cuuint64_t)p.N2};
    const cuuint64_t strides[1] = {(cuuint64_t)kw2 * 4};
    const cuuint32_t box[2] = {32, N2};
    if (encode_map(ctx, &m2, p.W2, 2, dims, strides, box, ones)) return 1;
  }
  if (direct) {
    m1 = m2;                                                 // (not used by the DIRECT kernel)
  } else {
    const cuuint64_t dims[2] = {(cuuint64_t)p.K1, (cuuint64_t)p.N1};
    const cuuint64_t strides[1] = {(cuuint64_t)p.K1 * 4};
    const cuuint32_t box[2] = {32, CH};
    if (encode_map(ctx, &m1, p.W1, 2, dims, strides, box, ones)) return 1;
  }
  {
    // residual in / outputs: [M, N2] in tiles of 64 rows x 32-column slabs (stores clip the M and column tails)
    const cuuint64_t dims[2] = {(cuuint64_t)p.N2, (cuuint64_t)p.M};
    const cuuint64_t strides[1] = {(cuuint64_t)p.N2 * 4};
    const cuuint32_t box[2] = {32, HR};
    if (encode_map(ctx, &mr, p.resid, 2, dims, strides, box, ones)) return 1;
    if (encode_map(ctx, &mc, p.C, 2, dims, strides, box, ones)) return 1;
    if (encode_map(ctx, &mc2, p.C2, 2, dims, strides, box, ones)) return 1;
  }
  if (direct) {
    if (epilogue == EPI_RESID_LN) return launch_pair_t<EPI_RESID_LN, true>(ctx, mx, m1, m2, mr, mc, mc2, pp, stream);
    return launch_pair_t<EPI_RESID_LN2, true>(ctx, mx, m1, m2, mr, mc, mc2, pp, stream);
  }
  if (epilogue == EPI_RESID_LN) return launch_pair_t<EPI_RESID_LN, false>(ctx, mx, m1, m2, mr, mc, mc2, pp, stream);
  return launch_pair_t<EPI_RESID_LN2, false>(ctx, mx, m1, m2, mr, mc, mc2, pp, stream);
}

}  // namespace b200asr
