// Two chained GEMMs in one tcgen05 kernel:   Y = epilogue( act(X . W1^T + b1) . W2^T + b2 )
//
// Used for the FFModule (conformer_blocks.py:126-134: Dense 4D + swish + Dense D, 0.5-residual, LayerNorm of the next
// module) and for the tail of the ConvModule (:214-218: pointwise conv with folded BatchNorm + swish, pointwise conv,
// residual, LayerNorm).  The hidden activations H = act(X W1^T + b1) ([M, 576] / [M, 288] per call, 18 MB at the benchmark
// shape) never leave the SM: they are produced chunk by chunk (CH columns) into TMEM accumulators, activated IN PLACE by the
// epilogue warps (tcgen05.ld -> bias + swish -> round to tf32 -> tcgen05.st) and consumed straight from TMEM as the A operand
// of the second GEMM (tcgen05.mma, TS form), whose accumulator (128 x N2) stays in TMEM for the whole tile.
//
//   warp 0   TMA producer: the X tile (all K1 slabs, resident for the tile) + a ring of weight slabs in MMA issue order
//   warp 1   MMA issuer + TMEM owner; issue order G1(0) G1(1) G2(0) G1(2) G2(1) ... G2(n-1) so that the tensor core works
//            on chunk j+1 while the epilogue warps activate chunk j
//   warps 2-5 activation of every chunk, then the final fused residual + LayerNorm epilogue (tc_common.cuh: epilogue_ln)
// TMEM columns: [0,CH) and [CH,2CH) = double-buffered chunk accumulators / activated A operand, [2CH, 2CH+N2) = output.
#include "tc_common.cuh"

#include <cstdlib>

namespace b200asr {

using namespace tc;

namespace {

constexpr int kChainThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 activation / epilogue (two per TMEM lane quadrant)

struct ChainParams {
  TcParams ep;          // final epilogue (bias = b2, resid, C, C2, ln params, alpha, M, N = N2, ldc)
  const float* bias1;   // [N1]
  int n_chunks;         // N1 / CH
  int kb1;              // ceil(K1 / 32) slabs of X
  int ks1;              // ceil(K1 / 8) k-steps of the first GEMM
  int num_m_tiles;
  long long* dbg;       // optional timeline buffer (B200ASR_CHAIN_DBG=1): [role][event] clock64 stamps of CTA 0
};

template <int EPI, int CH, int N2, int STAGES>
__global__ void __launch_bounds__(kChainThreads, 1)
gemm_chain_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                  const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_r,
                  const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_c2, const ChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int BM = 128;
  constexpr uint32_t kXSlab = BM * 128;                               // 16 KB per 32-column slab of X
  constexpr uint32_t kRing = (CH > N2 ? CH : N2) * 128;               // one weight slab: rows x 32 floats
  constexpr int KB2 = (CH + 31) / 32;                                 // W2 slabs per chunk
  constexpr int KSTEPS2 = CH / 8;                                     // k-steps of the second GEMM per chunk
  // 1024-byte alignment by POINTER arithmetic on the shared array (an integer round trip would strip the address space and turn every
  // access through a derived pointer into a generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  uint8_t* xs = smem;                                                 // kb1 slabs
  uint8_t* ring = xs + (size_t)p.kb1 * kXSlab;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)STAGES * kRing);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* x_full = empty_bar + STAGES;
  uint64_t* x_empty = x_full + 1;
  uint64_t* acc1_full = x_empty + 1;     // [2]
  uint64_t* act_done = acc1_full + 2;    // [2]
  uint64_t* acc2_full = act_done + 2;
  uint64_t* acc2_empty = acc2_full + 1;
  uint64_t* r_full = acc2_empty + 1;     // residual tile landed in the (recycled) X slabs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_full + 1);
  float* statbuf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + (((2 * STAGES + 9) * 8 + 4 + 15) / 16) * 16);   // [128][2][4] LN partial sums (16-byte aligned: bars sit on a 1024-byte boundary)
  float* pcache = statbuf + 128 * 2 * 4;                      // bias1[N1] | bias2 | ln1_g | ln1_b | ln2_g | ln2_b  (N2 each)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nch = p.n_chunks;
  int ev = 0;
  auto stamp = [&](int role) {
    if (p.dbg && blockIdx.x == 0 && lane == 0 && ev < 64) p.dbg[role * 64 + ev++] = clock64();
  };

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
    mbar_init(acc2_empty, 8);
    mbar_init(r_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc2 = tmem_base + 2 * CH;
  pdl_trigger();   // TMEM is allocated: the next kernel's CTAs may start their prologue under this kernel's main body

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, xphase = 0;
      auto ring_load = [&](const CUtensorMap* map, int c0, int c1, uint32_t bytes) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], bytes);
        tma_load_2d(map, &full_bar[stage], ring + (size_t)stage * kRing, c0, c1);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      };
      auto g1 = [&](int j, int kb0 = 0) { for (int kb = kb0; kb < p.kb1; ++kb) ring_load(&map_w1, kb * 32, j * CH, CH * 128); };
      auto g2 = [&](int j) { for (int kb = 0; kb < KB2; ++kb) ring_load(&map_w2, j * CH + kb * 32, 0, N2 * 128); };
      // the first weight slabs are constants: they are requested BEFORE griddepcontrol.wait, i.e. while the previous kernel
      // of the schedule is still draining; X and the residual stream (its outputs) only after the wait
      // (no more than the ring holds: the MMA warp cannot free a stage before X has arrived)
      bool first = true;
      const int npre = p.kb1 < STAGES ? p.kb1 : STAGES;
      if (blockIdx.x < p.num_m_tiles)
        for (int kb = 0; kb < npre; ++kb) ring_load(&map_w1, kb * 32, 0, CH * 128);
      pdl_wait();
      for (int tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x) {
        stamp(0);
        mbar_wait(acc2_empty, xphase ^ 1);                    // previous tile's epilogue has left the slabs (they stage its output)
        mbar_expect_tx(x_full, (uint32_t)p.kb1 * kXSlab);
        for (int kb = 0; kb < p.kb1; ++kb) tma_load_2d(&map_x, x_full, xs + (size_t)kb * kXSlab, kb * 32, tile * BM);
        stamp(0);
        g1(0, first ? npre : 0);
        first = false;
        stamp(0);
        for (int j = 1; j < nch; ++j) { g1(j); stamp(0); g2(j - 1); stamp(0); }
        g2(nch - 1);
        // the X slabs are dead once every first-GEMM MMA has retired: recycle them for the residual tile of this output
        mbar_wait(x_empty, xphase);
        xphase ^= 1;
        mbar_expect_tx(r_full, (uint32_t)p.kb1 * kXSlab);
        for (int kb = 0; kb < p.kb1; ++kb) tma_load_2d(&map_r, r_full, xs + (size_t)kb * kXSlab, kb * 32, tile * BM);
        stamp(0);
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    constexpr uint32_t idesc1 = make_idesc(BM, CH);
    constexpr uint32_t idesc2 = make_idesc(BM, N2);
    int stage = 0;
    uint32_t phase = 0, xphase = 0, tphase = 0;
    uint32_t full1_cnt[2] = {0, 0}, act_cnt[2] = {0, 0};
    (void)full1_cnt;
    for (int tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x) {
      stamp(1);
      mbar_wait(x_full, xphase);
      stamp(1);
      xphase ^= 1;
      tcgen05_fence_after();
      auto g1 = [&](int j) {   // acc1[j&1] = X . W1_j^T
        const uint32_t d = tmem_base + (uint32_t)((j & 1) * CH);
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
              tcgen05_commit(&acc1_full[j & 1]);
              if (j == nch - 1) tcgen05_commit(x_empty);       // X slabs free once every first-GEMM MMA of the tile retired
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      };
      auto g2 = [&](int j) {   // acc2 += act(acc1[j&1]) . W2_j^T   (A operand from TMEM)
        stamp(1);
        mbar_wait(&act_done[j & 1], act_cnt[j & 1] & 1);
        stamp(1);
        act_cnt[j & 1]++;
        tcgen05_fence_after();
        const uint32_t a = tmem_base + (uint32_t)((j & 1) * CH);
        for (int kb = 0; kb < KB2; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (lane == 0) {
            const uint64_t db = make_smem_desc(smem_u32(ring + (size_t)stage * kRing));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int step = kb * 4 + k;
              if (step < KSTEPS2) umma_tf32_ts(tmem_acc2, a + (uint32_t)(8 * step), db + (uint64_t)(2 * k), idesc2, (j > 0 || step > 0) ? 1u : 0u);
            }
            tcgen05_commit(&empty_bar[stage]);
            if (j == nch - 1 && kb == KB2 - 1) tcgen05_commit(acc2_full);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      };
      mbar_wait(acc2_empty, tphase ^ 1);                      // previous tile's final epilogue has drained acc2
      tcgen05_fence_after();
      stamp(1);
      g1(0);
      stamp(1);
      for (int j = 1; j < nch; ++j) { g1(j); stamp(1); g2(j - 1); stamp(1); }
      g2(nch - 1);
      stamp(1);
      tphase ^= 1;
    }
  } else {
    // ===================================================================== activation + final epilogue (warps 2..9)
    const int quad = warp & 3;                       // TMEM lane quadrant (hardware: warp id % 4)
    const int half = (warp - 2) >> 2;                // which half of the columns this warp handles
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    constexpr int kUnits = CH / 16;
    constexpr int kSplit = (kUnits + 1) / 2;
    const int cu0 = half ? kSplit : 0, cu1 = half ? kUnits : kSplit;
    uint32_t full_cnt[2] = {0, 0};
    uint32_t tphase = 0;
    // every per-column parameter is copied to shared memory once, while the first MMAs are in flight: the epilogue
    // warps would otherwise pay a first-touch L2 round trip per chunk for them
    const int n1 = nch * CH;
    const int et = threadIdx.x - 64;                   // 0..255 among the epilogue warps
    for (int i = et; i < n1; i += 256) pcache[i] = p.bias1[i];
    for (int i = et; i < N2; i += 256) {
      pcache[n1 + i] = p.ep.bias[i];
      pcache[n1 + N2 + i] = p.ep.ln1_g[i];
      pcache[n1 + 2 * N2 + i] = p.ep.ln1_b[i];
      pcache[n1 + 3 * N2 + i] = p.ep.ln2_g ? p.ep.ln2_g[i] : 0.f;
      pcache[n1 + 4 * N2 + i] = p.ep.ln2_g ? p.ep.ln2_b[i] : 0.f;
    }
    pdl_wait();   // (the parameter cache above only reads weights: it fills while the previous kernel is still running)
    epi_bar_sync<256>();
    TcParams ep = p.ep;
    ep.bias = pcache + n1;
    ep.ln1_g = pcache + n1 + N2;
    ep.ln1_b = pcache + n1 + 2 * N2;
    if (p.ep.ln2_g) {
      ep.ln2_g = pcache + n1 + 3 * N2;
      ep.ln2_b = pcache + n1 + 4 * N2;
    }
    const float* bias1 = pcache;
    for (int tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x) {
      for (int j = 0; j < nch; ++j) {
        const int a = j & 1;
        if (warp == 2) stamp(2);
        mbar_wait(&acc1_full[a], full_cnt[a] & 1);
        if (warp == 2) stamp(2);
        full_cnt[a]++;
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + lane_addr + (uint32_t)(a * CH);
        // all of this warp's columns are fetched from TMEM before the first use (one exposed TMEM round trip per chunk)
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
              const float4 b = *reinterpret_cast<const float4*>(bias1 + j * CH + 16 * u + 4 * q);
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
        if (warp == 2) stamp(2);
      }
      // final epilogue on acc2
      if (warp == 2) stamp(2);
      mbar_wait(acc2_full, tphase);
      if (warp == 2) stamp(2);
      tphase ^= 1;
      tcgen05_fence_after();
      mbar_wait(r_full, tphase ^ 1);                          // residual tile is in the slabs (tphase already flipped above)
      // the weight ring is idle from here until the next tile's first load (the producer is gated on acc2_empty): use it as
      // the second staging tile so the C store is not waited for before C2 is produced
      constexpr bool kRingFits = (size_t)STAGES * kRing >= (size_t)BM * ((N2 + 31) / 32) * 128;
      epilogue_ln_tma<EPI, N2, BM, 2>(ep, tmem_acc2 + lane_addr, xs, &map_c, &map_c2, tile * BM, quad * 32 + lane,
                                      warp == 2 && lane == 0, half, statbuf, kRingFits ? ring : nullptr);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty);
      if (warp == 2) stamp(2);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

template <int CH, int N2, int STAGES>
size_t chain_smem(int kb1) {
  return (size_t)kb1 * 128 * 128 + (size_t)STAGES * (CH > N2 ? CH : N2) * 128 + 1024 + 256 + 128 * 2 * 4 * 4 /*LayerNorm partial sums*/ + (1024 + 5 * 256) * 4 /*parameter cache*/ + 16;
}

template <int EPI, int CH, int N2, int STAGES>
int launch_chain_t(TcContext& ctx, const CUtensorMap& mx, const CUtensorMap& m1, const CUtensorMap& m2, const CUtensorMap& mr,
                   const CUtensorMap& mc, const CUtensorMap& mc2, const ChainParams& cp, cudaStream_t stream) {
  auto kern = gemm_chain_kernel<EPI, CH, N2, STAGES>;
  const size_t smem = chain_smem<CH, N2, STAGES>(cp.kb1);
  static PerDeviceSmem configured;
  if (configured.need(smem)) B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = cp.num_m_tiles < ctx.num_sms ? cp.num_m_tiles : ctx.num_sms;
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    const char* e = getenv("B200ASR_CHAIN_DBG");
    dbg_on = (e && e[0] == '1') ? 1 : 0;
    if (dbg_on) cudaMalloc(&dbg, sizeof(long long) * 192);
  }
  ChainParams cp2 = cp;
  cp2.dbg = dbg_on ? dbg : nullptr;
  if (dbg_on) cudaMemset(dbg, 0, sizeof(long long) * 192);
  B200_CUDA_OK(launch_k(kern, dim3(grid), dim3(kChainThreads), smem, stream, mx, m1, m2, mr, mc, mc2, cp2));
  B200_CUDA_OK(cudaGetLastError());
  if (dbg_on) {
    long long hbuf[192];
    cudaDeviceSynchronize();
    cudaMemcpy(hbuf, dbg, sizeof(hbuf), cudaMemcpyDeviceToHost);
    long long t0 = 0;
    for (int i = 0; i < 192; ++i) if (hbuf[i] && (!t0 || hbuf[i] < t0)) t0 = hbuf[i];
    const char* names[3] = {"producer", "mma", "epilogue"};
    for (int r = 0; r < 3; ++r) {
      fprintf(stderr, "chain-dbg %s:", names[r]);
      for (int i = 0; i < 64 && hbuf[r * 64 + i]; ++i) fprintf(stderr, " %lld", hbuf[r * 64 + i] - t0);
      fprintf(stderr, "\n");
    }
  }
  return 0;
}

}  // namespace

bool tc_chain_supported(const ChainGemmParams& p, int epilogue) {
  if (epilogue != EPI_RESID_LN && epilogue != EPI_RESID_LN2) return false;
  if (p.M <= 0 || p.K1 % 4 != 0 || p.N1 % 4 != 0) return false;
  if (!p.bias1 || !p.bias2 || !p.resid || !p.C || !p.C2 || !p.ln1_g) return false;
  if ((reinterpret_cast<uintptr_t>(p.X) | reinterpret_cast<uintptr_t>(p.W1) | reinterpret_cast<uintptr_t>(p.W2)) & 15) return false;
  // (the X slabs are recycled for the N2-wide residual/output tile: needs ceil(K1/32) == ceil(N2/32))
  if (p.N2 == 144) return p.N1 % 144 == 0 && p.K1 > 128 && p.K1 <= 160;
  if (p.N2 == 256) return p.N1 % 128 == 0 && p.K1 > 224 && p.K1 <= 256;
  return false;
}

int launch_gemm_chain(TcContext& ctx, const ChainGemmParams& p, int epilogue, cudaStream_t stream) {
  if (!ctx.ready) {
    snprintf(g_errbuf, sizeof(g_errbuf), "gemm_chain: tensor-map encoder not initialised");
    return 1;
  }
  const int CH = (p.N2 == 144) ? 144 : 128;
  ChainParams cp{};
  cp.ep.bias = p.bias2; cp.ep.resid = p.resid; cp.ep.C = p.C; cp.ep.C2 = p.C2; cp.ep.M = p.M; cp.ep.N = p.N2; cp.ep.K = p.N1;
  cp.ep.ldc = p.N2; cp.ep.alpha = p.alpha; cp.ep.ln1_g = p.ln1_g; cp.ep.ln1_b = p.ln1_b; cp.ep.ln2_g = p.ln2_g; cp.ep.ln2_b = p.ln2_b;
  cp.ep.ln_eps = p.ln_eps;
  cp.ep.round_out = (p.round_c && p.ln2_g == nullptr) ? 1 : 0;
  cp.bias1 = p.bias1;
  cp.n_chunks = p.N1 / CH;
  cp.kb1 = ceil_div(p.K1, 32);
  cp.ks1 = ceil_div(p.K1, 8);
  cp.num_m_tiles = ceil_div(p.M, 128);
  const cuuint32_t ones[2] = {1, 1};
  CUtensorMap mx, m1, m2;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)p.K1, (cuuint64_t)p.M};
    const cuuint64_t strides[1] = {(cuuint64_t)p.ldx * 4};
    const cuuint32_t box[2] = {32, 128};
    if (encode_map(ctx, &mx, p.X, 2, dims, strides, box, ones)) return 1;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)p.K1, (cuuint64_t)p.N1};
    const cuuint64_t strides[1] = {(cuuint64_t)p.K1 * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)CH};
    if (encode_map(ctx, &m1, p.W1, 2, dims, strides, box, ones)) return 1;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)p.N1, (cuuint64_t)p.N2};
    const cuuint64_t strides[1] = {(cuuint64_t)p.N1 * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)p.N2};
    if (encode_map(ctx, &m2, p.W2, 2, dims, strides, box, ones)) return 1;
  }
  CUtensorMap mr, mc, mc2;
  {
    // residual in / outputs: [M, N2] tiles of 128 rows x 32-column slabs (stores clip the M and column tails)
    const cuuint64_t dims[2] = {(cuuint64_t)p.N2, (cuuint64_t)p.M};
    const cuuint64_t strides[1] = {(cuuint64_t)p.N2 * 4};
    const cuuint32_t box[2] = {32, 128};
    if (encode_map(ctx, &mr, p.resid, 2, dims, strides, box, ones)) return 1;
    if (encode_map(ctx, &mc, p.C, 2, dims, strides, box, ones)) return 1;
    if (encode_map(ctx, &mc2, p.C2, 2, dims, strides, box, ones)) return 1;
  }
  if (p.N2 == 144) {
    if (epilogue == EPI_RESID_LN) return launch_chain_t<EPI_RESID_LN, 144, 144, 6>(ctx, mx, m1, m2, mr, mc, mc2, cp, stream);
    return launch_chain_t<EPI_RESID_LN2, 144, 144, 6>(ctx, mx, m1, m2, mr, mc, mc2, cp, stream);
  }
  if (epilogue == EPI_RESID_LN) return launch_chain_t<EPI_RESID_LN, 128, 256, 2>(ctx, mx, m1, m2, mr, mc, mc2, cp, stream);
  return launch_chain_t<EPI_RESID_LN2, 128, 256, 2>(ctx, mx, m1, m2, mr, mc, mc2, cp, stream);
}

}  // namespace b200asr
