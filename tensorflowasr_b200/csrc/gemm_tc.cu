// tcgen05 tf32 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] . W[N,K]^T), fp32 in HBM, tf32 tensor-core inputs,
// fp32 accumulation in TMEM.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer   cp.async.bulk.tensor (SWIZZLE_128B boxes of 32 fp32 = one 128-byte swizzle row) into a
//                              4-stage shared-memory ring, completion on mbarriers
//   warp 1      MMA issuer     one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BLOCK_N, K=8 per
//                              instruction, 4 per stage); tcgen05.commit releases smem stages / publishes accumulators;
//                              also owns the TMEM allocation (512 columns = two accumulator buffers)
//   warps 2..5  epilogue       tcgen05.ld 32 lanes x 16 columns at a time (thread == output row), fused bias / ReLU /
//                              swish / GLU / residual, vectorised st.global; overlaps the next tile's MMAs through the
//                              double-buffered accumulator
// Ragged edges need no code: TMA zero-fills out-of-bounds rows (M, N tails) and columns (K tail: 144 = 4.5 x 32),
// the epilogue masks stores.
// The second subsampling conv (3x3, stride 2, 'same'; conformer_blocks.py:81-85) runs through the same kernel as an
// implicit GEMM: its A operand is fetched by a 4-D tensor map over conv1's NHWC activations with traversal stride 2
// along time and frequency, one (tap, 32-channel) slab per pipeline stage; 'same' padding is TMA out-of-bounds fill.
#include "gemm_tc.cuh"
#include "tc_common.cuh"

#include <cstring>

namespace b200asr {

using namespace tc;

namespace {

// pipeline depth: as many stages as fit beside the epilogue scratch, at most 6 (plain epilogues; the first `stages` weight slabs are
// requested before griddepcontrol.wait, i.e. under the previous kernel's tail) / 4 (LayerNorm epilogues, whose staging tile
// BLOCK_M x BLOCK_N fp32 shares the shared memory)
__host__ __device__ constexpr int stages_for(bool is_ln, int bn, int bm) {
  if (!is_ln) {
    for (int s = 6; s > 4; --s)
      if (s * (bm + bn) * 128 + 4 * kWsmFloats * 4 + 1280 <= 226 * 1024) return s;
    return 4;
  }
  const int tile = bm * ((bn + 31) / 32) * 128;
  for (int s = 4; s >= 2; --s)
    if (s * (bm + bn) * 128 + tile <= 220 * 1024) return s;
  return 2;
}

// ------------------------------------------------------------------------------------------------ kernel
template <int EPI, int BLOCK_N, int BLOCK_M, bool F16 = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_c,
               const __grid_constant__ CUtensorMap map_c2, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t kABytes = BLOCK_M * BLOCK_K * 4;     // 16 KB
  constexpr uint32_t kBBytes = BLOCK_N * BLOCK_K * 4;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;     // multiple of 1024 (BLOCK_N multiple of 8)
  constexpr bool kIsLN = (EPI == EPI_RESID_LN || EPI == EPI_RESID_LN2 || EPI == EPI_BIAS_LN);
  constexpr bool kHasResidTile = (EPI == EPI_RESID_LN || EPI == EPI_RESID_LN2);
  constexpr int kStages = stages_for(kIsLN, BLOCK_N, BLOCK_M);
  constexpr int kNSlab = (BLOCK_N + 31) / 32;
  constexpr uint32_t kStageTile = kIsLN ? (uint32_t)BLOCK_M * kNSlab * 128 : 0;   // LN epilogues: residual-in / output staging tile
  // 1024-byte alignment by POINTER arithmetic on the shared array (an integer round trip would strip the address space and turn every
  // access through a derived pointer into a generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  uint8_t* stile = smem + kStages * kStageBytes;          // (LN epilogues) [BLOCK_M rows x kNSlab slabs of 32 columns], SW128
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stile + kStageTile);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* r_full = tmem_empty + 2;      // residual tile landed in `stile`
  uint64_t* r_empty = r_full + 1;         // epilogue is done with `stile`
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    if (kIsLN) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_r) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c2) : "memory");
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    mbar_init(r_full, 1);
    mbar_init(r_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();   // prologue resources are taken (TMEM allocated): the next kernel may start its own prologue
  const uint32_t a_rows_bytes = (p.a_mode == 1) ? (uint32_t)(p.bt * p.F2 * BLOCK_K * 4) : kABytes;
  constexpr int bk = F16 ? 2 * BLOCK_K : BLOCK_K;   // operand columns per 128-byte swizzle row (TMA coordinates count elements)
  // B operand = weights (constants): the first pipeline stages' weight slabs are requested BEFORE griddepcontrol.wait, under
  // the tail of the previous kernel; everything that kernel produced (A operand, residual) is touched only after the wait
  int b_prefetched = 0;
  if (warp == 0 && lane == 0 && (int)blockIdx.x < num_tiles) {
    const int nt0 = (int)blockIdx.x % p.num_n_tiles;
    b_prefetched = p.num_k_blocks < kStages ? p.num_k_blocks : kStages;
    for (int kb = 0; kb < b_prefetched; ++kb) {
      uint8_t* sb = smem + kb * kStageBytes + kABytes;
      mbar_expect_tx(&full_bar[kb], a_rows_bytes + kBBytes);
      if (p.a_mode == 0) {
        tma_load_2d(&map_b, &full_bar[kb], sb, kb * bk, nt0 * BLOCK_N);
      } else {
        const int tap = kb / p.kc, j = kb - tap * p.kc;
        tma_load_2d(&map_b, &full_bar[kb], sb, tap * p.D + j * bk, nt0 * BLOCK_N);
      }
    }
  }
  pdl_wait();      // the previous kernel's outputs (A operand, residual) are complete and visible

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, rphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.num_n_tiles, nt = tile - mt * p.num_n_tiles;
        const int n0 = nt * BLOCK_N;
        if (kHasResidTile) {   // residual tile -> staging (free once the previous tile's epilogue has stored its outputs)
          mbar_wait(r_empty, rphase ^ 1);
          rphase ^= 1;
          mbar_expect_tx(r_full, kStageTile);
          for (int sl = 0; sl < kNSlab; ++sl) tma_load_2d(&map_r, r_full, stile + (size_t)sl * BLOCK_M * 128, 32 * sl, mt * BLOCK_M);
        }
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          const bool b_done = b_prefetched > 0;               // this stage's weights (and its expect_tx) were issued before the wait
          if (b_done) --b_prefetched;
          else mbar_expect_tx(&full_bar[stage], a_rows_bytes + kBBytes);
          if (p.a_mode == 0) {
            tma_load_2d(&map_a, &full_bar[stage], sa, kb * bk, mt * BLOCK_M);
            if (!b_done) tma_load_2d(&map_b, &full_bar[stage], sb, kb * bk, n0);
          } else {
            const int b = mt / p.tiles_per_b, tb = mt - b * p.tiles_per_b;
            const int tap = kb / p.kc, j = kb - tap * p.kc;
            const int kh = tap / 3, kw = tap - kh * 3;
            tma_load_4d(&map_a, &full_bar[stage], sa, j * bk, kw - p.pad_f, 2 * tb * p.bt + kh - p.pad_t, b);
            if (!b_done) tma_load_2d(&map_b, &full_bar[stage], sb, tap * p.D + j * bk, n0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    constexpr uint32_t idesc = F16 ? make_idesc_f16(BLOCK_M, BLOCK_N) : make_idesc(BLOCK_M, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    const int kdim = (p.a_mode == 0) ? p.K : p.D;                       // columns covered by the slabs of one K run
    constexpr int umma_k = F16 ? 2 * UMMA_K : UMMA_K;                     // columns per instruction (32 bytes of a swizzle row either way)
    const int tail_cols = kdim - (kdim / bk) * bk;
    const int tail_ksteps = tail_cols == 0 ? BLOCK_K / UMMA_K : (tail_cols + umma_k - 1) / umma_k;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
      int slab_j = 0;                        // conv2: channel slab inside the current tap
      for (int kb = 0; kb < p.num_k_blocks; ++kb, slab_j = (slab_j + 1 == p.kc) ? 0 : slab_j + 1) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t da = make_smem_desc(sa);
          const uint64_t db = make_smem_desc(sa + kABytes);
          // k-steps that hold real columns: the last slab of a K = 144 GEMM (of every tap of conv2) has 16 columns = two k-steps, the
          // rest is TMA zero fill.  (No division here: this single thread paces the tensor pipe.)
          const bool last_slab = (p.a_mode == 0) ? (kb == p.num_k_blocks - 1) : (slab_j == p.kc - 1);
          const int ksteps = last_slab ? tail_ksteps : BLOCK_K / UMMA_K;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 32 bytes (8 tf32) inside the 128-byte swizzle row: +2 in the 16-byte address field
            if (k < ksteps) {
              if constexpr (F16) umma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
              else umma_tf32(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          tcgen05_commit(&empty_bar[stage]);                           // frees this smem stage when the MMAs retire
          if (kb == p.num_k_blocks - 1) tcgen05_commit(&tmem_full[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int quad = warp & 3;              // TMEM lane quadrant this warp may access
    // accumulator row held by this thread: M=128 fills all 128 lanes; M=64 uses lanes 0..15 of each quadrant
    // (cute tmem_frg: (16,4) x N with lane strides (1,32))
    float* wsm = reinterpret_cast<float*>(tmem_slot + 4) + (warp - 2) * kWsmFloats;   // warp-private transpose scratch
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int mt = tile / p.num_n_tiles, nt = tile - mt * p.num_n_tiles;
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      // rows of this warp: consecutive output rows starting at grow0; nrows of them are real
      size_t grow0;
      int nrows;
      constexpr int kWarpRows = (BLOCK_M == 128) ? 32 : 16;
      if (p.a_mode == 0) {
        const int m0 = mt * BLOCK_M + quad * kWarpRows;
        grow0 = (size_t)m0;
        nrows = min(kWarpRows, p.M - m0);
      } else {
        // conv2: tile row r = i * F2 + f2 <-> output row ((b*T2 + tb*bt + i) * F2 + f2): consecutive; rows beyond bt / T2 clipped
        const int b = mt / p.tiles_per_b, tb = mt - b * p.tiles_per_b;
        const int valid = min(p.bt, p.T2 - tb * p.bt) * p.F2;
        grow0 = ((size_t)b * p.T2 + (size_t)tb * p.bt) * p.F2 + quad * 32;
        nrows = min(32, valid - quad * 32);
      }
      if (nrows < 0) nrows = 0;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
      if constexpr (kIsLN) {
        if (kHasResidTile) mbar_wait(r_full, local & 1);
        const int trow = (BLOCK_M == 128) ? quad * 32 + lane : (lane < 16 ? quad * 16 + lane : -1);
        epilogue_ln_tma<EPI, BLOCK_N, BLOCK_M>(p, taddr, stile, &map_c, &map_c2, mt * BLOCK_M, trow, warp == 2 && lane == 0);
        __syncwarp();
        if (kHasResidTile && lane == 0) mbar_arrive(r_empty);
      } else if constexpr (EPI == EPI_BIAS_ARGMAX) {
        epilogue_argmax<BLOCK_N>(p, taddr, grow0, nrows, nt * BLOCK_N, nt, lane);
      } else {
        epilogue_plain<EPI, BLOCK_N>(p, taddr, wsm, grow0, nrows, nt * BLOCK_N, lane);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
template <int EPI, int BLOCK_N, int BLOCK_M>
constexpr size_t smem_bytes() {
  constexpr bool is_ln = (EPI == EPI_RESID_LN || EPI == EPI_RESID_LN2 || EPI == EPI_BIAS_LN);
  return (size_t)stages_for(is_ln, BLOCK_N, BLOCK_M) * (BLOCK_M * BLOCK_K * 4 + BLOCK_N * BLOCK_K * 4) +
         (is_ln ? (size_t)BLOCK_M * ((BLOCK_N + 31) / 32) * 128 : (size_t)4 * kWsmFloats * 4 /*epilogue transpose scratch*/) +
         1024 /*align slack*/ + 256 /*barriers*/;
}

template <int EPI, int BLOCK_N, int BLOCK_M, bool F16 = false>
int launch_one(TcContext& ctx, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap* lnmaps, const TcParams& tp,
               cudaStream_t stream) {
  static PerDeviceSmem configured;
  auto kern = gemm_tc_kernel<EPI, BLOCK_N, BLOCK_M, F16>;
  if (configured.need(smem_bytes<EPI, BLOCK_N, BLOCK_M>()))
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<EPI, BLOCK_N, BLOCK_M>()));
  const int tiles = tp.num_m_tiles * tp.num_n_tiles;
  const int grid = tiles < ctx.num_sms ? tiles : ctx.num_sms;
  B200_CUDA_OK(launch_k(kern, dim3(grid), dim3(kThreads), smem_bytes<EPI, BLOCK_N, BLOCK_M>(), stream, ma, mb, lnmaps[0], lnmaps[1], lnmaps[2], tp));
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int BLOCK_N, int BLOCK_M>
int dispatch_epi(TcContext& ctx, int epi, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap* lnmaps, const TcParams& tp,
                 cudaStream_t s) {
  switch (epi) {
    case EPI_BIAS: return launch_one<EPI_BIAS, BLOCK_N, BLOCK_M>(ctx, ma, mb, lnmaps, tp, s);
    case EPI_BIAS_RELU: return launch_one<EPI_BIAS_RELU, BLOCK_N, BLOCK_M>(ctx, ma, mb, lnmaps, tp, s);
    case EPI_BIAS_SWISH: return launch_one<EPI_BIAS_SWISH, BLOCK_N, BLOCK_M>(ctx, ma, mb, lnmaps, tp, s);
    case EPI_GLU: return launch_one<EPI_GLU, BLOCK_N, BLOCK_M>(ctx, ma, mb, lnmaps, tp, s);
    case EPI_RESID: return launch_one<EPI_RESID, BLOCK_N, BLOCK_M>(ctx, ma, mb, lnmaps, tp, s);
    case EPI_NONE: return launch_one<EPI_NONE, BLOCK_N, BLOCK_M>(ctx, ma, mb, lnmaps, tp, s);
    case EPI_RESID_LN: return launch_one<EPI_RESID_LN, BLOCK_N, BLOCK_M>(ctx, ma, mb, lnmaps, tp, s);
    case EPI_RESID_LN2: return launch_one<EPI_RESID_LN2, BLOCK_N, BLOCK_M>(ctx, ma, mb, lnmaps, tp, s);
    case EPI_BIAS_LN: return launch_one<EPI_BIAS_LN, BLOCK_N, BLOCK_M>(ctx, ma, mb, lnmaps, tp, s);
    case EPI_BIAS_ARGMAX:
      if constexpr (BLOCK_M == 128) return launch_one<EPI_BIAS_ARGMAX, BLOCK_N, BLOCK_M>(ctx, ma, mb, lnmaps, tp, s);
      break;
  }
  snprintf(g_errbuf, sizeof(g_errbuf), "gemm_tc: bad epilogue %d", epi);
  return 1;
}

int pick_block_n(int N, int epilogue, int M, int num_sms) {
  // QKV projection of ConformerS (N = 432, no epilogue math): two 224-column tiles per 128 rows make one wave of 126 CTAs on
  // 148 SMs, where three 144-column tiles would need two waves (an MMA costs the same ~132 cycles for N = 144 and N = 224)
  if (N == 432 && epilogue == EPI_NONE && ceil_div(M, 128) * 3 > num_sms && ceil_div(M, 128) * 2 <= num_sms) return 224;
  if (N % 144 == 0) return 144;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  if (N <= 144) return 144;
  if (N <= 192) return 192;
  if (N <= 256) return 256;
  if (N % 256 == 0) return 256;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  // ragged N (e.g. the 1332-class CTC head): fewest wasted columns among the instantiated widths
  int best = 256, waste = (N + 255) / 256 * 256 - N;
  const int cands[3] = {192, 144, 128};
  for (int c : cands) {
    const int w = (N + c - 1) / c * c - N;
    if (w < waste) { waste = w; best = c; }
  }
  return best;
}

}  // namespace

int tc_init(TcContext* ctx) {
  ctx->ready = false;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    snprintf(g_errbuf, sizeof(g_errbuf), "tc_init: cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
    return 1;
  }
  ctx->encode_tiled = fn;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, dev);
  ctx->ready = true;
  return 0;
}

int tc_argmax_tiles(int N) { return ceil_div(N, pick_block_n(N, EPI_BIAS_ARGMAX, 1, 148)); }

bool tc_gemm_supported(const GemmParams& p, int epilogue) {
  if (p.M <= 0 || p.N % 4 != 0 || p.K % 4 != 0) return false;
  if (p.f16) {   // fp16 operands: conv2 (a_mode 1, bias + ReLU) and the subsampling linear layer (a_mode 0, bias + LayerNorm), N = 144 | 256
    if (!(p.N == 144 || p.N == 256)) return false;
    if (p.a_mode == 1 ? epilogue != EPI_BIAS_RELU : (epilogue != EPI_BIAS_LN || p.lda % 8 != 0 || p.K % 8 != 0)) return false;
  }
  if (p.out_f16 && (epi_is_ln(epilogue) || epilogue == EPI_GLU || epilogue == EPI_BIAS_ARGMAX || epilogue == EPI_RESID || p.ldc % 4 != 0)) return false;
  if (epilogue == EPI_GLU && p.N % 8 != 0) return false;
  if (epilogue == EPI_BIAS_ARGMAX && (p.a_mode != 0 || p.bias == nullptr)) return false;
  if (epi_is_ln(epilogue)) {
    // fused LayerNorm: the row must be exactly one instantiated tile width, plain A operand, bias present
    if (p.a_mode != 0 || p.bias == nullptr || p.C2 == nullptr || p.ln1_g == nullptr || p.N != p.ldc) return false;
    if (!(p.N == 64 || p.N == 128 || p.N == 144 || p.N == 192 || p.N == 256)) return false;
  }
  if ((reinterpret_cast<uintptr_t>(p.A) | reinterpret_cast<uintptr_t>(p.W) | reinterpret_cast<uintptr_t>(p.C)) & 15) return false;
  if (p.a_mode == 0) return (p.lda % 4) == 0;
  if (p.f16 && (p.D % 8) != 0) return false;                 // 16-byte global strides of the fp16 map
  return p.F2 <= 64 && p.F2 >= 1 && (128 / p.F2) >= 1 && (p.D % 4) == 0 && p.N == p.D && 2 * p.F2 <= 256;
}

int launch_gemm_tc(TcContext& ctx, const GemmParams& p, int epilogue, cudaStream_t stream) {
  if (!ctx.ready) {
    snprintf(g_errbuf, sizeof(g_errbuf), "gemm_tc: tensor-map encoder not initialised");
    return 1;
  }
  const int bn = epi_is_ln(epilogue) ? p.N : pick_block_n(p.N, epilogue, p.M, ctx.num_sms);
  TcParams tp{};
  tp.bias = p.bias; tp.resid = p.resid; tp.C = p.C; tp.M = p.M; tp.N = p.N; tp.K = p.K; tp.ldc = p.ldc; tp.alpha = p.alpha;
  tp.ln1_g = p.ln1_g; tp.ln1_b = p.ln1_b; tp.ln2_g = p.ln2_g; tp.ln2_b = p.ln2_b; tp.C2 = p.C2; tp.ln_eps = p.ln_eps;
  tp.num_n_tiles = ceil_div(p.N, bn);
  tp.a_mode = p.a_mode;
  tp.round_out = p.round_out;
  tp.f16 = p.f16 ? 1 : 0;
  tp.out_f16 = p.out_f16 ? 1 : 0;
  const bool f16 = tp.f16 != 0;
  const cuuint64_t esz = f16 ? 2 : 4;                       // operand element size
  const cuuint32_t bk = f16 ? 2 * BLOCK_K : BLOCK_K;         // operand columns per 128-byte swizzle row
  CUtensorMap ma, mb;
  const cuuint32_t ones[4] = {1, 1, 1, 1};
  {
    const cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.N};
    const cuuint64_t strides[1] = {(cuuint64_t)p.K * esz};
    const cuuint32_t box[2] = {bk, (cuuint32_t)bn};
    if (encode_map(ctx, &mb, p.W, 2, dims, strides, box, ones, f16)) return 1;
  }
  // M tile: 128 rows normally; 64 when 128-row tiles would leave most SMs idle (the 8000-row, N<=256 GEMMs of one batch)
  int bm = 128;
  if (p.a_mode == 0 && bn != 224 && epilogue != EPI_BIAS_ARGMAX && ceil_div(p.M, 128) * tp.num_n_tiles < (ctx.num_sms * 3) / 4) bm = 64;
  if (p.a_mode == 0) {
    const cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.M};
    const cuuint64_t strides[1] = {(cuuint64_t)p.lda * esz};
    const cuuint32_t box[2] = {bk, (cuuint32_t)bm};
    if (encode_map(ctx, &ma, p.A, 2, dims, strides, box, ones, f16)) return 1;
    tp.num_m_tiles = ceil_div(p.M, bm);
    tp.num_k_blocks = ceil_div(p.K, (int)bk);
  } else {
    const int B = p.M / (p.T2 * p.F2);
    tp.T2 = p.T2; tp.F2 = p.F2; tp.D = p.D; tp.pad_t = p.pad_t; tp.pad_f = p.pad_f;
    tp.bt = 128 / p.F2;
    tp.kc = ceil_div(p.D, (int)bk);
    tp.tiles_per_b = ceil_div(p.T2, tp.bt);
    tp.num_m_tiles = B * tp.tiles_per_b;
    tp.num_k_blocks = 9 * tp.kc;
    const cuuint64_t dims[4] = {(cuuint64_t)p.D, (cuuint64_t)p.F1, (cuuint64_t)p.T1, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)p.D * esz, (cuuint64_t)p.F1 * p.D * esz, (cuuint64_t)p.T1 * p.F1 * p.D * esz};
    const cuuint32_t box[4] = {bk, (cuuint32_t)(2 * p.F2), (cuuint32_t)(2 * tp.bt), 1};
    const cuuint32_t estr[4] = {1, 2, 2, 1};
    if (encode_map(ctx, &ma, p.A, 4, dims, strides, box, estr, f16)) return 1;
  }
  // LayerNorm epilogues: residual-in / output tiles travel by TMA ([M, N] in BLOCK_M x 32-column slabs; stores clip tails)
  CUtensorMap lnmaps[3];
  memset(lnmaps, 0, sizeof(lnmaps));
  if (epi_is_ln(epilogue)) {
    const cuuint64_t dims[2] = {(cuuint64_t)p.N, (cuuint64_t)p.M};
    const cuuint64_t strides[1] = {(cuuint64_t)p.ldc * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)bm};
    if (p.resid && encode_map(ctx, &lnmaps[0], p.resid, 2, dims, strides, box, ones)) return 1;
    if (encode_map(ctx, &lnmaps[1], p.C, 2, dims, strides, box, ones)) return 1;
    if (encode_map(ctx, &lnmaps[2], p.C2, 2, dims, strides, box, ones)) return 1;
  }
  if (f16) {   // instantiated for the two GEMMs of the subsampler only (tc_gemm_supported admits nothing else)
    if (epilogue == EPI_BIAS_RELU && bn == 144) return launch_one<EPI_BIAS_RELU, 144, 128, true>(ctx, ma, mb, lnmaps, tp, stream);
    if (epilogue == EPI_BIAS_RELU && bn == 256) return launch_one<EPI_BIAS_RELU, 256, 128, true>(ctx, ma, mb, lnmaps, tp, stream);
    if (epilogue == EPI_BIAS_LN && bn == 144 && bm == 64) return launch_one<EPI_BIAS_LN, 144, 64, true>(ctx, ma, mb, lnmaps, tp, stream);
    if (epilogue == EPI_BIAS_LN && bn == 144 && bm == 128) return launch_one<EPI_BIAS_LN, 144, 128, true>(ctx, ma, mb, lnmaps, tp, stream);
    if (epilogue == EPI_BIAS_LN && bn == 256 && bm == 64) return launch_one<EPI_BIAS_LN, 256, 64, true>(ctx, ma, mb, lnmaps, tp, stream);
    if (epilogue == EPI_BIAS_LN && bn == 256 && bm == 128) return launch_one<EPI_BIAS_LN, 256, 128, true>(ctx, ma, mb, lnmaps, tp, stream);
    snprintf(g_errbuf, sizeof(g_errbuf), "gemm_tc: no fp16-operand kernel for epilogue %d, tile %d x %d", epilogue, bm, bn);
    return 1;
  }
  if (bn == 224) return launch_one<EPI_NONE, 224, 128>(ctx, ma, mb, lnmaps, tp, stream);
  if (bm == 64) {
    switch (bn) {
      case 64: return dispatch_epi<64, 64>(ctx, epilogue, ma, mb, lnmaps, tp, stream);
      case 128: return dispatch_epi<128, 64>(ctx, epilogue, ma, mb, lnmaps, tp, stream);
      case 144: return dispatch_epi<144, 64>(ctx, epilogue, ma, mb, lnmaps, tp, stream);
      case 192: return dispatch_epi<192, 64>(ctx, epilogue, ma, mb, lnmaps, tp, stream);
      default: return dispatch_epi<256, 64>(ctx, epilogue, ma, mb, lnmaps, tp, stream);
    }
  }
  switch (bn) {
    case 64: return dispatch_epi<64, 128>(ctx, epilogue, ma, mb, lnmaps, tp, stream);
    case 128: return dispatch_epi<128, 128>(ctx, epilogue, ma, mb, lnmaps, tp, stream);
    case 144: return dispatch_epi<144, 128>(ctx, epilogue, ma, mb, lnmaps, tp, stream);
    case 192: return dispatch_epi<192, 128>(ctx, epilogue, ma, mb, lnmaps, tp, stream);
    default: return dispatch_epi<256, 128>(ctx, epilogue, ma, mb, lnmaps, tp, stream);
  }
}

}  // namespace b200asr
