// ConvSubsampling (conformer_blocks.py:67-96) as ONE tcgen05 kernel: conv1 (3x3, stride 2, 'same', 1 -> D, ReLU) is computed on
// the fly by producer warps straight into the swizzled shared-memory A tile of conv2's implicit GEMM (3x3, stride 2, 'same',
// D -> D, +bias, ReLU), so conv1's [B, T/2, 40, D] map (369 MB at 32 x 10 s) never exists in HBM.
//
//   GEMM view: rows = (b, t2, f2) output positions, N = D output channels, K = (kh, kw, ci) = 9 D.  One tile = `bt` consecutive
//   (b, t2) time rows x F2 frequency bins (bt = 128 / F2 = 6 -> 120 of the 128 MMA rows; time rows are taken from the flattened
//   (b, t2) axis, so tiles may straddle utterances and no tile is ragged except the last).
//   K blocks are ordered channel-slab major: kb = j * 9 + tap (slab j = input channels [32 j, 32 j + 32), tap = kh * 3 + kw), so
//   a producer thread keeps its conv1 weights (8 channels x 9 taps, packed fp32x2) in registers across 9 consecutive K blocks.
//
//   warp 0        TMA producer of the B operand (conv2 weights [D, 9 D], K-major): box {32, D} at K offset tap * D + 32 j
//   warp 1        MMA issuer + TMEM owner (two accumulators: the epilogue of tile i overlaps the MMAs of tile i + 1)
//   warps 2..5    epilogue: tcgen05.ld, + bias, ReLU, round to tf32 (the output only feeds the subsampling linear GEMM), coalesced stores
//   warps 6..13   A producers: per tile they stage the mel patch of every time row ([bt][7 mel rows][4 F2 + 4 columns], as (m, m)
//                 pairs for packed FMAs); per K block each thread computes 2 x 8 conv1 outputs (9 taps each, bias, ReLU, zero outside
//                 conv1's extent = conv2's 'same' padding), rounds them to nearest tf32 and writes two 16-byte chunks per row into
//                 the SWIZZLE_128B K-major A tile of the pipeline stage; fence.proxy.async + one mbarrier arrive per warp.
// Cost model per tile (D = 144): 162 MMAs; B operand 45 x 18 KB through TMA (55 B/clk per SM => ~15 k cycles); producers ~61 k warp
// instructions (=> ~15 k cycles on 4 schedulers); the conv1 recomputation factor is 2.25 (each conv1 output feeds up to 4 taps).
#include "tc_common.cuh"

namespace b200asr {

using namespace tc;

namespace {

constexpr int kCsThreads = 448;          // 14 warps
constexpr int kProdWarps = 8;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kProdTid0 = 6 * 32;        // first producer thread
constexpr int kPatchRows = 7;            // mel rows one (b, t2) output row touches: 2 kh + a, kh, a in 0..2

struct ConvSubKParams {
  TcParams ep;                 // bias = conv2 bias, C = out [rows, D], M = rows, N = D, ldc = D, round_out
  const float* mel;            // [B, T, F]
  const float* w1;             // [9, D] tap-major
  const float* b1;             // [D]
  int B, T, F, T1, F1, T2, F2, D;
  int pt1, pf1, pt2, pf2;
  int bt, num_tiles, kc, num_kb, PW;   // PW = 4 F2 + 4 patch columns
};

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  return (unsigned long long)__float_as_uint(a) | ((unsigned long long)__float_as_uint(b) << 32);
}
__device__ __forceinline__ unsigned long long cs_ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
  return c;
}
// relu + round to nearest tf32 of both halves of a packed pair
__device__ __forceinline__ float2 relu_rn2(unsigned long long v) {
  const float lo = fmaxf(__uint_as_float((uint32_t)v), 0.f), hi = fmaxf(__uint_as_float((uint32_t)(v >> 32)), 0.f);
  return make_float2(__uint_as_float(tf32_rn_bits(lo)), __uint_as_float(tf32_rn_bits(hi)));
}

// explicit shared-window accesses: pointers carved out of the dynamic shared-memory block with run-time offsets lose their address
// space, and the compiler then emits generic LD / ST (measured: the first version of this kernel ran 3x slower than planned)
__device__ __forceinline__ unsigned long long lds_u64(uint32_t saddr) {
  unsigned long long v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_v2(uint32_t saddr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(saddr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}

template <int BLOCK_N>
__host__ __device__ constexpr int cs_stages() { return BLOCK_N <= 160 ? 4 : 3; }

template <int BLOCK_N>
__global__ void __launch_bounds__(kCsThreads, 1)
conv_subsample_tc_kernel(const __grid_constant__ CUtensorMap map_b, const ConvSubKParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t kABytes = 128 * BLOCK_K * 4;          // 16 KB
  constexpr uint32_t kBBytes = BLOCK_N * BLOCK_K * 4;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  constexpr int kStg = cs_stages<BLOCK_N>();
  // 1024-byte alignment by POINTER arithmetic on the shared array (an integer round trip would strip the address space and turn every
  // access through a derived pointer into a generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  float* wsm = reinterpret_cast<float*>(smem + kStg * kStageBytes);                  // epilogue transpose scratch (4 warps)
  float* w1s = wsm + 4 * kWsmFloats;                                                 // conv1 weights [9][D] + bias [D]
  float2* patch = reinterpret_cast<float2*>(w1s + 10 * p.D);                         // [bt][7][PW] of (m, m)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(patch) + (size_t)p.bt * kPatchRows * p.PW * 8);
  uint64_t* empty_bar = full_bar + kStg;
  uint64_t* tmem_full = empty_bar + kStg;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_tile = p.bt * p.F2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < kStg; ++s) {
      mbar_init(&full_bar[s], 1 + kProdWarps);     // TMA's arrive.expect_tx + one arrive per producer warp
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // conv1 weights + bias are constants: staged before griddepcontrol.wait
  for (int i = threadIdx.x; i < 10 * p.D; i += kCsThreads) w1s[i] = (i < 9 * p.D) ? p.w1[i] : p.b1[i - 9 * p.D];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();

  if (warp == 0) {
    // ===================================================================== TMA producer (conv2 weights)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int j = 0; j < p.kc; ++j)
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], kBBytes);
            tma_load_2d(&map_b, &full_bar[stage], smem + stage * kStageBytes + kABytes, tap * p.D + j * BLOCK_K, 0);
            if (++stage == kStg) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    constexpr uint32_t idesc = make_idesc(128, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int j = 0, kb = 0; j < p.kc; ++j) {
        const int ksteps = min(BLOCK_K / UMMA_K, (p.D - j * BLOCK_K + UMMA_K - 1) / UMMA_K);   // the last slab may hold < 32 channels
        for (int tap = 0; tap < 9; ++tap, ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t da = make_smem_desc(sa);
          const uint64_t db = make_smem_desc(sa + kABytes);
          for (int k = 0; k < ksteps; ++k) umma_tf32(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          tcgen05_commit(&empty_bar[stage]);
          if (kb == p.num_kb - 1) tcgen05_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == kStg) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < 6) {
    // ===================================================================== epilogue (warps 2..5)
    const int quad = warp & 3;
    float* wsm_w = wsm + (warp - 2) * kWsmFloats;
    int local = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const long long row0 = (long long)tile * rows_per_tile;
      const int valid = (int)min((long long)rows_per_tile, (long long)p.ep.M - row0);
      int nrows = min(32, valid - quad * 32);
      if (nrows < 0) nrows = 0;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
      epilogue_plain<EPI_BIAS_RELU, BLOCK_N>(p.ep, taddr, wsm_w, (size_t)(row0 + quad * 32), nrows, 0, lane);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  } else {
    // ===================================================================== A producers (warps 6..13)
    const int pt = threadIdx.x - kProdTid0;       // 0..255
    const int cg = pt & 3;                        // 8-channel group inside the 32-channel slab
    const int rsub = pt >> 2;                     // 0..63: row inside a pass
    const int PW = p.PW;
    const int col_shift = 2 * p.pf2 + p.pf1;      // patch column pc <-> mel column pc - col_shift
    const int row_shift = 2 * p.pt2 + p.pt1;      // patch row pr of time row (b, t2) <-> mel row 4 t2 - row_shift + pr
    const int total_trows = p.B * p.T2;
    const uint32_t patch_s = smem_u32(patch), w1s_s = smem_u32(w1s), smem_s = smem_u32(smem);
    pdl_wait();                                   // mel (written by the frontend kernel) is complete and visible
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      // ---- stage the mel patch of this tile's time rows: one warp per (time row, patch row), lanes over the columns (no divisions,
      // independent loads)
      asm volatile("bar.sync 3, %0;" ::"n"(kProdThreads) : "memory");        // every producer is done with the previous patch
      const int per_row = kPatchRows * PW;
      const int pwarp = pt >> 5;
      for (int rr = pwarp; rr < p.bt * kPatchRows; rr += kProdWarps) {
        const int i = rr / kPatchRows, pr = rr - i * kPatchRows;
        const int g = tile * p.bt + i;
        const bool gv = g < total_trows;
        const int b = gv ? g / p.T2 : 0, t2 = gv ? g - b * p.T2 : 0;
        const int t = 4 * t2 - row_shift + pr;
        const bool tv = gv && t >= 0 && t < p.T;
        const float* mrow = p.mel + ((size_t)b * p.T + (tv ? t : 0)) * p.F;
        for (int pc = lane; pc < PW; pc += 32) {
          const int f = pc - col_shift;
          const float v = (tv && f >= 0 && f < p.F) ? __ldg(mrow + f) : 0.f;
          sts_v2(patch_s + (uint32_t)(rr * PW + pc) * 8u, v, v);
        }
      }
      asm volatile("bar.sync 3, %0;" ::"n"(kProdThreads) : "memory");
      // ---- per-pass row geometry (fixed for the tile): patch byte address of (kh = 0, a = 0, kw = 0, bc = 0), A-tile byte offsets of the
      // two 16-byte chunks, and one validity bit per conv2 tap (the row exists, and the conv1 position the tap reads lies inside
      // conv1's output: outside it conv2 sees its own zero padding, not relu(bias))
      uint32_t pb8[2], soff0[2], soff1[2];
      uint32_t okmask[2];
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        const int r = ps * 64 + rsub;
        const int i = r / p.F2, f2 = r - i * p.F2;
        const int g = tile * p.bt + i;
        const bool rv = (r < rows_per_tile) && (g < total_trows);
        const int t2 = rv ? g % p.T2 : 0;
        pb8[ps] = patch_s + (uint32_t)((rv ? i : 0) * per_row + (rv ? 4 * f2 : 0)) * 8u;
        soff0[ps] = (uint32_t)r * 128u + (uint32_t)((((2 * cg) ^ (r & 7))) << 4);
        soff1[ps] = (uint32_t)r * 128u + (uint32_t)((((2 * cg + 1) ^ (r & 7))) << 4);
        uint32_t m = 0;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int t1 = 2 * t2 + kh - p.pt2, f1 = 2 * f2 + kw - p.pf2;
            if (rv && t1 >= 0 && t1 < p.T1 && f1 >= 0 && f1 < p.F1) m |= 1u << (kh * 3 + kw);
          }
        okmask[ps] = m;
      }
      const uint32_t PW8 = (uint32_t)PW * 8u;
      for (int j = 0; j < p.kc; ++j) {
        // conv1 weights of this thread's 8 channels for slab j: w[tap1][4 pairs], bias[4 pairs]
        const int c0 = j * BLOCK_K + 8 * cg;
        const bool cvalid = c0 < p.D;             // D % 8 == 0: the 8 channels are valid together
        unsigned long long w[9][4], bias[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {      // (c0 is even and w1s is 8-byte aligned: one 64-bit shared load per pair)
          bias[q] = cvalid ? lds_u64(w1s_s + (uint32_t)(9 * p.D + c0 + 2 * q) * 4u) : 0ull;
#pragma unroll
          for (int t = 0; t < 9; ++t) w[t][q] = cvalid ? lds_u64(w1s_s + (uint32_t)(t * p.D + c0 + 2 * q) * 4u) : 0ull;
        }
        // one conv2 tap = one K block.  kh / kw are compile-time constants (no index arithmetic in the body); the nine patch loads of a
        // pass are issued before the first FMA (they are independent); invalid positions are computed like valid ones and zeroed by a
        // select, so the body is branch-free.
        auto do_tap = [&](auto KH, auto KW) {
          constexpr int kh = decltype(KH)::value, kw = decltype(KW)::value;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t sa = smem_s + (uint32_t)stage * kStageBytes;
#pragma unroll
          for (int ps = 0; ps < 2; ++ps) {
            const uint32_t pp = pb8[ps] + (uint32_t)(2 * kh) * PW8 + (uint32_t)(2 * kw) * 8u;
            unsigned long long m[9];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
              for (int bc = 0; bc < 3; ++bc) m[a * 3 + bc] = lds_u64(pp + (uint32_t)a * PW8 + (uint32_t)bc * 8u);
            unsigned long long a0 = bias[0], a1 = bias[1], a2 = bias[2], a3 = bias[3];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              a0 = cs_ffma2(m[t], w[t][0], a0);
              a1 = cs_ffma2(m[t], w[t][1], a1);
              a2 = cs_ffma2(m[t], w[t][2], a2);
              a3 = cs_ffma2(m[t], w[t][3], a3);
            }
            const bool ok = cvalid && ((okmask[ps] >> (kh * 3 + kw)) & 1u);
            const float2 r0 = relu_rn2(a0), r1 = relu_rn2(a1), r2 = relu_rn2(a2), r3 = relu_rn2(a3);
            const float4 o0 = ok ? make_float4(r0.x, r0.y, r1.x, r1.y) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 o1 = ok ? make_float4(r2.x, r2.y, r3.x, r3.y) : make_float4(0.f, 0.f, 0.f, 0.f);
            // K-major SWIZZLE_128B: 16-byte chunk q of row r lives at r * 128 + ((q ^ (r & 7)) << 4)
            sts_v4(sa + soff0[ps], o0);
            sts_v4(sa + soff1[ps], o1);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_bar[stage]);
          if (++stage == kStg) { stage = 0; phase ^= 1; }
        };
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        using I2 = std::integral_constant<int, 2>;
        do_tap(I0{}, I0{}); do_tap(I0{}, I1{}); do_tap(I0{}, I2{});
        do_tap(I1{}, I0{}); do_tap(I1{}, I1{}); do_tap(I1{}, I2{});
        do_tap(I2{}, I0{}); do_tap(I2{}, I1{}); do_tap(I2{}, I2{});
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

template <int BLOCK_N>
size_t cs_smem_bytes(const ConvSubKParams& kp) {
  return (size_t)cs_stages<BLOCK_N>() * (128 * BLOCK_K * 4 + BLOCK_N * BLOCK_K * 4) + (size_t)4 * kWsmFloats * 4 + (size_t)10 * kp.D * 4 +
         (size_t)kp.bt * kPatchRows * kp.PW * 8 + 256 + 1024;
}

template <int BLOCK_N>
int launch_cs(TcContext& ctx, const CUtensorMap& mb, const ConvSubKParams& kp, cudaStream_t stream) {
  static PerDeviceSmem configured;
  auto kern = conv_subsample_tc_kernel<BLOCK_N>;
  const size_t smem = cs_smem_bytes<BLOCK_N>(kp);
  if (smem > 227 * 1024) {
    snprintf(g_errbuf, sizeof(g_errbuf), "conv_subsample_tc: %zu bytes of shared memory needed", smem);
    return 1;
  }
  if (configured.need(smem)) B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = kp.num_tiles < ctx.num_sms ? kp.num_tiles : ctx.num_sms;
  B200_CUDA_OK(launch_k(kern, dim3(grid), dim3(kCsThreads), smem, stream, mb, kp));
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

bool conv_subsample_tc_supported(const ConvSubParams& p) {
  if (p.D != 144 && p.D != 256) return false;                       // instantiated tile widths (D % 8 == 0)
  if (p.F2 < 1 || p.F2 > 64 || p.B <= 0 || p.T <= 0) return false;
  if ((reinterpret_cast<uintptr_t>(p.w2) | reinterpret_cast<uintptr_t>(p.out)) & 15) return false;
  return true;
}

int launch_conv_subsample_tc(TcContext& ctx, const ConvSubParams& p, cudaStream_t stream) {
  if (!ctx.ready) {
    snprintf(g_errbuf, sizeof(g_errbuf), "conv_subsample_tc: tensor-map encoder not initialised");
    return 1;
  }
  ConvSubKParams kp{};
  kp.mel = p.mel; kp.w1 = p.w1; kp.b1 = p.b1;
  kp.B = p.B; kp.T = p.T; kp.F = p.F; kp.T1 = p.T1; kp.F1 = p.F1; kp.T2 = p.T2; kp.F2 = p.F2; kp.D = p.D;
  kp.pt1 = p.pt1; kp.pf1 = p.pf1; kp.pt2 = p.pt2; kp.pf2 = p.pf2;
  kp.bt = 128 / p.F2;
  const long long trows = (long long)p.B * p.T2;
  kp.num_tiles = (int)((trows + kp.bt - 1) / kp.bt);
  kp.kc = ceil_div(p.D, BLOCK_K);
  kp.num_kb = 9 * kp.kc;
  kp.PW = 4 * p.F2 + 4;
  kp.ep.bias = p.b2; kp.ep.C = p.out; kp.ep.M = (int)(trows * p.F2); kp.ep.N = p.D; kp.ep.K = 9 * p.D; kp.ep.ldc = p.D;
  kp.ep.round_out = p.round_out; kp.ep.num_n_tiles = 1;
  CUtensorMap mb;
  const cuuint32_t ones[2] = {1, 1};
  const cuuint64_t dims[2] = {(cuuint64_t)(9 * p.D), (cuuint64_t)p.D};
  const cuuint64_t strides[1] = {(cuuint64_t)(9 * p.D) * 4};
  const cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)p.D};
  if (encode_map(ctx, &mb, p.w2, 2, dims, strides, box, ones)) return 1;
  if (p.D == 144) return launch_cs<144>(ctx, mb, kp, stream);
  return launch_cs<256>(ctx, mb, kp, stream);
}

}  // namespace b200asr
