// Exact-fp32 tiled GEMM (CUDA cores) with fused epilogues + the first subsampling conv.
//
// This is the "precise" arithmetic mode of the engine (b200asr_config.precision = 1) and the checker the tcgen05
// tf32 kernels (gemm_tc.cu) are validated against on the GPU.  C[M,N] = A[M,K] . W[N,K]^T, both K-major.
// The A operand may be gathered on the fly as the im2col view of the second subsampling conv
// (conformer_blocks.py:81-85: 3x3, stride 2, 'same'), so conv1's activations are read in place.
#include "kernels.cuh"
#include <cuda_fp16.h>

#include <cstdlib>

namespace b200asr {

namespace {

constexpr int BM = 128, BN = 64, BK = 16;

struct ARow {
  const float* base;  // pointer to k = 0 of this row for plain; unused for conv2
  int b, t2, f2;
  bool valid;
};

template <int AMODE>
__device__ __forceinline__ float4 load_a(const GemmParams& p, int m, int k) {
  if (m >= p.M) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (AMODE == 0) {
    return *reinterpret_cast<const float4*>(p.A + (size_t)m * p.lda + k);
  } else {
    const int f2 = m % p.F2;
    const int r = m / p.F2;
    const int t2 = r % p.T2;
    const int b = r / p.T2;
    const int tap = k / p.D;
    const int c = k - tap * p.D;
    const int kh = tap / 3, kw = tap - kh * 3;
    const int y = 2 * t2 + kh - p.pad_t;
    const int x = 2 * f2 + kw - p.pad_f;
    if (y < 0 || y >= p.T1 || x < 0 || x >= p.F1) return make_float4(0.f, 0.f, 0.f, 0.f);
    return *reinterpret_cast<const float4*>(p.A + (((size_t)b * p.T1 + y) * p.F1 + x) * p.D + c);
  }
}

template <int AMODE, int EPI>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmParams p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  const int a_row = tid >> 2, a_kq = (tid & 3) * 4;
  const int b_row = tid >> 2, b_kq = (tid & 3) * 4;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 ra0, ra1, rb;
  auto gload = [&](int k0) {
    ra0 = load_a<AMODE>(p, m0 + a_row, k0 + a_kq);
    ra1 = load_a<AMODE>(p, m0 + a_row + 64, k0 + a_kq);
    const int n = n0 + b_row;
    rb = (n < p.N) ? *reinterpret_cast<const float4*>(p.W + (size_t)n * p.K + k0 + b_kq) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  gload(0);
  for (int k0 = 0; k0 < p.K; k0 += BK) {
    As[a_kq + 0][a_row] = ra0.x; As[a_kq + 1][a_row] = ra0.y; As[a_kq + 2][a_row] = ra0.z; As[a_kq + 3][a_row] = ra0.w;
    As[a_kq + 0][a_row + 64] = ra1.x; As[a_kq + 1][a_row + 64] = ra1.y; As[a_kq + 2][a_row + 64] = ra1.z; As[a_kq + 3][a_row + 64] = ra1.w;
    Bs[b_kq + 0][b_row] = rb.x; Bs[b_kq + 1][b_row] = rb.y; Bs[b_kq + 2][b_row] = rb.z; Bs[b_kq + 3][b_row] = rb.w;
    __syncthreads();
    if (k0 + BK < p.K) gload(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 bb = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int n = n0 + tx * 4;
  if (n >= p.N) return;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (EPI != EPI_NONE && p.bias != nullptr) {
    const float4 bq = *reinterpret_cast<const float4*>(p.bias + n);
    bias[0] = bq.x; bias[1] = bq.y; bias[2] = bq.z; bias[3] = bq.w;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
    if (EPI == EPI_GLU) {
      float2 o;
      o.x = v[0] * sigmoidf_(v[1]);
      o.y = v[2] * sigmoidf_(v[3]);
      *reinterpret_cast<float2*>(p.C + (size_t)m * p.ldc + (n >> 1)) = o;
    } else {
      if (EPI == EPI_BIAS_RELU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
      } else if (EPI == EPI_BIAS_SWISH) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = swishf_(v[j]);
      } else if (EPI == EPI_RESID) {
        const float4 r = *reinterpret_cast<const float4*>(p.resid + (size_t)m * p.ldc + n);
        v[0] = r.x + p.alpha * v[0]; v[1] = r.y + p.alpha * v[1]; v[2] = r.z + p.alpha * v[2]; v[3] = r.w + p.alpha * v[3];
      }
      *reinterpret_cast<float4*>(p.C + (size_t)m * p.ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

template <int AMODE>
int dispatch(const GemmParams& p, int epi, dim3 grid, cudaStream_t s) {
  switch (epi) {
    case EPI_BIAS: B200_CUDA_OK(launch_k(gemm_simt_kernel<AMODE, EPI_BIAS>, grid, dim3(256), 0, s, p)); break;
    case EPI_BIAS_RELU: B200_CUDA_OK(launch_k(gemm_simt_kernel<AMODE, EPI_BIAS_RELU>, grid, dim3(256), 0, s, p)); break;
    case EPI_BIAS_SWISH: B200_CUDA_OK(launch_k(gemm_simt_kernel<AMODE, EPI_BIAS_SWISH>, grid, dim3(256), 0, s, p)); break;
    case EPI_GLU: B200_CUDA_OK(launch_k(gemm_simt_kernel<AMODE, EPI_GLU>, grid, dim3(256), 0, s, p)); break;
    case EPI_RESID: B200_CUDA_OK(launch_k(gemm_simt_kernel<AMODE, EPI_RESID>, grid, dim3(256), 0, s, p)); break;
    case EPI_NONE: B200_CUDA_OK(launch_k(gemm_simt_kernel<AMODE, EPI_NONE>, grid, dim3(256), 0, s, p)); break;
    default: snprintf(g_errbuf, sizeof(g_errbuf), "gemm_simt: bad epilogue %d", epi); return 1;
  }
  return 0;
}

// conv1: 3x3 stride (2,2) 'same' on a single input channel + ReLU (conformer_blocks.py:76-80).  HBM-write bound
// (B*T1*F1*D floats out, tiny input): one CTA per (b, group of ROWS output time rows); a thread owns 4 consecutive output
// channels (its 9x4 weights live in registers for the whole CTA) and walks over frequency; the three mel rows an output
// row needs are staged in shared memory; stores are 16-byte, coalesced along channels.
constexpr int kConv1Rows = 4;
__global__ void __launch_bounds__(256) conv1_kernel(const Conv1Params p, int groups, int flanes) {
  extern __shared__ float mel_s[];  // [2*ROWS+1][F + 2] with one zero column of padding on each side
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int t1_0 = blockIdx.x * kConv1Rows;
  const int FW = p.F + 2;
  const int nrows = 2 * kConv1Rows + 1;
  for (int i = threadIdx.x; i < nrows * FW; i += blockDim.x) {
    const int r = i / FW, xx = i - r * FW;
    const int y = 2 * t1_0 + r - p.pad_t;
    const int x = xx - 1;
    float v = 0.f;
    if (y >= 0 && y < p.T && x >= 0 && x < p.F) v = p.mel[((size_t)b * p.T + y) * p.F + x];
    mel_s[i] = v;
  }
  const int g = threadIdx.x % groups, fl = threadIdx.x / groups;
  float4 w[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) w[k] = *reinterpret_cast<const float4*>(p.w + k * p.D + 4 * g);
  const float4 bias = *reinterpret_cast<const float4*>(p.bias + 4 * g);
  __syncthreads();
  if (fl >= flanes) return;
  for (int r = 0; r < kConv1Rows; ++r) {
    const int t1 = t1_0 + r;
    if (t1 >= p.T1) break;
    float* orow = p.out + (((size_t)b * p.T1 + t1) * p.F1) * p.D + 4 * g;
    for (int f1 = fl; f1 < p.F1; f1 += flanes) {
      float4 acc = bias;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const float* mrow = mel_s + (2 * r + kh) * FW + (2 * f1 - p.pad_f + 1);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float m = mrow[kw];
          const float4 ww = w[kh * 3 + kw];
          acc.x = fmaf(m, ww.x, acc.x); acc.y = fmaf(m, ww.y, acc.y); acc.z = fmaf(m, ww.z, acc.z); acc.w = fmaf(m, ww.w, acc.w);
        }
      }
      acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
      if (p.round_tf32) { acc.x = tf32_rn(acc.x); acc.y = tf32_rn(acc.y); acc.z = tf32_rn(acc.z); acc.w = tf32_rn(acc.w); }
      *reinterpret_cast<float4*>(orow + (size_t)f1 * p.D) = acc;
    }
  }
}

// (An eight-channels-per-thread variant of the fp16-output path -- nine shared loads feeding 36 packed FMAs instead of 18 -- was
// measured SLOWER: 98.7 us against 76.7 us; the kernel needs the parallelism of four channels per thread more than the saved loads.)
// Second generation of conv1: same mapping, but the mel patch is staged as (m, m) pairs and the four channels of a thread are
// two packed fp32x2 accumulators, so a tap costs one 8-byte shared load + two fma.rn.f32x2 instead of one load + four FFMA
// (the kernel is issue-bound: 68 M warp instructions for 369 MB of output).  Bit-identical results: f32x2 is two independent
// round-to-nearest FMAs.
__device__ __forceinline__ unsigned long long ffma2_u64(unsigned long long a, unsigned long long b, unsigned long long c) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
  return c;
}
__global__ void __launch_bounds__(256) conv1_f32x2_kernel(const Conv1Params p, int groups, int flanes) {
  extern __shared__ __align__(8) float2 mel2_s[];  // [2*ROWS+1][F + 2] of (m, m), one zero column of padding on each side
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int t1_0 = blockIdx.x * kConv1Rows;
  const int FW = p.F + 2;
  const int nrows = 2 * kConv1Rows + 1;
  for (int i = threadIdx.x; i < nrows * FW; i += blockDim.x) {
    const int r = i / FW, xx = i - r * FW;
    const int y = 2 * t1_0 + r - p.pad_t;
    const int x = xx - 1;
    float v = 0.f;
    if (y >= 0 && y < p.T && x >= 0 && x < p.F) v = p.mel[((size_t)b * p.T + y) * p.F + x];
    mel2_s[i] = make_float2(v, v);
  }
  const int g = threadIdx.x % groups, fl = threadIdx.x / groups;
  unsigned long long wlo[9], whi[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float4 w = *reinterpret_cast<const float4*>(p.w + k * p.D + 4 * g);
    float2 lo = make_float2(w.x, w.y), hi = make_float2(w.z, w.w);
    wlo[k] = *reinterpret_cast<unsigned long long*>(&lo);
    whi[k] = *reinterpret_cast<unsigned long long*>(&hi);
  }
  const float4 bias = *reinterpret_cast<const float4*>(p.bias + 4 * g);
  float2 blo = make_float2(bias.x, bias.y), bhi = make_float2(bias.z, bias.w);
  const unsigned long long b_lo = *reinterpret_cast<unsigned long long*>(&blo), b_hi = *reinterpret_cast<unsigned long long*>(&bhi);
  __syncthreads();
  if (fl >= flanes) return;
  for (int r = 0; r < kConv1Rows; ++r) {
    const int t1 = t1_0 + r;
    if (t1 >= p.T1) break;
    float* orow = p.out + (((size_t)b * p.T1 + t1) * p.F1) * p.D + 4 * g;
    for (int f1 = fl; f1 < p.F1; f1 += flanes) {
      unsigned long long a_lo = b_lo, a_hi = b_hi;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const unsigned long long* mrow = reinterpret_cast<const unsigned long long*>(mel2_s + (2 * r + kh) * FW + (2 * f1 - p.pad_f + 1));
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const unsigned long long m = mrow[kw];
          a_lo = ffma2_u64(m, wlo[kh * 3 + kw], a_lo);
          a_hi = ffma2_u64(m, whi[kh * 3 + kw], a_hi);
        }
      }
      const float2 lo = *reinterpret_cast<float2*>(&a_lo), hi = *reinterpret_cast<float2*>(&a_hi);
      float4 o4 = make_float4(fmaxf(lo.x, 0.f), fmaxf(lo.y, 0.f), fmaxf(hi.x, 0.f), fmaxf(hi.y, 0.f));
      if (p.out_f16) {            // 4 halves = 8 bytes per thread; the row pointer counts halves here
        __half2 h01 = __floats2half2_rn(o4.x, o4.y), h23 = __floats2half2_rn(o4.z, o4.w);
        uint2 pk;
        pk.x = *reinterpret_cast<unsigned int*>(&h01);
        pk.y = *reinterpret_cast<unsigned int*>(&h23);
        __half* hrow = reinterpret_cast<__half*>(p.out) + (((size_t)b * p.T1 + t1) * p.F1) * p.D + 4 * g;
        *reinterpret_cast<uint2*>(hrow + (size_t)f1 * p.D) = pk;
        continue;
      }
      if (p.round_tf32) { o4.x = tf32_rn(o4.x); o4.y = tf32_rn(o4.y); o4.z = tf32_rn(o4.z); o4.w = tf32_rn(o4.w); }
      *reinterpret_cast<float4*>(orow + (size_t)f1 * p.D) = o4;
    }
  }
}

}  // namespace

int launch_gemm_simt(const GemmParams& p, int epilogue, cudaStream_t stream) {
  if (p.K % BK != 0 || p.N % 4 != 0 || (p.a_mode == 1 && p.D % BK != 0)) {
    snprintf(g_errbuf, sizeof(g_errbuf), "gemm_simt: unsupported shape M=%d N=%d K=%d", p.M, p.N, p.K);
    return 1;
  }
  if (p.M == 0) return 0;
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.N, BN));
  int rc = p.a_mode == 0 ? dispatch<0>(p, epilogue, grid, stream) : dispatch<1>(p, epilogue, grid, stream);
  if (rc) return rc;
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_conv1(const Conv1Params& p, cudaStream_t stream) {
  if ((size_t)p.B * p.T1 * p.F1 * p.D == 0) return 0;
  if (p.D % 4 != 0 || p.D / 4 > 256 || p.pad_f > 1) {
    snprintf(g_errbuf, sizeof(g_errbuf), "conv1: unsupported geometry D=%d pad_f=%d", p.D, p.pad_f);
    return 1;
  }
  const int groups = p.D / 4;
  const int flanes = 256 / groups;
  const int threads = groups * flanes;
  const size_t smem = sizeof(float) * (2 * kConv1Rows + 1) * (p.F + 2);
  dim3 grid(ceil_div(p.T1, kConv1Rows), p.B);
  static int legacy = -1;
  if (legacy < 0) {
    const char* e = getenv("B200ASR_CONV1_LEGACY");
    legacy = (e && e[0] == '1') ? 1 : 0;
  }
  if (legacy && p.out_f16) {
    snprintf(g_errbuf, sizeof(g_errbuf), "conv1: the legacy kernel (B200ASR_CONV1_LEGACY=1) has no fp16 output");
    return 1;
  }
  if (!legacy) {
    B200_CUDA_OK(launch_k(conv1_f32x2_kernel, grid, dim3(threads), 2 * smem, stream, p, groups, flanes));
    B200_CUDA_OK(cudaGetLastError());
    return 0;
  }
  B200_CUDA_OK(launch_k(conv1_kernel, grid, dim3(threads), smem, stream, p, groups, flanes));
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace b200asr
