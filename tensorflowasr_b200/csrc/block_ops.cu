// LayerNorm, multi-head self-attention (no positional term, no mask unless banded) and the depthwise conv of the
// Conformer conv module, fp32 CUDA-core kernels.  Reference semantics: conformer_blocks.py:116,158,190,256 (LN eps 1e-3),
// multihead_attention.py:151-188, conformer_blocks.py:196-199 (SeparableConv1D depthwise part, 'same' padding).
#include "kernels.cuh"

namespace b200asr {

namespace {

// one warp per row, D <= 512
// pe (nullable): a [U, D] table added to the row before normalising, row m uses pe row m % U (RMHSAModule: LN(x + positional
// encoding), conformer_blocks.py:455-456).  round_tf32: the output only feeds a tensor-core GEMM.
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* __restrict__ y, int M, int D,
                                                        float eps, const float* __restrict__ pe, int U, int round_tf32) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + (size_t)row * D;
  const float* per = pe ? pe + (size_t)(row % U) * D : nullptr;
  float v[16];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int d = lane + i * 32;
    v[i] = (d < D) ? xr[d] + (per ? per[d] : 0.f) : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int d = lane + i * 32;
    const float c = (d < D) ? v[i] - mean : 0.f;
    q += c * c;
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)D + eps);
  float* yr = y + (size_t)row * D;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int d = lane + i * 32;
    if (d < D) {
      const float o = (v[i] - mean) * rstd * gamma[d] + beta[d];
      yr[d] = round_tf32 ? tf32_rn(o) : o;
    }
  }
}

// x[m, :] = table[clamp(ids[m]), :]   (tf.keras.layers.Embedding)
__global__ void __launch_bounds__(256) embed_kernel(const int* __restrict__ ids, const float* __restrict__ table, float* __restrict__ x, int M,
                                                    int D, int n_classes) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (size_t)M * D) return;
  const int m = (int)(i / D), d = (int)(i - (size_t)m * D);
  int id = ids[m];
  id = id < 0 ? 0 : (id >= n_classes ? n_classes - 1 : id);
  x[i] = table[(size_t)id * D + d];
}

// Cross attention (RMHSAModule -> MultiHeadAttention([q, enc, enc]), multihead_attention.py:151-188): one warp per (batch, head,
// query); keys stream 32 at a time with an online softmax.  q [B*U, H*dh] (pre-scaled), kv [B*Tk, 2*H*dh] (k | v), no mask.
__global__ void __launch_bounds__(128) cross_attention_kernel(const float* __restrict__ q, const float* __restrict__ kv, float* __restrict__ out,
                                                              int B, int U, int Tk, int H, int dh, int round_tf32) {
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 4 + warp;
  if (gw >= B * H * U) return;
  const int u = gw % U, h = (gw / U) % H, b = gw / (U * H);
  const int HD = H * dh;
  const float* qrow = q + ((size_t)b * U + u) * HD + h * dh;
  float m_run = -INFINITY, l_run = 0.f, o0 = 0.f, o1 = 0.f;
  for (int k0 = 0; k0 < Tk; k0 += 32) {
    const int key = k0 + lane;
    float sc = -INFINITY;
    if (key < Tk) {
      const float* krow = kv + ((size_t)b * Tk + key) * 2 * HD + h * dh;
      float acc = 0.f;
      for (int d = 0; d < dh; ++d) acc = fmaf(qrow[d], krow[d], acc);
      sc = acc;
    }
    const float m_new = fmaxf(m_run, warp_max(sc));
    const float corr = (m_run == -INFINITY) ? 0.f : expf(m_run - m_new);
    const float pj = (key < Tk) ? expf(sc - m_new) : 0.f;
    l_run = l_run * corr + warp_sum(pj);
    o0 *= corr;
    o1 *= corr;
    const int nk = min(32, Tk - k0);
    for (int j = 0; j < nk; ++j) {
      const float pw = __shfl_sync(0xffffffffu, pj, j);
      const float* vrow = kv + ((size_t)b * Tk + k0 + j) * 2 * HD + HD + h * dh;
      if (lane < dh) o0 = fmaf(pw, vrow[lane], o0);
      if (lane + 32 < dh) o1 = fmaf(pw, vrow[lane + 32], o1);
    }
    m_run = m_new;
  }
  const float inv = 1.0f / l_run;
  float* orow = out + ((size_t)b * U + u) * HD + h * dh;
  o0 *= inv;
  o1 *= inv;
  if (round_tf32) { o0 = tf32_rn(o0); o1 = tf32_rn(o1); }
  if (lane < dh) orow[lane] = o0;
  if (lane + 32 < dh) orow[lane + 32] = o1;
}

// Flash-style attention on CUDA cores: CTA = (batch, head, 64 queries), 256 threads as 16 x 16; keys/values stream through
// shared memory 64 at a time.  S = Q.K^T as 4x4 register micro-tiles (rows 4*ty.., columns tx+16*j so that every
// shared-memory access is either a broadcast or conflict-free), online softmax with warp-shuffle row reductions across the
// 16 lanes that share a query row, P staged in shared memory, O += P.V with CPT = ceil(dh/16) columns per thread.
// No mask and no positional term in the reference (multihead_attention.py:151-188); an optional band restricts keys
// for the chunk-streaming variant (chunk_conformer_blocks.py:158-176).
constexpr int kAttQ = 64, kAttK = 64;
template <int CPT>
__global__ void __launch_bounds__(256) attention_kernel(const AttnParams p, int dhs) {
  extern __shared__ __align__(16) float sm[];
  pdl_trigger();
  pdl_wait();
  const int dh = p.dh;
  float* Qs = sm;                      // [64][dhs]
  float* Ks = Qs + kAttQ * dhs;        // [64][dhs]
  float* Vs = Ks + kAttK * dhs;        // [64][dhs]
  float* Ps = Vs + kAttK * dhs;        // [64][68]
  constexpr int PS = kAttK + 4;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kAttQ;
  const int ld = 3 * p.H * dh;
  const float* base = p.qkv + (size_t)b * p.T * ld;
  const int nvec = dh >> 2;            // dh % 4 == 0
  for (int i = tid; i < kAttQ * nvec; i += 256) {
    const int r = i / nvec, v4 = i - r * nvec;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < p.T) v = *reinterpret_cast<const float4*>(base + (size_t)(q0 + r) * ld + h * dh + 4 * v4);
    *reinterpret_cast<float4*>(Qs + r * dhs + 4 * v4) = v;
  }
  float m[4], l[4], o[4][CPT];
  int lo[4], hi[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
#pragma unroll
    for (int c = 0; c < CPT; ++c) o[i][c] = 0.f;
    const int qi = q0 + 4 * ty + i;
    lo[i] = 0;
    hi[i] = p.T - 1;
    if (p.win_front >= 0) {
      lo[i] = max(min(max(qi - p.win_front, 0), p.T - p.win_back), 0);
      hi[i] = min(max(min(qi + p.win_back, p.T), p.win_back), p.T - 1);
    }
  }
  const float* kbase = base + p.H * dh + h * dh;
  const float* vbase = base + 2 * p.H * dh + h * dh;
  for (int j0 = 0; j0 < p.T; j0 += kAttK) {
    __syncthreads();   // previous tile fully consumed (also orders the Q staging before first use)
    for (int i = tid; i < kAttK * nvec; i += 256) {
      const int r = i / nvec, v4 = i - r * nvec;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (j0 + r < p.T) {
        kv = *reinterpret_cast<const float4*>(kbase + (size_t)(j0 + r) * ld + 4 * v4);
        vv = *reinterpret_cast<const float4*>(vbase + (size_t)(j0 + r) * ld + 4 * v4);
      }
      *reinterpret_cast<float4*>(Ks + r * dhs + 4 * v4) = kv;
      *reinterpret_cast<float4*>(Vs + r * dhs + 4 * v4) = vv;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    for (int d = 0; d < dh; d += 4) {
      float4 qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(Qs + (4 * ty + i) * dhs + d);
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * dhs + d);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] = fmaf(qv[i].x, kv[j].x, s[i][j]);
          s[i][j] = fmaf(qv[i].y, kv[j].y, s[i][j]);
          s[i][j] = fmaf(qv[i].z, kv[j].z, s[i][j]);
          s[i][j] = fmaf(qv[i].w, kv[j].w, s[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = j0 + tx + 16 * j;
        if (key >= p.T || key < lo[i] || key > hi[i]) s[i][j] = -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float mn = fmaxf(m[i], mx);
      float corr = 1.f, ps = 0.f;
      float pv[4] = {0.f, 0.f, 0.f, 0.f};
      if (mn != -INFINITY) {
        corr = (m[i] == -INFINITY) ? 0.f : expf(m[i] - mn);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          pv[j] = (s[i][j] == -INFINITY) ? 0.f : expf(s[i][j] - mn);
          ps += pv[j];
        }
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      l[i] = l[i] * corr + ps;
      m[i] = mn;
#pragma unroll
      for (int c = 0; c < CPT; ++c) o[i][c] *= corr;
#pragma unroll
      for (int j = 0; j < 4; ++j) Ps[(4 * ty + i) * PS + tx + 16 * j] = pv[j];
    }
    __syncthreads();
    for (int k = 0; k < kAttK; k += 4) {
      float4 pr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pr[i] = *reinterpret_cast<const float4*>(Ps + (4 * ty + i) * PS + k);
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const int col = tx + 16 * c;
        if (col < dh) {
          const float v0 = Vs[(k + 0) * dhs + col], v1 = Vs[(k + 1) * dhs + col], v2 = Vs[(k + 2) * dhs + col],
                      v3 = Vs[(k + 3) * dhs + col];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            o[i][c] = fmaf(pr[i].x, v0, o[i][c]);
            o[i][c] = fmaf(pr[i].y, v1, o[i][c]);
            o[i][c] = fmaf(pr[i].z, v2, o[i][c]);
            o[i][c] = fmaf(pr[i].w, v3, o[i][c]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + 4 * ty + i;
    if (qi >= p.T) continue;
    const float inv = 1.0f / l[i];
    float* orow = p.out + ((size_t)b * p.T + qi) * (p.H * dh) + h * dh;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int col = tx + 16 * c;
      if (col < dh) orow[col] = o[i][c] * inv;
    }
  }
}

// generic fallback (any kernel size)
__global__ void __launch_bounds__(256) dwconv_kernel(const DwConvParams p) {
  pdl_trigger();
  pdl_wait();
  const size_t total = (size_t)p.B * p.T * p.D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % p.D);
    const size_t r = i / p.D;
    const int t = (int)(r % p.T);
    const int b = (int)(r / p.T);
    const float* xb = p.x + (size_t)b * p.T * p.D + c;
    float acc = 0.f;
    for (int j = 0; j < p.K; ++j) {
      const int tt = t + j - p.pad_left;
      if (tt >= 0 && tt < p.T) acc = fmaf(xb[(size_t)tt * p.D], p.w[j * p.D + c], acc);
    }
    p.y[i] = p.round_tf32 ? tf32_rn(acc) : acc;
  }
}

// thread = channel, TT consecutive output frames per thread: the TT+K-1 inputs and the K taps live in registers, every
// global access is a coalesced row of D floats.  grid (ceil(T/TT), B), block = D rounded up to a warp multiple.
template <int K, int TT>
__global__ void __launch_bounds__(512) dwconv_reg_kernel(const DwConvParams p) {
  pdl_trigger();
  const int c = threadIdx.x;
  if (c >= p.D) return;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TT;
  const float* xb = p.x + (size_t)b * p.T * p.D + c;
  float w[K];
#pragma unroll
  for (int j = 0; j < K; ++j) w[j] = p.w[j * p.D + c];   // taps are weights: fetched while the producer of x is still running
  pdl_wait();
  float in[TT + K - 1];
#pragma unroll
  for (int i = 0; i < TT + K - 1; ++i) {
    const int tt = t0 + i - p.pad_left;
    in[i] = (tt >= 0 && tt < p.T) ? xb[(size_t)tt * p.D] : 0.f;
  }
  float* yb = p.y + (size_t)b * p.T * p.D + c;
#pragma unroll
  for (int o = 0; o < TT; ++o) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) acc = fmaf(in[o + j], w[j], acc);
    if (t0 + o < p.T) yb[(size_t)(t0 + o) * p.D] = p.round_tf32 ? tf32_rn(acc) : acc;
  }
}

// Two channels per thread with packed fp32x2 FMAs (bit-identical to the scalar kernel: two independent round-to-nearest FMAs):
// half the instructions, 8-byte coalesced rows.  D must be even.
__device__ __forceinline__ unsigned long long dw_ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
  return c;
}
template <int K, int TT>
__global__ void __launch_bounds__(256) dwconv_reg2_kernel(const DwConvParams p) {
  pdl_trigger();
  const int c = 2 * threadIdx.x;
  if (c >= p.D) return;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TT;
  const float* xb = p.x + (size_t)b * p.T * p.D + c;
  unsigned long long w[K];
#pragma unroll
  for (int j = 0; j < K; ++j) w[j] = *reinterpret_cast<const unsigned long long*>(p.w + j * p.D + c);   // taps are weights
  pdl_wait();
  unsigned long long in[TT + K - 1];
#pragma unroll
  for (int i = 0; i < TT + K - 1; ++i) {
    const int tt = t0 + i - p.pad_left;
    in[i] = (tt >= 0 && tt < p.T) ? *reinterpret_cast<const unsigned long long*>(xb + (size_t)tt * p.D) : 0ull;
  }
  float* yb = p.y + (size_t)b * p.T * p.D + c;
#pragma unroll
  for (int o = 0; o < TT; ++o) {
    unsigned long long acc = 0ull;
#pragma unroll
    for (int j = 0; j < K; ++j) acc = dw_ffma2(in[o + j], w[j], acc);
    if (p.round_tf32) acc = (acc + 0x0000100000001000ull) & 0xFFFFE000FFFFE000ull;   // both halves to nearest tf32 (a carry out of the low half needs bits >= 0xFFFFF000, a NaN)
    if (t0 + o < p.T) *reinterpret_cast<unsigned long long*>(yb + (size_t)(t0 + o) * p.D) = acc;
  }
}

}  // namespace

int launch_layernorm(const float* x, const float* gamma, const float* beta, float* y, int M, int D, float eps,
                     cudaStream_t stream, const float* pe, int U, int round_tf32) {
  if (D > 512) {
    snprintf(g_errbuf, sizeof(g_errbuf), "layernorm: D=%d > 512 unsupported", D);
    return 1;
  }
  if (M == 0) return 0;
  B200_CUDA_OK(launch_k(layernorm_kernel, dim3(ceil_div(M, 8)), dim3(256), 0, stream, x, gamma, beta, y, M, D, eps, pe, U > 0 ? U : 1, round_tf32));
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_embed(const int* ids, const float* table, float* x, int M, int D, int n_classes, cudaStream_t stream) {
  if ((size_t)M * D == 0) return 0;
  B200_CUDA_OK(launch_k(embed_kernel, dim3((unsigned)(((size_t)M * D + 255) / 256)), dim3(256), 0, stream, ids, table, x, M, D, n_classes));
  return 0;
}

int launch_cross_attention(const float* q, const float* kv, float* out, int B, int U, int Tk, int H, int dh, int round_tf32, cudaStream_t stream) {
  if (dh > 64) {
    snprintf(g_errbuf, sizeof(g_errbuf), "cross_attention: head_size=%d unsupported (needs <= 64)", dh);
    return 1;
  }
  const int total = B * H * U;
  if (total == 0 || Tk == 0) return 0;
  B200_CUDA_OK(launch_k(cross_attention_kernel, dim3(ceil_div(total, 4)), dim3(128), 0, stream, q, kv, out, B, U, Tk, H, dh, round_tf32));
  return 0;
}

int launch_attention(const AttnParams& p, cudaStream_t stream) {
  if (p.dh > 64 || p.dh % 4 != 0) {
    snprintf(g_errbuf, sizeof(g_errbuf), "attention: head_size=%d unsupported (needs <= 64, multiple of 4)", p.dh);
    return 1;
  }
  if (p.B == 0 || p.T == 0) return 0;
  int dhs = p.dh;                       // row stride = 4 (mod 32) floats: 16-lane float4 reads hit every bank once
  while (dhs % 32 != 4) dhs += 4;
  const size_t smem = sizeof(float) * (size_t)(3 * 64 * dhs + 64 * 68);
  dim3 grid(ceil_div(p.T, kAttQ), p.H, p.B);
  const int cpt = ceil_div(p.dh, 16);
  static PerDeviceSmem configured;
  if (configured.need(1)) {
    const int big = (int)(sizeof(float) * (3 * 64 * 68 + 64 * 68));
    B200_CUDA_OK(cudaFuncSetAttribute(attention_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    B200_CUDA_OK(cudaFuncSetAttribute(attention_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    B200_CUDA_OK(cudaFuncSetAttribute(attention_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    B200_CUDA_OK(cudaFuncSetAttribute(attention_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  }
  switch (cpt) {
    case 1: B200_CUDA_OK(launch_k(attention_kernel<1>, grid, dim3(256), smem, stream, p, dhs)); break;
    case 2: B200_CUDA_OK(launch_k(attention_kernel<2>, grid, dim3(256), smem, stream, p, dhs)); break;
    case 3: B200_CUDA_OK(launch_k(attention_kernel<3>, grid, dim3(256), smem, stream, p, dhs)); break;
    default: B200_CUDA_OK(launch_k(attention_kernel<4>, grid, dim3(256), smem, stream, p, dhs)); break;
  }
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_dwconv(const DwConvParams& p, cudaStream_t stream) {
  const size_t total = (size_t)p.B * p.T * p.D;
  if (total == 0) return 0;
  const int threads = ceil_div(p.D, 32) * 32;
  if (p.K == 32 && p.D % 2 == 0 && p.D <= 512 && ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.y) | reinterpret_cast<uintptr_t>(p.w)) & 7) == 0) {
    constexpr int TT = 8;
    const int th2 = ceil_div(p.D / 2, 32) * 32;
    B200_CUDA_OK(launch_k(dwconv_reg2_kernel<32, TT>, dim3(ceil_div(p.T, TT), p.B), dim3(th2), 0, stream, p));
  } else if (p.K == 32 && threads <= 512) {
    constexpr int TT = 8;
    B200_CUDA_OK(launch_k(dwconv_reg_kernel<32, TT>, dim3(ceil_div(p.T, TT), p.B), dim3(threads), 0, stream, p));
  } else if (p.K == 5 && threads <= 512) {
    constexpr int TT = 16;
    B200_CUDA_OK(launch_k(dwconv_reg_kernel<5, TT>, dim3(ceil_div(p.T, TT), p.B), dim3(threads), 0, stream, p));
  } else {
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    B200_CUDA_OK(launch_k(dwconv_kernel, dim3(blocks), dim3(256), 0, stream, p));
  }
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace b200asr
