// LayerNorm, multi-head self-attention (no positional term, no mask unless banded) and the depthwise conv of the
// Conformer conv module, fp32 CUDA-core kernels.  Reference semantics: conformer_blocks.py:116,158,190,256 (LN eps 1e-3),
// multihead_attention.py:151-188, conformer_blocks.py:196-199 (SeparableConv1D depthwise part, 'same' padding).
#include "kernels.cuh"

namespace b200asr {

namespace {

// one warp per row, D <= 512
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* __restrict__ y, int M, int D,
                                                        float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + (size_t)row * D;
  float v[16];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int d = lane + i * 32;
    v[i] = (d < D) ? xr[d] : 0.f;
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
    if (d < D) yr[d] = (v[i] - mean) * rstd * gamma[d] + beta[d];
  }
}

// One warp per query row; keys/values streamed through shared memory in tiles of 32.  dh <= 64.
constexpr int kQW = 8;  // queries (warps) per CTA
__global__ void __launch_bounds__(kQW * 32) attention_kernel(const AttnParams p) {
  extern __shared__ float sm[];
  const int dh = p.dh, ldk = dh + 1;
  float* Ks = sm;                 // [32][dh+1]
  float* Vs = Ks + 32 * ldk;      // [32][dh+1]
  float* Qs = Vs + 32 * ldk;      // [kQW][dh]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * kQW + warp;
  const int ld = 3 * p.H * dh;
  const float* base = p.qkv + (size_t)b * p.T * ld;
  const bool q_ok = qi < p.T;
  if (q_ok) {
    for (int d = lane; d < dh; d += 32) Qs[warp * dh + d] = base[(size_t)qi * ld + h * dh + d];
  }
  int lo = 0, hi = p.T - 1;
  if (p.win_front >= 0) {  // chunk_conformer_blocks.py:158-176
    lo = min(max(qi - p.win_front, 0), p.T - p.win_back);
    hi = max(min(qi + p.win_back, p.T), p.win_back);
    lo = max(lo, 0);
    hi = min(hi, p.T - 1);
  }
  float m = -INFINITY, l = 0.f, acc0 = 0.f, acc1 = 0.f;
  const float* kbase = base + p.H * dh + h * dh;
  const float* vbase = base + 2 * p.H * dh + h * dh;
  for (int j0 = 0; j0 < p.T; j0 += 32) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * dh; i += kQW * 32) {
      const int j = i / dh, d = i - j * dh;
      const int key = j0 + j;
      float kv = 0.f, vv = 0.f;
      if (key < p.T) {
        kv = kbase[(size_t)key * ld + d];
        vv = vbase[(size_t)key * ld + d];
      }
      Ks[j * ldk + d] = kv;
      Vs[j * ldk + d] = vv;
    }
    __syncthreads();
    if (!q_ok) continue;
    const int key = j0 + lane;
    float s = -INFINITY;
    if (key < p.T && key >= lo && key <= hi) {
      s = 0.f;
      const float* kr = Ks + lane * ldk;
      const float* qr = Qs + warp * dh;
      for (int d = 0; d < dh; ++d) s = fmaf(qr[d], kr[d], s);
    }
    const float mt = warp_max(s);
    if (mt == -INFINITY) continue;
    const float mn = fmaxf(m, mt);
    const float corr = (m == -INFINITY) ? 0.f : expf(m - mn);
    const float pj = (s == -INFINITY) ? 0.f : expf(s - mn);
    l = l * corr + warp_sum(pj);
    acc0 *= corr;
    acc1 *= corr;
    m = mn;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const float pb = __shfl_sync(0xffffffffu, pj, j);
      const float* vr = Vs + j * ldk;
      if (lane < dh) acc0 = fmaf(pb, vr[lane], acc0);
      if (lane + 32 < dh) acc1 = fmaf(pb, vr[lane + 32], acc1);
    }
  }
  if (q_ok) {
    float* o = p.out + ((size_t)b * p.T + qi) * (p.H * dh) + h * dh;
    const float inv = 1.0f / l;
    if (lane < dh) o[lane] = acc0 * inv;
    if (lane + 32 < dh) o[lane + 32] = acc1 * inv;
  }
}

__global__ void __launch_bounds__(256) dwconv_kernel(const DwConvParams p) {
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
    p.y[i] = acc;
  }
}

}  // namespace

int launch_layernorm(const float* x, const float* gamma, const float* beta, float* y, int M, int D, float eps,
                     cudaStream_t stream) {
  if (D > 512) {
    snprintf(g_errbuf, sizeof(g_errbuf), "layernorm: D=%d > 512 unsupported", D);
    return 1;
  }
  if (M == 0) return 0;
  layernorm_kernel<<<ceil_div(M, 8), 256, 0, stream>>>(x, gamma, beta, y, M, D, eps);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_attention(const AttnParams& p, cudaStream_t stream) {
  if (p.dh > 64) {
    snprintf(g_errbuf, sizeof(g_errbuf), "attention: head_size=%d > 64 unsupported", p.dh);
    return 1;
  }
  if (p.B == 0 || p.T == 0) return 0;
  const size_t smem = sizeof(float) * (2 * 32 * (p.dh + 1) + kQW * p.dh);
  dim3 grid(ceil_div(p.T, kQW), p.H, p.B);
  attention_kernel<<<grid, kQW * 32, smem, stream>>>(p);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_dwconv(const DwConvParams& p, cudaStream_t stream) {
  const size_t total = (size_t)p.B * p.T * p.D;
  if (total == 0) return 0;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  dwconv_kernel<<<blocks, 256, 0, stream>>>(p);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace b200asr
