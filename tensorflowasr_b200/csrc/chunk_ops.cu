// ChunkConformer state-cache streaming (asr/models/chunk_conformer_blocks.py): the small HBM / latency-bound kernels around the
// GEMMs of one streaming step -- waveform / mel cache roll, attention of the new frames over [cache | new frames] with the chunk
// band mask, causal depthwise conv over [cache | new frames], cache roll with look-ahead drop, feature_pick compaction.
//
// Cache representation.  The reference caches the INPUT rows of the MHSA / conv modules and recomputes LayerNorm + K/V (resp.
// LayerNorm + pointwise conv + GLU) of the cached rows at every step (:202-216, :294-311).  Both are per-row functions, so the
// caches here hold their RESULTS instead: K|V rows [B, W = win_front, 2 H dh] and GLU rows [B, K - 1, D], right-aligned in a
// fixed-size buffer (invalid / not-yet-filled rows on the left).  A zero GLU row is exactly the 'causal' zero padding of an
// empty cache; for K|V the number of valid rows `c` is tracked by the host (all streams of a state advance in lockstep).
// The reference's cache roll  cat(cache, cur)[:, :-win_back][:, -W:]  (:532-563, :626-658; it drops the look-ahead rows, and even
// valid cache rows when fewer than win_back rows arrive) is  new[r] = cat(old, cur)[r + Tc - win_back]  on this representation.
#include "kernels.cuh"

namespace b200asr {

namespace {

// wavbuf [B, 2S] <- [wavbuf[:, S:], chunk];  one thread per element of the first half (reads its own second-half element first)
__global__ void __launch_bounds__(256) stream_wav_shift_kernel(float* __restrict__ wavbuf, const float* __restrict__ chunk, int B, int S) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (size_t)B * S) return;
  const size_t b = i / S, k = i - b * S;
  float* row = wavbuf + b * 2 * S;
  row[k] = row[S + k];
  row[S + k] = chunk[i];
}

// melcat [B, sub + n, F] = [sub_cache | mel_new];  sub_cache <- last `sub` rows of melcat (ConvSubsampling.stream_call :74-91)
__global__ void __launch_bounds__(256) stream_mel_cat_kernel(float* __restrict__ melcat, float* __restrict__ sub_cache,
                                                             const float* __restrict__ mel_new, int B, int sub, int n, int F) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t per_b = (size_t)(sub + n) * F;
  if (i >= (size_t)B * per_b) return;
  const size_t b = i / per_b, rem = i - b * per_b;
  const int r = (int)(rem / F), f = (int)(rem - (size_t)r * F);
  if (r < sub) {
    // this thread owns cache element (b, r, f): it reads the old value and installs the new one (= row n + r of the concatenation
    // = mel_new row n + r - sub; n >= sub is enforced by the launcher), so no other thread touches it
    const size_t ci = (b * sub + r) * F + f;
    melcat[i] = sub_cache[ci];
    sub_cache[ci] = mel_new[(b * n + (n + r - sub)) * F + f];
  } else {
    melcat[i] = mel_new[(b * n + (r - sub)) * F + f];
  }
}

// out [B, na + nb, D] = [a [B, na, D] | b [B, nb, D]]
__global__ void __launch_bounds__(256) rows_cat_kernel(const float* __restrict__ a, int na, const float* __restrict__ bsrc, int nb,
                                                       float* __restrict__ out, int B, int D) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t per_b = (size_t)(na + nb) * D;
  if (i >= (size_t)B * per_b) return;
  const size_t b = i / per_b, rem = i - b * per_b;
  const int r = (int)(rem / D), d = (int)(rem - (size_t)r * D);
  out[i] = r < na ? a[(b * na + r) * D + d] : bsrc[(b * nb + (r - na)) * D + d];
}

// dst [B, n, D] = src [B, T, D][:, t0 : t0 + n]
__global__ void __launch_bounds__(256) rows_slice_kernel(const float* __restrict__ src, int T, int t0, int n, float* __restrict__ dst, int B, int D) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t per_b = (size_t)n * D;
  if (i >= (size_t)B * per_b) return;
  const size_t b = i / per_b, rem = i - b * per_b;
  dst[i] = src[(b * T + t0) * D + rem];
}

// new_cache[b, r, :] = cat(old_cache[b] (W rows), cur[b] (Tc rows))[r + shift], zero for a negative index
__global__ void __launch_bounds__(256) stream_cache_update_kernel(const CacheUpdateParams p) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t per_b = (size_t)p.W * p.C;
  if (i >= (size_t)p.B * per_b) return;
  const size_t b = i / per_b, rem = i - b * per_b;
  const int r = (int)(rem / p.C), cc = (int)(rem - (size_t)r * p.C);
  const int s = r + p.shift;
  float v = 0.f;
  if (s >= 0) {
    if (s < p.W) v = p.old_cache[(b * p.W + s) * p.C + cc];
    else if (s - p.W < p.Tc) v = p.cur[(b * p.Tc + (s - p.W)) * (size_t)p.cur_ld + p.cur_col0 + cc];
  }
  p.new_cache[i] = v;
}

// One warp per (stream, head, query row).  Keys = the last `c` rows of the K|V cache followed by the Tc new rows; the band mask
// of ChunkMHSAModule._compute_chunk_mask (:158-176) is evaluated on positions inside that concatenation (n = c + Tc rows), exactly
// as stream_call does (:202-216).  fp32 throughout (the step is launch-latency bound, not math bound).
__global__ void __launch_bounds__(128) stream_attention_kernel(const StreamAttnParams p) {
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 4 + warp;
  const int total = p.B * p.H * p.Tc;
  if (gw >= total) return;
  const int j = gw % p.Tc, h = (gw / p.Tc) % p.H, b = gw / (p.Tc * p.H);
  const int HD = p.H * p.dh, n = p.c + p.Tc, idx = p.c + j;
  // mask row (chunk_mask): attend to columns [low, high]
  int low = max(idx - p.win_front, 0);
  int high = min(max(idx + p.win_back, 0), n);
  low = low - max(low - n + p.win_back, 0);
  high = high + max(p.win_back - high, 0);
  const float* qrow = p.qkv + ((size_t)b * p.Tc + j) * 3 * HD + h * p.dh;
  auto krow = [&](int col) -> const float* {   // K row of concatenation position `col` (V follows HD floats later)
    return col < p.c ? p.kv_cache + ((size_t)b * p.W + (p.W - p.c + col)) * 2 * HD + h * p.dh
                     : p.qkv + ((size_t)b * p.Tc + (col - p.c)) * 3 * HD + HD + h * p.dh;
  };
  // scores: lane handles columns lane, lane + 32, ... (n <= 4 * 32)
  float sc[4];
  float mx = -INFINITY;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int col = lane + 32 * u;
    sc[u] = -INFINITY;
    if (col < n && col >= low && col <= high) {
      const float* k = krow(col);
      float acc = 0.f;
      for (int d = 0; d < p.dh; ++d) acc = fmaf(qrow[d], k[d], acc);
      sc[u] = acc;
      mx = fmaxf(mx, acc);
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    sc[u] = (sc[u] == -INFINITY) ? 0.f : expf(sc[u] - mx);
    sum += sc[u];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  // output: lane handles dims lane, lane + 32 (dh <= 64)
  float o0 = 0.f, o1 = 0.f;
  for (int col = max(low, 0); col <= min(high, n - 1); ++col) {
    const int u = col >> 5;
    const float mine = u == 0 ? sc[0] : (u == 1 ? sc[1] : (u == 2 ? sc[2] : sc[3]));
    const float pj = __shfl_sync(0xffffffffu, mine, col & 31);
    const float* v = krow(col) + HD;            // V sits HD floats after K in both the cache row (K|V) and the qkv row (q|K|V)
    if (lane < p.dh) o0 = fmaf(pj, v[lane], o0);
    if (lane + 32 < p.dh) o1 = fmaf(pj, v[lane + 32], o1);
  }
  float* orow = p.out + ((size_t)b * p.Tc + j) * HD + h * p.dh;
  o0 *= inv;
  o1 *= inv;
  if (p.round_tf32) { o0 = tf32_rn(o0); o1 = tf32_rn(o1); }
  if (lane < p.dh) orow[lane] = o0;
  if (lane + 32 < p.dh) orow[lane + 32] = o1;
}

// y[b, j, c] = sum_k cat(cache[b] (K - 1 rows), cur[b] (Tc rows))[j + k, c] * w[k, c]   ('causal' SeparableConv1D depthwise part)
__global__ void __launch_bounds__(256) stream_dwconv_kernel(const StreamDwParams p) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t per_b = (size_t)p.Tc * p.D;
  if (i >= (size_t)p.B * per_b) return;
  const size_t b = i / per_b, rem = i - b * per_b;
  const int j = (int)(rem / p.D), c = (int)(rem - (size_t)j * p.D);
  const int W = p.K - 1;
  float acc = 0.f;
  for (int k = 0; k < p.K; ++k) {
    const int s = j + k;
    const float x = s < W ? p.cache[(b * W + s) * p.D + c] : p.cur[(b * p.Tc + (s - W)) * p.D + c];
    acc = fmaf(x, p.w[k * p.D + c], acc);
  }
  p.y[i] = p.round_tf32 ? tf32_rn(acc) : acc;
}

// ChunkConformer.feature_pick (:913-999): per stream keep the frames whose phone argmax (first maximum) is not blank, compacted to
// the front of feats / picked, zero rows behind; counts[b] = kept frames, *n_max = max over the batch.  One CTA per stream.
__global__ void __launch_bounds__(256) feature_pick_kernel(const PickParams p) {
  extern __shared__ int keep_s[];   // [T] flags, then [T] exclusive prefix
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* pref = keep_s + p.T;
  for (int t = warp; t < p.T; t += 8) {
    const float* row = p.logits + ((size_t)b * p.T + t) * p.ldv;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int v = lane; v < p.V; v += 32) {
      const float x = row[v];
      if (x > best) { best = x; bi = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) keep_s[t] = (bi != p.blank) ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int t = 0; t < p.T; ++t) { pref[t] = run; run += keep_s[t]; }
    p.counts[b] = run;
    atomicMax(p.n_max, run);
  }
  __syncthreads();
  const int cnt = p.counts[b];
  for (int i = threadIdx.x; i < p.T * p.D; i += 256) {
    const int t = i / p.D, d = i - t * p.D;
    if (keep_s[t]) p.feats[((size_t)b * p.T + pref[t]) * p.D + d] = p.hidden[((size_t)b * p.T + t) * p.D + d];
    if (t >= cnt) p.feats[((size_t)b * p.T + t) * p.D + d] = 0.f;
  }
  if (p.picked) {
    for (int i = threadIdx.x; i < p.T * p.V; i += 256) {
      const int t = i / p.V, v = i - t * p.V;
      if (keep_s[t]) p.picked[((size_t)b * p.T + pref[t]) * p.V + v] = p.logits[((size_t)b * p.T + t) * p.ldv + v];
      if (t >= cnt) p.picked[((size_t)b * p.T + t) * p.V + v] = 0.f;
    }
  }
}

// dst = round-to-nearest-tf32(src)
__global__ void __launch_bounds__(256) round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) dst[i] = tf32_rn(src[i]);
}

inline int blocks_for(size_t n) { return (int)((n + 255) / 256); }

}  // namespace

int launch_stream_wav_shift(float* wavbuf, const float* chunk, int B, int S, cudaStream_t stream) {
  if (B * S == 0) return 0;
  B200_CUDA_OK(launch_k(stream_wav_shift_kernel, dim3(blocks_for((size_t)B * S)), dim3(256), 0, stream, wavbuf, chunk, B, S));
  return 0;
}

int launch_stream_mel_cat(float* melcat, float* sub_cache, const float* mel_new, int B, int sub, int n, int F, cudaStream_t stream) {
  if (n < sub) {
    snprintf(g_errbuf, sizeof(g_errbuf), "stream_mel_cat: chunk of %d mel frames is shorter than the %d-frame cache", n, sub);
    return 1;
  }
  B200_CUDA_OK(launch_k(stream_mel_cat_kernel, dim3(blocks_for((size_t)B * (sub + n) * F)), dim3(256), 0, stream, melcat, sub_cache, mel_new, B, sub, n, F));
  return 0;
}

int launch_rows_cat(const float* a, int na, const float* b, int nb, float* out, int B, int D, cudaStream_t stream) {
  if ((size_t)B * (na + nb) * D == 0) return 0;
  B200_CUDA_OK(launch_k(rows_cat_kernel, dim3(blocks_for((size_t)B * (na + nb) * D)), dim3(256), 0, stream, a, na, b, nb, out, B, D));
  return 0;
}

int launch_rows_slice(const float* src, int T, int t0, int n, float* dst, int B, int D, cudaStream_t stream) {
  if ((size_t)B * n * D == 0) return 0;
  B200_CUDA_OK(launch_k(rows_slice_kernel, dim3(blocks_for((size_t)B * n * D)), dim3(256), 0, stream, src, T, t0, n, dst, B, D));
  return 0;
}

int launch_stream_cache_update(const CacheUpdateParams& p, cudaStream_t stream) {
  B200_CUDA_OK(launch_k(stream_cache_update_kernel, dim3(blocks_for((size_t)p.B * p.W * p.C)), dim3(256), 0, stream, p));
  return 0;
}

int launch_stream_attention(const StreamAttnParams& p, cudaStream_t stream) {
  if (p.dh > 64 || p.c + p.Tc > 128 || p.c > p.W) {
    snprintf(g_errbuf, sizeof(g_errbuf), "stream_attention: unsupported geometry dh=%d keys=%d", p.dh, p.c + p.Tc);
    return 1;
  }
  const int total = p.B * p.H * p.Tc;
  if (total == 0) return 0;
  B200_CUDA_OK(launch_k(stream_attention_kernel, dim3(ceil_div(total, 4)), dim3(128), 0, stream, p));
  return 0;
}

int launch_stream_dwconv(const StreamDwParams& p, cudaStream_t stream) {
  if ((size_t)p.B * p.Tc * p.D == 0) return 0;
  B200_CUDA_OK(launch_k(stream_dwconv_kernel, dim3(blocks_for((size_t)p.B * p.Tc * p.D)), dim3(256), 0, stream, p));
  return 0;
}

int launch_feature_pick(const PickParams& p, cudaStream_t stream) {
  if (p.B == 0 || p.T == 0) return 0;
  B200_CUDA_OK(cudaMemsetAsync(p.n_max, 0, sizeof(int), stream));
  B200_CUDA_OK(launch_k(feature_pick_kernel, dim3(p.B), dim3(256), (size_t)2 * p.T * sizeof(int), stream, p));
  return 0;
}

int launch_round_tf32(const float* src, float* dst, size_t n, cudaStream_t stream) {
  if (n == 0) return 0;
  B200_CUDA_OK(launch_k(round_tf32_kernel, dim3(blocks_for(n)), dim3(256), 0, stream, src, dst, n));
  return 0;
}

}  // namespace b200asr
