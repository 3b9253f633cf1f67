// Kernel launch interfaces (internal).  All pointers are device pointers, all launches are asynchronous on `stream`.
#pragma once
#include "common.cuh"

namespace b200asr {

// ------------------------------------------------------------------------------------------- frontend.cu
struct FrontendParams {
  const float* wav;        // [B, L]
  const float* window;     // [1024]
  const float2* twiddle;   // [1024] exp(-2*pi*i*m/1024)
  const float* melw;       // [513, n_mels]
  const int* mel_lo;       // [n_mels] first bin with non-zero weight
  const int* mel_hi;       // [n_mels] one past last
  const int* mel_off;      // [n_mels + 1] offsets into mel_wc (compact band weights: mel_wc[mel_off[m] + k - mel_lo[m]] = melw[k, m])
  const float* mel_wc;     // [mel_nnz]
  int mel_nnz;
  float* power;            // [B, T, power_stride] scratch
  unsigned int* pmax;      // [B] scratch
  float* mel;              // [B, T, n_mels] out
  int B, L, T, pad_left, hop, power_stride, n_mels;
  int mode;                // 0 offline ('same', dB w/ per-utterance max), 1 chunk ('valid', log10 only)
};
int launch_frontend(const FrontendParams& p, cudaStream_t stream);

// ------------------------------------------------------------------------------------------- subsample.cu
struct Conv1Params {
  const float* mel;   // [B, T, F]
  const float* w;     // [9, D] tap-major
  const float* bias;  // [D]
  float* out;         // [B, T1, F1, D]
  int B, T, F, T1, F1, D, pad_t, pad_f;
  int round_tf32 = 0;   // 1: outputs rounded to nearest tf32 (they feed conv2's tensor-core A operand only)
  int out_f16 = 0;      // 1: `out` holds IEEE fp16 [B, T1, F1, D] (same 11-bit significand as tf32, half the bytes; conv2 then runs kind::f16)
};
int launch_conv1(const Conv1Params& p, cudaStream_t stream);

// ------------------------------------------------------------------------------------------- gemm_simt.cu
enum Epilogue : int {
  EPI_BIAS = 0,        // C = acc + bias
  EPI_BIAS_RELU = 1,   // C = relu(acc + bias)
  EPI_BIAS_SWISH = 2,  // C = swish(acc + bias)
  EPI_GLU = 3,         // C[:, j] = (acc[2j]+b[2j]) * sigmoid(acc[2j+1]+b[2j+1])   (weights pre-interleaved), ldc = N/2
  EPI_RESID = 4,       // C = resid + alpha * (acc + bias)
  EPI_NONE = 5,        // C = acc
  // fused LayerNorm epilogues (tcgen05 path only; need the whole row in one tile: N <= 256)
  EPI_RESID_LN = 6,    // x = resid + alpha*(acc+bias) -> C;  LN(x; ln1) -> C2
  EPI_RESID_LN2 = 7,   // y = LN(resid + alpha*(acc+bias); ln1) -> C;  LN(y; ln2) -> C2 (skipped when ln2_g == null)
  EPI_BIAS_LN = 8,     // x = acc + bias -> C;  LN(x; ln1) -> C2
  // CTC head fused with the greedy decoder's per-frame argmax (tcgen05 path only): no logits are stored;
  // C = float2 [M, ceil(N / BLOCK_N)] holding (max, bit pattern of the first arg max) of acc + bias over each N tile
  EPI_BIAS_ARGMAX = 9,
};
inline bool epi_is_ln(int e) { return e == EPI_RESID_LN || e == EPI_RESID_LN2 || e == EPI_BIAS_LN; }

struct GemmParams {
  const float* A;      // [M, K] row-major (lda), or conv2 source [B, T1, F1, D] when a_mode == 1
  const float* W;      // [N, K] row-major (K-major "B" operand)
  const float* bias;   // [N] or null
  const float* resid;  // [M, ldc] (EPI_RESID)
  float* C;            // [M, ldc]
  int M, N, K, lda, ldc;
  float alpha;
  // fused LayerNorm epilogues
  const float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  float* C2 = nullptr;  // [M, ldc]
  float ln_eps = 1e-3f;
  // implicit-GEMM geometry for the second subsampling conv (3x3, stride 2, TF 'same')
  int a_mode;          // 0 plain, 1 conv2 im2col
  int T1, F1, T2, F2, D, pad_t, pad_f;
  // tcgen05 path: store C (plain epilogues) / C2 (LayerNorm epilogues) rounded to nearest tf32 -- set when the tensor is only
  // read as a tensor-core operand again (the datapath would truncate raw fp32 bits)
  int round_out = 0;
  // tcgen05 path (conv2: a_mode 1 + EPI_BIAS_RELU; subsampling linear layer: a_mode 0 + EPI_BIAS_LN): A and W hold IEEE fp16 --
  // products of 11-bit significands accumulated in fp32, exactly what kind::tf32 computes on tf32-rounded operands, at twice the MMA
  // rate and half the operand bytes.  lda counts halves then.
  int f16 = 0;
  int out_f16 = 0;     // plain epilogues: C holds IEEE fp16 [M, ldc] (ldc in halves)
};
int launch_gemm_simt(const GemmParams& p, int epilogue, cudaStream_t stream);

// ------------------------------------------------------------------------------------------- block_ops.cu
// pe (nullable): [U, D] table added before normalising (row m uses pe row m % U); round_tf32: output rounded to nearest tf32
int launch_layernorm(const float* x, const float* gamma, const float* beta, float* y, int M, int D, float eps,
                     cudaStream_t stream, const float* pe = nullptr, int U = 1, int round_tf32 = 0);
int launch_embed(const int* ids, const float* table, float* x, int M, int D, int n_classes, cudaStream_t stream);
// q [B*U, H*dh] (pre-scaled), kv [B*Tk, 2*H*dh] (k | v) -> out [B*U, H*dh]; no mask
int launch_cross_attention(const float* q, const float* kv, float* out, int B, int U, int Tk, int H, int dh, int round_tf32, cudaStream_t stream);

struct AttnParams {
  const float* qkv;   // [B*T, 3*H*dh]: q | k | v, each h-major (q already scaled by 1/sqrt(dh) through Wq)
  float* out;         // [B*T, H*dh]
  int B, T, H, dh;
  // optional band mask (ChunkConformer, chunk_conformer_blocks.py:158-176); win_front < 0 => full attention
  int win_front, win_back;
  long long* dbg = nullptr;   // optional clock64 timeline of CTA 0 (B200ASR_ATTN_DBG=1)
  int round_tf32 = 0;         // 1: outputs rounded to nearest tf32 (they feed the out-projection's tensor-core A operand only)
  int async_stage = 0;        // 1 (tcgen05 kernel): qkv holds tf32 numbers already (its producer rounded them): stage Q / K / V^T with cp.async
};
int launch_attention(const AttnParams& p, cudaStream_t stream);          // fp32 CUDA cores (block_ops.cu)
bool attention_tc_supported(const AttnParams& p);
int launch_attention_tc(const AttnParams& p, cudaStream_t stream);       // tcgen05 tf32 (attention_tc.cu)

struct DwConvParams {
  const float* x;   // [B*T, D]
  const float* w;   // [K, D]
  float* y;         // [B*T, D]
  int B, T, D, K, pad_left;
  int round_tf32 = 0;   // 1: outputs rounded to nearest tf32 (they feed the pointwise conv's tensor-core A operand only)
};
int launch_dwconv(const DwConvParams& p, cudaStream_t stream);

// ------------------------------------------------------------------------------------------- ctc_decode.cu
// argmax over V per frame (first maximum wins, as ctc_greedy_decoder.h:11-18), then merge repeats / drop blank.
int launch_ctc_greedy(const float* logits, const int* lengths /*nullable*/, int B, int T, int V, int blank,
                      int* frame_argmax /*[B*T] scratch*/, int* ids /*[B, T]*/, int* out_len /*[B]*/, cudaStream_t stream);
// the same from per-tile (max, argmax) partials written by the EPI_BIAS_ARGMAX GEMM epilogue: part [B*T, n_tiles] float2
int launch_ctc_greedy_partials(const float2* part, int n_tiles, const int* lengths /*nullable*/, int B, int T, int blank,
                               int* frame_argmax, int* ids, int* out_len, cudaStream_t stream);
// number of N tiles the tcgen05 GEMM uses for an N-column EPI_BIAS_ARGMAX launch (gemm_tc.cu)
int tc_argmax_tiles(int N);

struct BeamParams {
  const float* logits;   // [B, T, V] (pre-softmax) -- softmax is applied inside, as the callers of ctc_beam_search_decoder do
  const int* lengths;    // nullable
  int B, T, V, blank, beam, cutoff_top_n;
  float cutoff_prob;
  int* ids;              // [B, beam, T]
  int* out_len;          // [B, beam]
  float* scores;         // [B, beam]
  void* workspace;       // beam_workspace_bytes(B, T, beam)
  int is_prob = 0;       // 1: `logits` already holds probabilities (the reference decoder's own input convention, probs_seq)
};
size_t beam_workspace_bytes(int B, int T, int beam);
int launch_ctc_beam(const BeamParams& p, cudaStream_t stream);

// ------------------------------------------------------------------------------------------- chunk_ops.cu
// ChunkConformer state-cache streaming helpers (see the file header for the cache representation).
int launch_stream_wav_shift(float* wavbuf /*[B, 2S]*/, const float* chunk /*[B, S]*/, int B, int S, cudaStream_t stream);
int launch_stream_mel_cat(float* melcat /*[B, sub+n, F]*/, float* sub_cache /*[B, sub, F]*/, const float* mel_new /*[B, n, F]*/, int B, int sub,
                          int n, int F, cudaStream_t stream);
int launch_rows_cat(const float* a, int na, const float* b, int nb, float* out, int B, int D, cudaStream_t stream);
int launch_rows_slice(const float* src, int T, int t0, int n, float* dst, int B, int D, cudaStream_t stream);
struct CacheUpdateParams {
  const float* old_cache;   // [B, W, C]
  const float* cur;         // [B*Tc rows, cur_ld], columns [cur_col0, cur_col0 + C)
  float* new_cache;         // [B, W, C]
  int B, W, C, Tc, shift, cur_ld, cur_col0;   // shift = Tc - win_back
};
int launch_stream_cache_update(const CacheUpdateParams& p, cudaStream_t stream);
struct StreamAttnParams {
  const float* qkv;        // [B*Tc, 3*H*dh] new rows: q | k | v (q pre-scaled, biases added)
  const float* kv_cache;   // [B, W, 2*H*dh] right-aligned; the last c rows are valid
  float* out;              // [B*Tc, H*dh]
  int B, Tc, H, dh, W, c, win_front, win_back, round_tf32;
};
int launch_stream_attention(const StreamAttnParams& p, cudaStream_t stream);
struct StreamDwParams {
  const float* cache;   // [B, K-1, D] GLU rows of the previous frames (zeros = 'causal' padding)
  const float* cur;     // [B*Tc, D]
  const float* w;       // [K, D]
  float* y;             // [B*Tc, D]
  int B, Tc, D, K, round_tf32;
};
int launch_stream_dwconv(const StreamDwParams& p, cudaStream_t stream);
struct PickParams {
  const float* hidden;   // [B, T, D]
  const float* logits;   // [B, T, ldv] (first V columns are classes)
  float* feats;          // [B, T, D] out: kept rows first, zero rows behind
  float* picked;         // [B, T, V] out (nullable): the kept rows' logits
  int* counts;           // [B] out
  int* n_max;            // [1] out: max over the batch
  int B, T, D, V, ldv, blank;
};
int launch_feature_pick(const PickParams& p, cudaStream_t stream);
int launch_round_tf32(const float* src, float* dst, size_t n, cudaStream_t stream);

}  // namespace b200asr
