// b200asr engine: weight residency, workspace, and the launch schedule of the Conformer-CTC path behind the
// C ABI declared in include/b200asr.h.  No CPU fallback: creation fails without a usable CUDA device.
#include "engine_internal.cuh"

namespace b200asr {
thread_local char g_errbuf[512] = {0};
bool g_pdl_enabled = true;
}

namespace {


// ------------------------------------------------------------------------------------------------ workspace
struct Shapes {
  int B, L, T, pad_left, T1, T2, pt1, pf1, pt2, pf2, M;
};

Shapes shapes_for(b200asr_handle h, int B, int L) {
  Shapes s;
  const b200asr_config& c = h->cfg;
  s.B = B;
  s.L = L;
  SamePad fp = same_pad(L, c.n_dft, c.hop);
  s.T = fp.out;
  s.pad_left = fp.before;
  SamePad t1 = same_pad(s.T, 3, 2), f1 = same_pad(c.n_mels, 3, 2);
  s.T1 = t1.out; s.pt1 = t1.before; s.pf1 = f1.before;
  SamePad t2 = same_pad(s.T1, 3, 2), f2 = same_pad(h->F1, 3, 2);
  s.T2 = t2.out; s.pt2 = t2.before; s.pf2 = f2.before;
  s.M = B * s.T2;
  return s;
}


size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t carve(b200asr_handle h, const Shapes& s, Buffers* b, char* base) {
  const b200asr_config& c = h->cfg;
  size_t off = 0;
  auto take = [&](size_t nfloat) {
    char* p = base ? base + off : nullptr;
    off += align_up(nfloat * sizeof(float), 256);
    return reinterpret_cast<float*>(p);
  };
  const size_t D = c.dmodel, M = s.M;
  const size_t wide = (size_t)std::max(std::max(c.ff_dim, 3 * c.num_heads * c.head_size), 2 * c.dmodel);
  b->power = take((size_t)s.B * s.T * kPowerStride);
  b->mel = take((size_t)s.B * s.T * c.n_mels);
  b->c1 = take(h->use_fused_sub ? 0 : (size_t)s.B * s.T1 * h->F1 * D);   // conv1's map only exists on the unfused path
  b->c2 = take((size_t)s.B * s.T2 * h->F2 * D);
  b->x = take(M * D);
  b->xn = take(M * D);
  b->h = take(M * wide);
  b->att = take(M * D);
  b->g = take(M * D);
  b->logits = take(M * (size_t)std::max(c.vocab, 1));
  b->pmax = reinterpret_cast<unsigned int*>(take(s.B));
  b->am = reinterpret_cast<int*>(take(M));
  b->ids = reinterpret_cast<int*>(take(M));
  b->lens = reinterpret_cast<int*>(take(s.B));
  b->amp = reinterpret_cast<float2*>(take(M * 2 * (size_t)(std::max(c.vocab, 1) / 64 + 2)));
  return off;
}

int ensure_workspace(b200asr_handle h, const Shapes& s, Buffers* b) {
  if (h->chunk) return fail(h, "this handle is a ChunkConformer engine: use the b200asr_stream_* entry points");
  if (h->vad) return fail(h, "this handle is a voice-activity model: use b200asr_vad_infer");
  if (h->punc) return fail(h, "this handle is a punctuation model: use b200asr_punc_infer");
  size_t need = carve(h, s, b, nullptr);
  if (need > h->ws.bytes) {
    // growing the workspace invalidates captured graphs (they hold the old addresses)
    for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second.exec);
    h->graphs.clear();
    ENG_CUDA(h, cudaDeviceSynchronize());
    if (h->ws.base) ENG_CUDA(h, cudaFree(h->ws.base));
    h->ws.base = nullptr;
    h->ws.bytes = 0;
    ENG_CUDA(h, cudaMalloc(&h->ws.base, need));
    h->ws.bytes = need;
  }
  carve(h, s, b, reinterpret_cast<char*>(h->ws.base));
  return 0;
}


// One ConformerBlock with every LayerNorm folded into the epilogue of the GEMM that produces its input (11 launches).
// Pre-condition: b.xn == LN(b.x; w.ffn1.ln).  Post-condition: b.x = block output, b.xn = LN(b.x; *next_ln) if next_ln.
// round_out: the block output only feeds another GEMM (the CTC head): store it rounded to nearest tf32.
int run_block_fused(Ctx& c, const BlockW& w, const Buffers& b, int B, int T, int D, int F, int H, int dh, float eps,
                    const LNW* next_ln, bool round_out = false) {
  const int M = B * T, HD = H * dh;
  if (chain_resid_ln(c, b.xn, D, w.ffn1.w1, w.ffn1.b1, F, w.ffn1.w2, w.ffn1.b2, 0.5f, b, M, D, w.mhsa.ln, nullptr, eps)) return 1;
  if (gemm(c, b.xn, D, w.mhsa.wqkv, nullptr, nullptr, 0.f, b.h, 3 * HD, M, 3 * HD, D, EPI_NONE)) return 1;
  AttnParams ap{};
  ap.qkv = b.h; ap.out = b.att; ap.B = B; ap.T = T; ap.H = H; ap.dh = dh; ap.win_front = -1; ap.win_back = 0;
  ap.round_tf32 = 1;   // (this schedule only runs in tf32 precision)
  if (attention(c, ap)) return 1;
  if (gemm_resid_ln(c, b.att, HD, w.mhsa.wo, w.mhsa.bo, 1.0f, b, M, D, w.conv.ln, nullptr, eps)) return 1;
  if (gemm(c, b.xn, D, w.conv.pw1w, w.conv.pw1b, nullptr, 0.f, b.g, D, M, 2 * D, D, EPI_GLU)) return 1;
  DwConvParams dp{};
  dp.x = b.g; dp.w = w.conv.dww; dp.y = b.att; dp.B = B; dp.T = T; dp.D = D; dp.K = w.kernel_size;
  dp.pad_left = same_pad(T, w.kernel_size, 1).before;
  dp.round_tf32 = 1;
  c.h->launches++;
  if (launch_dwconv(dp, c.s)) return 1;
  if (chain_resid_ln(c, b.att, D, w.conv.pww, w.conv.pwb, 2 * D, w.conv.pw2w, w.conv.pw2b, 1.0f, b, M, D, w.ffn2.ln, nullptr, eps)) return 1;
  LNW none{nullptr, nullptr};
  if (chain_resid_ln(c, b.xn, D, w.ffn2.w1, w.ffn2.b1, F, w.ffn2.w2, w.ffn2.b2, 0.5f, b, M, D, w.ln, next_ln ? next_ln : &none, eps,
                     round_out && !next_ln)) return 1;
  return 0;
}

int run_block(Ctx& c, const BlockW& w, const Buffers& b, int B, int T, int D, int F, int H, int dh, float eps) {
  const int M = B * T;
  const FFNW* ff[2] = {&w.ffn1, &w.ffn2};
  auto ffn = [&](const FFNW& f) -> int {
    c.h->launches++;
    if (launch_layernorm(b.x, f.ln.g, f.ln.b, b.xn, M, D, eps, c.s)) return 1;
    if (gemm(c, b.xn, D, f.w1, f.b1, nullptr, 0.f, b.h, F, M, F, D, EPI_BIAS_SWISH)) return 1;
    if (gemm(c, b.h, F, f.w2, f.b2, b.x, 0.5f, b.x, D, M, D, F, EPI_RESID)) return 1;
    return 0;
  };
  if (ffn(*ff[0])) return 1;
  // MHSA
  c.h->launches++;
  if (launch_layernorm(b.x, w.mhsa.ln.g, w.mhsa.ln.b, b.xn, M, D, eps, c.s)) return 1;
  const int HD = H * dh;
  if (gemm(c, b.xn, D, w.mhsa.wqkv, nullptr, nullptr, 0.f, b.h, 3 * HD, M, 3 * HD, D, EPI_NONE)) return 1;
  AttnParams ap{};
  ap.qkv = b.h; ap.out = b.att; ap.B = B; ap.T = T; ap.H = H; ap.dh = dh; ap.win_front = -1; ap.win_back = 0;
  if (attention(c, ap)) return 1;
  if (gemm(c, b.att, HD, w.mhsa.wo, w.mhsa.bo, b.x, 1.0f, b.x, D, M, D, HD, EPI_RESID)) return 1;
  // conv module
  c.h->launches++;
  if (launch_layernorm(b.x, w.conv.ln.g, w.conv.ln.b, b.xn, M, D, eps, c.s)) return 1;
  if (gemm(c, b.xn, D, w.conv.pw1w, w.conv.pw1b, nullptr, 0.f, b.g, D, M, 2 * D, D, EPI_GLU)) return 1;
  DwConvParams dp{};
  dp.x = b.g; dp.w = w.conv.dww; dp.y = b.att; dp.B = B; dp.T = T; dp.D = D; dp.K = w.kernel_size;
  dp.pad_left = same_pad(T, w.kernel_size, 1).before;
  c.h->launches++;
  if (launch_dwconv(dp, c.s)) return 1;
  if (gemm(c, b.att, D, w.conv.pww, w.conv.pwb, nullptr, 0.f, b.h, 2 * D, M, 2 * D, D, EPI_BIAS_SWISH)) return 1;
  if (gemm(c, b.h, 2 * D, w.conv.pw2w, w.conv.pw2b, b.x, 1.0f, b.x, D, M, D, 2 * D, EPI_RESID)) return 1;
  if (ffn(*ff[1])) return 1;
  c.h->launches++;
  if (launch_layernorm(b.x, w.ln.g, w.ln.b, b.x, M, D, eps, c.s)) return 1;
  return 0;
}

// test hook (b200asr_debug_encode_taps): keep a copy of the residual stream [M, D]
int tap(Ctx& c, const float* x, size_t floats) {
  b200asr_handle h = c.h;
  if (!h->tap_dst || h->tap_count >= h->tap_max) return 0;
  B200_CUDA_OK(cudaMemcpyAsync(h->tap_dst + (size_t)h->tap_count * floats, x, floats * sizeof(float), cudaMemcpyDeviceToDevice, c.s));
  h->tap_count++;
  return 0;
}

int run_frontend(Ctx& c, const float* wav, const Shapes& s, const Buffers& b, float* mel_out) {
  b200asr_handle h = c.h;
  FrontendParams fp{};
  fp.wav = wav; fp.window = h->window; fp.twiddle = h->twiddle; fp.melw = h->melw; fp.mel_lo = h->mel_lo;
  fp.mel_hi = h->mel_hi; fp.mel_off = h->mel_off; fp.mel_wc = h->mel_wc; fp.mel_nnz = h->mel_nnz; fp.power = b.power; fp.pmax = b.pmax; fp.mel = mel_out; fp.B = s.B; fp.L = s.L; fp.T = s.T;
  fp.pad_left = s.pad_left; fp.hop = h->cfg.hop; fp.power_stride = kPowerStride; fp.n_mels = h->cfg.n_mels; fp.mode = 0;
  h->launches += 3;
  return launch_frontend(fp, c.s);
}

int run_subsample_convs(Ctx& c, const float* mel, const Shapes& s, const Buffers& b);
int run_encoder_tail(Ctx& c, const Shapes& s, const Buffers& b);

// wav [B, L] -> b.x [B*T2, D]
int run_encoder(Ctx& c, const float* wav, const Shapes& s, const Buffers& b) {
  b200asr_handle h = c.h;
  const b200asr_config& cfg = h->cfg;
  const int D = cfg.dmodel;
  if (run_frontend(c, wav, s, b, b.mel)) return 1;
  if (run_subsample_convs(c, b.mel, s, b)) return 1;
  return run_encoder_tail(c, s, b);
}

// the subsampling linear layer as an fp16-operand GEMM with fused LayerNorm (conv2 then writes fp16)
GemmParams sublin_params(b200asr_handle h, const Shapes& s, const Buffers& b, bool f16) {
  const b200asr_config& cfg = h->cfg;
  const int D = cfg.dmodel;
  GemmParams lp{};
  lp.A = b.c2; lp.W = f16 ? h->linw16 : h->linw; lp.bias = h->linb; lp.C = b.x; lp.C2 = b.xn; lp.M = s.M; lp.N = D; lp.K = h->F2 * D;
  lp.lda = h->F2 * D; lp.ldc = D; lp.ln_eps = cfg.ln_eps; lp.f16 = f16 ? 1 : 0;
  if (!h->enc_blocks.empty()) { lp.ln1_g = h->enc_blocks[0].ffn1.ln.g; lp.ln1_b = h->enc_blocks[0].ffn1.ln.b; }
  return lp;
}
GemmParams conv2_params(b200asr_handle h, const Shapes& s, const Buffers& b) {
  const int D = h->cfg.dmodel;
  GemmParams g{};
  g.A = b.c1; g.W = h->c2w; g.bias = h->c2b; g.C = b.c2; g.M = s.B * s.T2 * h->F2; g.N = D; g.K = 9 * D; g.lda = 0; g.ldc = D;
  g.a_mode = 1; g.T1 = s.T1; g.F1 = h->F1; g.T2 = s.T2; g.F2 = h->F2; g.D = D; g.pad_t = s.pt2; g.pad_f = s.pf2;
  g.round_out = 1;   // conv2's output is only read by the subsampling linear layer's GEMM (ignored by the fp32 kernel)
  return g;
}
// How the two-kernel subsampler runs (one decision, taken the same way by every stage that touches b.c1 / b.c2):
//   0  fp32 maps (exact-fp32 mode, or fp16 switched off)
//   1  conv1 map in fp16, conv2 kind::f16, conv2's output fp32
//   2  ... and conv2's output in fp16, read by the subsampling linear layer as an fp16-operand GEMM with fused LayerNorm
int sub_f16_mode(b200asr_handle h, const Shapes& s, const Buffers& b) {
  if (!h->conv_f16 || h->cfg.precision != B200ASR_PRECISION_TF32 || h->use_fused_sub) return 0;
  GemmParams g = conv2_params(h, s, b);
  g.f16 = 1; g.W = h->c2w16;
  if (!tc_gemm_supported(g, EPI_BIAS_RELU)) return 0;
  if (!h->sub_out_f16 || !fused_ln_ok(h) || h->enc_blocks.empty()) return 1;
  g.out_f16 = 1;
  if (!tc_gemm_supported(g, EPI_BIAS_RELU) || !tc_gemm_supported(sublin_params(h, s, b, true), EPI_BIAS_LN)) return 1;
  return 2;
}

// mel [B, T, n_mels] -> b.c2 = relu(conv2(relu(conv1(mel))))  [B, T2, F2, D]  (fp16 in sub_f16_mode 2, else fp32)
int run_subsample_convs(Ctx& c, const float* mel, const Shapes& s, const Buffers& b) {
  b200asr_handle h = c.h;
  const b200asr_config& cfg = h->cfg;
  const int D = cfg.dmodel;
  if (h->use_fused_sub) {
    ConvSubParams cp{};
    cp.mel = mel; cp.w1 = h->c1w; cp.b1 = h->c1b; cp.w2 = h->c2w; cp.b2 = h->c2b; cp.out = b.c2;
    cp.B = s.B; cp.T = s.T; cp.F = cfg.n_mels; cp.T1 = s.T1; cp.F1 = h->F1; cp.T2 = s.T2; cp.F2 = h->F2; cp.D = D;
    cp.pt1 = s.pt1; cp.pf1 = s.pf1; cp.pt2 = s.pt2; cp.pf2 = s.pf2; cp.round_out = 1;
    h->launches++;
    return launch_conv_subsample_tc(h->tc, cp, c.s);
  }
  Conv1Params c1{};
  c1.mel = mel; c1.w = h->c1w; c1.bias = h->c1b; c1.out = b.c1; c1.B = s.B; c1.T = s.T; c1.F = cfg.n_mels; c1.T1 = s.T1;
  c1.F1 = h->F1; c1.D = D; c1.pad_t = s.pt1; c1.pad_f = s.pf1;
  c1.round_tf32 = (cfg.precision == B200ASR_PRECISION_TF32) ? 1 : 0;
  GemmParams g = conv2_params(h, s, b);
  const int mode = sub_f16_mode(h, s, b);
  if (mode >= 1) { c1.out_f16 = 1; g.f16 = 1; g.W = h->c2w16; }
  if (mode == 2) g.out_f16 = 1;
  h->launches++;
  if (launch_conv1(c1, c.s)) return 1;
  h->launches++;
  if (cfg.precision == B200ASR_PRECISION_TF32 && tc_gemm_supported(g, EPI_BIAS_RELU)) {
    if (launch_gemm_tc(h->tc, g, EPI_BIAS_RELU, c.s)) return 1;
  } else {
    if (launch_gemm_simt(g, EPI_BIAS_RELU, c.s)) return 1;
  }
  return 0;
}

// b.c2 -> subsampling linear layer -> encoder blocks -> b.x
int run_encoder_tail(Ctx& c, const Shapes& s, const Buffers& b) {
  b200asr_handle h = c.h;
  const b200asr_config& cfg = h->cfg;
  const int D = cfg.dmodel;
  if (fused_ln_ok(h) && !h->enc_blocks.empty()) {
    // (the same predicate run_subsample_convs used when it chose conv2's output type)
    GemmParams lp = sublin_params(h, s, b, sub_f16_mode(h, s, b) == 2);
    if (gemm_p(c, lp, EPI_BIAS_LN)) return 1;
    if (tap(c, b.x, (size_t)s.M * D)) return 1;
    for (size_t i = 0; i < h->enc_blocks.size(); ++i) {
      const LNW* next = (i + 1 < h->enc_blocks.size()) ? &h->enc_blocks[i + 1].ffn1.ln : nullptr;
      if (run_block_fused(c, h->enc_blocks[i], b, s.B, s.T2, D, cfg.ff_dim, cfg.num_heads, cfg.head_size, cfg.ln_eps, next)) return 1;
      if (tap(c, b.x, (size_t)s.M * D)) return 1;
    }
    return 0;
  }
  if (gemm(c, b.c2, h->F2 * D, h->linw, h->linb, nullptr, 0.f, b.x, D, s.M, D, h->F2 * D, EPI_BIAS)) return 1;
  if (tap(c, b.x, (size_t)s.M * D)) return 1;
  for (const BlockW& w : h->enc_blocks) {
    if (run_block(c, w, b, s.B, s.T2, D, cfg.ff_dim, cfg.num_heads, cfg.head_size, cfg.ln_eps)) return 1;
    if (tap(c, b.x, (size_t)s.M * D)) return 1;
  }
  return 0;
}

// enc [B*Tp, D] -> logits [B*Tp, V]   (uses b.x as the running activation; enc may alias b.x's source)
int run_ctc(Ctx& c, const float* enc, int B, int Tp, const Buffers& b, float* logits) {
  b200asr_handle h = c.h;
  const b200asr_config& cfg = h->cfg;
  const int D = cfg.dmodel, M = B * Tp;
  if (fused_ln_ok(h) && !h->ctc_blocks.empty()) {
    // projection -> (x', LN(x'; blk0.ffn1.ln)); x' must not alias the input, so the stream moves to b.g when enc == b.x
    Buffers bb = b;
    if (enc == b.x) std::swap(bb.x, bb.g);
    GemmParams pp{};
    pp.A = enc; pp.W = h->ctc_projw; pp.bias = h->ctc_projb; pp.C = bb.x; pp.C2 = bb.xn; pp.M = M; pp.N = D; pp.K = D; pp.lda = D;
    pp.ldc = D; pp.ln1_g = h->ctc_blocks[0].ffn1.ln.g; pp.ln1_b = h->ctc_blocks[0].ffn1.ln.b; pp.ln_eps = cfg.ln_eps;
    if (gemm_p(c, pp, EPI_BIAS_LN)) return 1;
    for (size_t i = 0; i < h->ctc_blocks.size(); ++i) {
      const LNW* next = (i + 1 < h->ctc_blocks.size()) ? &h->ctc_blocks[i + 1].ffn1.ln : nullptr;
      if (run_block_fused(c, h->ctc_blocks[i], bb, B, Tp, D, cfg.ff_dim, cfg.num_heads, cfg.head_size, cfg.ln_eps, next, true)) return 1;
    }
    if (logits == nullptr) {
      // greedy path: the CTC head keeps only per-tile (max, argmax) pairs (EPI_BIAS_ARGMAX); no logits are written
      GemmParams fp{};
      fp.A = bb.x; fp.W = h->ctc_fcw; fp.bias = h->ctc_fcb; fp.C = reinterpret_cast<float*>(b.amp); fp.M = M; fp.N = cfg.vocab; fp.K = D;
      fp.lda = D; fp.ldc = cfg.vocab;
      return gemm_p(c, fp, EPI_BIAS_ARGMAX);
    }
    return gemm(c, bb.x, D, h->ctc_fcw, h->ctc_fcb, nullptr, 0.f, logits, cfg.vocab, M, cfg.vocab, D, EPI_BIAS);
  }
  // the block schedule runs in place on its `x` buffer; when the input *is* b.x, project into b.xn and swap roles
  Buffers bb = b;
  if (enc == b.x) std::swap(bb.x, bb.xn);
  if (gemm(c, enc, D, h->ctc_projw, h->ctc_projb, nullptr, 0.f, bb.x, D, M, D, D, EPI_BIAS)) return 1;
  for (const BlockW& w : h->ctc_blocks)
    if (run_block(c, w, bb, B, Tp, D, cfg.ff_dim, cfg.num_heads, cfg.head_size, cfg.ln_eps)) return 1;
  if (gemm(c, bb.x, D, h->ctc_fcw, h->ctc_fcb, nullptr, 0.f, logits, cfg.vocab, M, cfg.vocab, D, EPI_BIAS)) return 1;
  return 0;
}

// Block streaming (StreamingConformerEncoder.call, conformer_blocks.py:574-594): an utterance longer than one chunk is reshaped
// into independent chunks -- tf.reshape there raises unless L is a whole number of chunks, and so does this (encoding such an
// utterance as one block would silently be a different model: global attention, one dB maximum).  L <= chunk is one (short) chunk.
int effective_batch(b200asr_handle h, int* B, int* L) {
  const int cs = h->cfg.chunk_samples;
  if (cs > 0 && *L > cs) {
    if ((*L % cs) != 0) {
      snprintf(g_errbuf, sizeof(g_errbuf), "streaming engine: %d samples is not a whole number of %d-sample chunks (split the ragged tail "
               "off and pass it on its own, as test_asr.py:120-128 does)", *L, cs);
      if (h) h->err = g_errbuf;
      return 1;
    }
    *B = *B * (*L / cs);
    *L = cs;
  }
  return 0;
}


}  // namespace

// ================================================================================================== shared with chunk_engine.cu
namespace b200asr {

// device / blob checks, new handle with the weight blob resident on `device` and the tensor table filled.  h->cfg is left to the caller.
int engine_alloc(const void* weight_blob, size_t blob_bytes, int device, const char* who, b200asr_engine** out) {
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    snprintf(g_errbuf, sizeof(g_errbuf), "%s: no CUDA device (this library has no CPU fallback)", who);
    return 1;
  }
  if (device < 0 || device >= ndev) { snprintf(g_errbuf, sizeof(g_errbuf), "%s: bad device index", who); return 1; }
  cudaDeviceProp prop;
  ENG_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    snprintf(g_errbuf, sizeof(g_errbuf), "%s: device %d is sm_%d%d; this build targets sm_100a (B200) only", who, device, prop.major, prop.minor);
    return 1;
  }
  ENG_CUDA(nullptr, cudaSetDevice(device));
  if (blob_bytes < 16 || memcmp(weight_blob, "B2ASRW01", 8) != 0) { snprintf(g_errbuf, sizeof(g_errbuf), "%s: bad weight blob magic", who); return 1; }
  const char* hb = static_cast<const char*>(weight_blob);
  uint32_t n_entries;
  memcpy(&n_entries, hb + 8, 4);
  const size_t table_end = 16 + (size_t)n_entries * sizeof(BlobEntry);
  if (table_end > blob_bytes) { snprintf(g_errbuf, sizeof(g_errbuf), "%s: truncated weight blob", who); return 1; }
  b200asr_engine* h = new b200asr_engine();
  h->device = device;
  if (cudaMalloc(&h->blob_dev, blob_bytes) != cudaSuccess ||
      cudaMemcpy(h->blob_dev, weight_blob, blob_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    snprintf(g_errbuf, sizeof(g_errbuf), "%s: cannot place %zu weight bytes on device", who, blob_bytes);
    b200asr_destroy(h);
    return 1;
  }
  for (uint32_t i = 0; i < n_entries; ++i) {
    BlobEntry e;
    memcpy(&e, hb + 16 + (size_t)i * sizeof(BlobEntry), sizeof(BlobEntry));
    e.name[47] = 0;
    if (e.offset % 128 != 0 || e.offset + e.numel * 4 > blob_bytes) {
      snprintf(g_errbuf, sizeof(g_errbuf), "%s: entry '%s' out of bounds / misaligned", who, e.name);
      std::string keep = g_errbuf;
      b200asr_destroy(h);
      snprintf(g_errbuf, sizeof(g_errbuf), "%s", keep.c_str());
      return 1;
    }
    h->tensors[e.name] = {reinterpret_cast<const float*>(h->blob_dev + e.offset), e.numel};
  }
  *out = h;
  return 0;
}

// frontend tables (window, FFT twiddles, sparse mel filters), the tensor-map encoder and the environment switches.
// Needs h->cfg.{n_dft, n_mels} and the host copy of the blob (the mel filters' sparsity pattern is analysed on the host).
int engine_init_frontend(b200asr_engine* h, const void* weight_blob) {
  const b200asr_config& c = h->cfg;
  const char* hb = static_cast<const char*>(weight_blob);
  uint32_t n_entries;
  memcpy(&n_entries, hb + 8, 4);
  const int nb = c.n_dft / 2 + 1;
  bool ok = true;
  h->window = lookup(h, "fe.window", c.n_dft, &ok);
  h->melw = ok ? lookup(h, "fe.mel", (uint64_t)nb * c.n_mels, &ok) : nullptr;
  if (!ok) return 1;
  // FFT twiddles exp(-2 pi i m / 1024), rounded once from double
  std::vector<float2> tw(c.n_dft);
  for (int m = 0; m < c.n_dft; ++m) {
    const double a = -2.0 * M_PI * (double)m / (double)c.n_dft;
    tw[m] = make_float2((float)cos(a), (float)sin(a));
  }
  // sparse extent of each mel filter (the reference multiplies by the dense matrix; zeros contribute nothing)
  const float* melw_host = nullptr;
  for (uint32_t i = 0; i < n_entries; ++i) {
    BlobEntry e;
    memcpy(&e, hb + 16 + (size_t)i * sizeof(BlobEntry), sizeof(BlobEntry));
    e.name[47] = 0;
    if (strcmp(e.name, "fe.mel") == 0) melw_host = reinterpret_cast<const float*>(hb + e.offset);
  }
  std::vector<int> lo(c.n_mels, 0), hi(c.n_mels, 0);
  for (int m = 0; m < c.n_mels; ++m) {
    int l = nb, r = 0;
    for (int k = 0; k < nb; ++k)
      if (melw_host[(size_t)k * c.n_mels + m] != 0.0f) {
        l = std::min(l, k);
        r = std::max(r, k + 1);
      }
    if (l >= r) l = r = 0;
    lo[m] = l;
    hi[m] = r;
  }
  // compact band weights of the (sparse, triangular) mel filters, staged in shared memory by db_mel_kernel
  std::vector<int> off(c.n_mels + 1, 0);
  for (int m = 0; m < c.n_mels; ++m) off[m + 1] = off[m] + (hi[m] - lo[m]);
  std::vector<float> wc(std::max(off[c.n_mels], 1), 0.f);
  for (int m = 0; m < c.n_mels; ++m)
    for (int k = lo[m]; k < hi[m]; ++k) wc[off[m] + k - lo[m]] = melw_host[(size_t)k * c.n_mels + m];
  h->mel_nnz = off[c.n_mels];
  if (cudaMalloc(&h->mel_off, sizeof(int) * (c.n_mels + 1)) != cudaSuccess ||
      cudaMalloc(&h->mel_wc, sizeof(float) * wc.size()) != cudaSuccess ||
      cudaMemcpy(h->mel_off, off.data(), sizeof(int) * (c.n_mels + 1), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(h->mel_wc, wc.data(), sizeof(float) * wc.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    snprintf(g_errbuf, sizeof(g_errbuf), "engine: device allocation of the mel tables failed");
    return 1;
  }
  if (cudaMalloc(&h->twiddle, sizeof(float2) * c.n_dft) != cudaSuccess ||
      cudaMalloc(&h->mel_lo, sizeof(int) * c.n_mels) != cudaSuccess ||
      cudaMalloc(&h->mel_hi, sizeof(int) * c.n_mels) != cudaSuccess ||
      cudaMemcpy(h->twiddle, tw.data(), sizeof(float2) * c.n_dft, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(h->mel_lo, lo.data(), sizeof(int) * c.n_mels, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(h->mel_hi, hi.data(), sizeof(int) * c.n_mels, cudaMemcpyHostToDevice) != cudaSuccess) {
    snprintf(g_errbuf, sizeof(g_errbuf), "engine: device allocation of frontend tables failed");
    return 1;
  }
  if (tc_init(&h->tc) != 0) return 1;
  if (const char* e = getenv("B200ASR_NO_CHAIN")) h->use_chain = !(e[0] == '1');
  if (const char* e = getenv("B200ASR_NO_PDL")) g_pdl_enabled = !(e[0] == '1');
  if (const char* e = getenv("B200ASR_NO_PAIR")) h->use_pair = !(e[0] == '1');
  if (const char* e = getenv("B200ASR_ATTN_ASYNC")) h->attn_async = (e[0] == '1');
  return 0;
}

}  // namespace b200asr

// ================================================================================================== C ABI
extern "C" {

B200ASR_API int b200asr_abi_version(void) { return B200ASR_ABI_VERSION; }

B200ASR_API const char* b200asr_last_error(b200asr_handle h) { return h ? h->err.c_str() : g_errbuf; }

B200ASR_API int b200asr_create(const void* weight_blob, size_t blob_bytes, const b200asr_config* cfg, int device, b200asr_handle* out) {
  if (!weight_blob || !cfg || !out) return fail(nullptr, "b200asr_create: null argument");
  if (cfg->abi_version != B200ASR_ABI_VERSION) return fail(nullptr, "b200asr_create: ABI version mismatch");
  *out = nullptr;
  b200asr_engine* h = nullptr;
  if (b200asr::engine_alloc(weight_blob, blob_bytes, device, "b200asr_create", &h)) return 1;
  h->cfg = *cfg;
  auto bail = [&](int) {
    std::string e = g_errbuf;
    b200asr_destroy(h);
    snprintf(g_errbuf, sizeof(g_errbuf), "%s", e.c_str());
    return 1;
  };
  const b200asr_config& c = h->cfg;
  if (c.n_dft != 1024) { snprintf(g_errbuf, sizeof(g_errbuf), "b200asr_create: n_dft must be 1024 (reference hard-codes it)"); return bail(1); }
  if (c.dmodel % 16 != 0 || c.dmodel > 512) { snprintf(g_errbuf, sizeof(g_errbuf), "b200asr_create: dmodel must be a multiple of 16, <= 512"); return bail(1); }
  const int D = c.dmodel;
  h->F1 = same_pad(c.n_mels, 3, 2).out;
  h->F2 = same_pad(h->F1, 3, 2).out;
  bool ok = true;
  h->c1w = lookup(h, "sub.conv1.w", 9ull * D, &ok);
  h->c1b = ok ? lookup(h, "sub.conv1.b", D, &ok) : nullptr;
  h->c2w = ok ? lookup(h, "sub.conv2.w", 9ull * D * D, &ok) : nullptr;
  h->c2b = ok ? lookup(h, "sub.conv2.b", D, &ok) : nullptr;
  h->linw = ok ? lookup(h, "sub.lin.w", (uint64_t)h->F2 * D * D, &ok) : nullptr;
  h->linb = ok ? lookup(h, "sub.lin.b", D, &ok) : nullptr;
  if (!ok) return bail(1);
  {
    auto it = h->tensors.find("sub.conv2.w16");
    if (it != h->tensors.end() && it->second.second == 9ull * D * D / 2) h->c2w16 = it->second.first;
    const char* off = getenv("B200ASR_NO_CONV_F16");
    h->conv_f16 = h->c2w16 != nullptr && c.precision == B200ASR_PRECISION_TF32 && D % 8 == 0 && !(off && off[0] == '1');
    auto il = h->tensors.find("sub.lin.w16");
    if (il != h->tensors.end() && il->second.second == (uint64_t)h->F2 * D * D / 2) h->linw16 = il->second.first;
    h->sub_out_f16 = h->conv_f16 && h->linw16 != nullptr && !(off && off[0] == '2');   // (needs the fused-LayerNorm GEMM too: checked where it is used)
  }
  h->enc_blocks.resize(c.num_blocks);
  for (int i = 0; i < c.num_blocks; ++i)
    if (!load_block(h, "enc." + std::to_string(i) + ".", D, c.ff_dim, c.num_heads, c.head_size, c.kernel_size, &h->enc_blocks[i]))
      return bail(1);
  if (c.vocab > 0) {
    h->ctc_projw = lookup(h, "ctc.proj.w", (uint64_t)D * D, &ok);
    h->ctc_projb = ok ? lookup(h, "ctc.proj.b", D, &ok) : nullptr;
    h->ctc_fcw = ok ? lookup(h, "ctc.fc.w", (uint64_t)c.vocab * D, &ok) : nullptr;
    h->ctc_fcb = ok ? lookup(h, "ctc.fc.b", c.vocab, &ok) : nullptr;
    if (!ok) return bail(1);
    h->ctc_blocks.resize(c.ctc_blocks);
    for (int i = 0; i < c.ctc_blocks; ++i)
      if (!load_block(h, "ctc.blk" + std::to_string(i) + ".", D, c.ff_dim, c.num_heads, c.head_size, c.ctc_kernel_size,
                      &h->ctc_blocks[i]))
        return bail(1);
  }
  if (c.tr_blocks > 0) {
    h->tr_emb = lookup(h, "tr.emb", (uint64_t)c.tr_inp_classes * D, &ok);
    h->tr_fcw = ok ? lookup(h, "tr.fc.w", (uint64_t)c.tr_vocab * D, &ok) : nullptr;
    h->tr_fcb = ok ? lookup(h, "tr.fc.b", c.tr_vocab, &ok) : nullptr;
    if (!ok) return bail(1);
    auto it = h->tensors.find("tr.pe");
    if (it == h->tensors.end() || it->second.second % D != 0) { snprintf(g_errbuf, sizeof(g_errbuf), "b200asr_create: tr.pe (positional table) missing"); return bail(1); }
    h->tr_pe = it->second.first;
    h->tr_pe_rows = (int)(it->second.second / D);
    h->tr_blocks.resize(c.tr_blocks);
    for (int i = 0; i < c.tr_blocks; ++i)
      if (!load_block(h, "tr." + std::to_string(i) + ".", D, c.ff_dim, c.num_heads, c.head_size, c.tr_kernel_size, &h->tr_blocks[i])) return bail(1);
  }
  if (b200asr::engine_init_frontend(h, weight_blob)) return bail(1);
  {
    ConvSubParams probe{};
    probe.D = D; probe.F2 = h->F2; probe.B = 1; probe.T = 1; probe.w2 = h->c2w;
    // The fused conv1 -> conv2 kernel removes the conv1 map from HBM (47 MB of DRAM traffic instead of 764 MB per 32 x 10 s batch) but is
    // SLOWER than the two-kernel path on a B200 (340 us against 262 us: its producer warps recompute conv1 2.25x on CUDA cores and stall
    // between K blocks, profiles/r02_fused_subsampler.md), and HBM is nowhere near the step's bottleneck: opt-in (B200ASR_FUSED_SUB=1).
    const char* fe = getenv("B200ASR_FUSED_SUB");
    h->use_fused_sub = fe && fe[0] == '1' && c.precision == B200ASR_PRECISION_TF32 && h->tc.ready && conv_subsample_tc_supported(probe);
  }
  *out = h;
  return 0;
}

B200ASR_API int b200asr_destroy(b200asr_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second.exec);
  if (h->chunk) b200asr::chunk_model_free(h->chunk);
  if (h->vad) b200asr::vad_model_free(h->vad);
  if (h->punc) b200asr::punc_model_free(h->punc);
  if (h->stage_wav) cudaFree(h->stage_wav);
  if (h->tr_ws) cudaFree(h->tr_ws);
  if (h->blob_dev) cudaFree(h->blob_dev);
  if (h->twiddle) cudaFree(h->twiddle);
  if (h->mel_lo) cudaFree(h->mel_lo);
  if (h->mel_hi) cudaFree(h->mel_hi);
  if (h->mel_off) cudaFree(h->mel_off);
  if (h->mel_wc) cudaFree(h->mel_wc);
  if (h->ws.base) cudaFree(h->ws.base);
  if (h->beam_ws) cudaFree(h->beam_ws);
  for (auto& ps : h->pipe) {
    if (ps.wav) cudaFree(ps.wav);
    if (ps.ids) cudaFree(ps.ids);
    if (ps.lens) cudaFree(ps.lens);
    if (ps.h2d) cudaEventDestroy(ps.h2d);
    if (ps.done) cudaEventDestroy(ps.done);
  }
  if (h->pipe_copy) cudaStreamDestroy(h->pipe_copy);
  if (h->pipe_compute) cudaStreamDestroy(h->pipe_compute);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->ev_in) cudaEventDestroy(h->ev_in);
  if (h->ev_out) cudaEventDestroy(h->ev_out);
  delete h;
  return 0;
}

B200ASR_API int b200asr_out_frames(b200asr_handle h, int num_samples) {
  if (!h || num_samples <= 0) return 0;
  int B = 1, L = num_samples;
  if (effective_batch(h, &B, &L)) return 0;
  return shapes_for(h, 1, L).T2 * B;
}

B200ASR_API int b200asr_reserve(b200asr_handle h, int B, int L) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (B <= 0 || L <= 0) return fail(h, "b200asr_reserve: B and L must be positive");
  if (effective_batch(h, &B, &L)) return 1;
  Buffers b;
  Shapes s = shapes_for(h, B, L);
  return ensure_workspace(h, s, &b);
}

B200ASR_API int64_t b200asr_launch_count(b200asr_handle h) { return h ? h->launches : 0; }

B200ASR_API int b200asr_mel(b200asr_handle h, const float* wav_dev, int B, int L, float* mel_dev, void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (B < 0 || L <= 0 || !wav_dev || !mel_dev) return fail(h, "b200asr_mel: bad arguments");
  if (B == 0) return 0;
  if (effective_batch(h, &B, &L)) return 1;
  Shapes s = shapes_for(h, B, L);
  Buffers b;
  if (ensure_workspace(h, s, &b)) return 1;
  Ctx c{h, static_cast<cudaStream_t>(stream)};
  ENG_TRY(h, run_frontend(c, wav_dev, s, b, mel_dev));
  return 0;
}

B200ASR_API int b200asr_encode(b200asr_handle h, const float* wav_dev, int B, int L, float* enc_dev, void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (B < 0 || L <= 0 || !wav_dev || !enc_dev) return fail(h, "b200asr_encode: bad arguments");
  if (B == 0) return 0;
  if (effective_batch(h, &B, &L)) return 1;
  Shapes s = shapes_for(h, B, L);
  Buffers b;
  if (ensure_workspace(h, s, &b)) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GraphKey key{};
  key.kind = 1; key.B = B; key.L = L; key.p0 = wav_dev; key.p1 = enc_dev;
  return with_graph(h, st, key, [&](cudaStream_t st) -> int {
    Ctx c{h, st};
    ENG_TRY(h, run_encoder(c, wav_dev, s, b));
    ENG_CUDA(h, cudaMemcpyAsync(enc_dev, b.x, sizeof(float) * (size_t)s.M * h->cfg.dmodel, cudaMemcpyDeviceToDevice, st));
    return 0;
  });
}

// Test hook: the encoder without CUDA graph, keeping a copy of the residual stream after the subsampler (tap 0) and after each
// block (taps 1 .. num_blocks) in taps_dev [n_taps][B*T', D]  (per-stage parity table against the reference's ONNX taps).
B200ASR_API int b200asr_debug_encode_taps(b200asr_handle h, const float* wav_dev, int B, int L, float* taps_dev, int n_taps, void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (B <= 0 || L <= 0 || !wav_dev || !taps_dev || n_taps <= 0) return fail(h, "b200asr_debug_encode_taps: bad arguments");
  if (effective_batch(h, &B, &L)) return 1;
  Shapes s = shapes_for(h, B, L);
  Buffers b;
  if (ensure_workspace(h, s, &b)) return 1;
  Ctx c{h, static_cast<cudaStream_t>(stream)};
  h->tap_dst = taps_dev; h->tap_count = 0; h->tap_max = n_taps;
  const int rc = run_encoder(c, wav_dev, s, b);
  h->tap_dst = nullptr;
  if (rc) return fail_cuda(h);
  return 0;
}

B200ASR_API int b200asr_ctc_logits(b200asr_handle h, const float* enc_dev, int B, int Tp, float* logits_dev, void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (h->cfg.vocab <= 0) return fail(h, "b200asr_ctc_logits: engine was created without a CTC decoder");
  if (B < 0 || Tp < 0 || !enc_dev || !logits_dev) return fail(h, "b200asr_ctc_logits: bad arguments");
  if (B == 0 || Tp == 0) return 0;
  // workspace sized through an equivalent (B, L): T2 = Tp  <=  L = Tp * 4 * hop
  Shapes s = shapes_for(h, B, Tp * 4 * h->cfg.hop);
  if (s.T2 != Tp) return fail(h, "b200asr_ctc_logits: internal shape inference mismatch");
  Buffers b;
  if (ensure_workspace(h, s, &b)) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GraphKey key{};
  key.kind = 2; key.B = B; key.L = Tp; key.p0 = enc_dev; key.p1 = logits_dev;
  return with_graph(h, st, key, [&](cudaStream_t st) -> int {
    Ctx c{h, st};
    ENG_TRY(h, run_ctc(c, enc_dev, B, Tp, b, logits_dev));
    return 0;
  });
}

B200ASR_API int b200asr_ctc_greedy(b200asr_handle h, const float* logits_dev, const int32_t* lengths_dev, int B, int Tp, int V, int blank,
                       int32_t* ids_dev, int32_t* out_len_dev, void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (B < 0 || Tp < 0 || V <= 0 || !ids_dev || !out_len_dev || (!logits_dev && B * Tp > 0))
    return fail(h, "b200asr_ctc_greedy: bad arguments");
  if (B == 0) return 0;
  Shapes s = shapes_for(h, B, std::max(Tp, 1) * 4 * h->cfg.hop);
  Buffers b;
  if (ensure_workspace(h, s, &b)) return 1;
  h->launches += 2;
  ENG_TRY(h, launch_ctc_greedy(logits_dev, lengths_dev, B, Tp, V, blank, b.am, ids_dev, out_len_dev,
                               static_cast<cudaStream_t>(stream)));
  return 0;
}

static int ctc_beam_impl(b200asr_handle h, const float* logits_dev, const int32_t* lengths_dev, int B, int Tp, int V, int blank, int beam,
                         int cutoff_top_n, float cutoff_prob, int32_t* ids_dev, int32_t* out_len_dev, float* scores_dev, void* stream, int is_prob);

B200ASR_API int b200asr_ctc_beam(b200asr_handle h, const float* logits_dev, const int32_t* lengths_dev, int B, int Tp, int V, int blank,
                     int beam, int cutoff_top_n, float cutoff_prob, int32_t* ids_dev, int32_t* out_len_dev, float* scores_dev,
                     void* stream) {
  return ctc_beam_impl(h, logits_dev, lengths_dev, B, Tp, V, blank, beam, cutoff_top_n, cutoff_prob, ids_dev, out_len_dev, scores_dev, stream, 0);
}

B200ASR_API int b200asr_ctc_beam_probs(b200asr_handle h, const float* probs_dev, const int32_t* lengths_dev, int B, int Tp, int V, int blank,
                           int beam, int cutoff_top_n, float cutoff_prob, int32_t* ids_dev, int32_t* out_len_dev, float* scores_dev,
                           void* stream) {
  return ctc_beam_impl(h, probs_dev, lengths_dev, B, Tp, V, blank, beam, cutoff_top_n, cutoff_prob, ids_dev, out_len_dev, scores_dev, stream, 1);
}

static int ctc_beam_impl(b200asr_handle h, const float* logits_dev, const int32_t* lengths_dev, int B, int Tp, int V, int blank, int beam,
                         int cutoff_top_n, float cutoff_prob, int32_t* ids_dev, int32_t* out_len_dev, float* scores_dev, void* stream, int is_prob) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (B < 0 || Tp < 0 || V <= 0 || !ids_dev || !out_len_dev || !scores_dev || (!logits_dev && B * Tp > 0))
    return fail(h, "b200asr_ctc_beam: bad arguments");
  if (B == 0) return 0;
  const size_t need = beam_workspace_bytes(B, Tp, std::max(beam, 1));
  if (need > h->beam_ws_bytes) {
    ENG_CUDA(h, cudaDeviceSynchronize());
    if (h->beam_ws) ENG_CUDA(h, cudaFree(h->beam_ws));
    h->beam_ws = nullptr;
    h->beam_ws_bytes = 0;
    ENG_CUDA(h, cudaMalloc(&h->beam_ws, need));
    h->beam_ws_bytes = need;
  }
  BeamParams p{};
  p.logits = logits_dev; p.lengths = lengths_dev; p.B = B; p.T = Tp; p.V = V; p.blank = blank; p.beam = beam;
  p.cutoff_top_n = cutoff_top_n; p.cutoff_prob = cutoff_prob; p.ids = ids_dev; p.out_len = out_len_dev; p.scores = scores_dev;
  p.workspace = h->beam_ws;
  p.is_prob = is_prob;
  h->launches += 2;
  ENG_TRY(h, launch_ctc_beam(p, static_cast<cudaStream_t>(stream)));
  return 0;
}

// Translator (conformer_blocks.py:504-552).  The sequences are tiny (U = tokens + 10, two blocks): the schedule is the plain one
// (LayerNorm kernels + GEMMs through whichever arithmetic the engine runs in), every GEMM on the same kernels as the encoder.
B200ASR_API int b200asr_translate(b200asr_handle h, const int32_t* ids_dev, const float* enc_dev, int B, int U, int Tp, float* logits_dev,
                                  void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (h->tr_blocks.empty()) return fail(h, "b200asr_translate: the engine was created without a translator (tr_blocks = 0)");
  if (!ids_dev || !enc_dev || !logits_dev || B < 0 || U < 0 || Tp <= 0) return fail(h, "b200asr_translate: bad arguments");
  if (B == 0 || U == 0) return 0;
  if (U > h->tr_pe_rows) return fail(h, "b200asr_translate: sequence longer than the positional table in the blob");
  const b200asr_config& cfg = h->cfg;
  const int D = cfg.dmodel, F = cfg.ff_dim, H = cfg.num_heads, dh = cfg.head_size, HD = H * dh, M = B * U, Mk = B * Tp;
  // workspace: x, xn, att, g [M, D]; hid [M, max(F, 2D)]; q [M, HD]; kv [Mk, 2 HD]
  const size_t wide = (size_t)std::max(F, 2 * D);
  const size_t need = (size_t)M * (4 * D + wide + HD) + (size_t)Mk * 2 * HD + 1024;
  if (need > h->tr_ws_floats) {
    ENG_CUDA(h, cudaDeviceSynchronize());
    if (h->tr_ws) ENG_CUDA(h, cudaFree(h->tr_ws));
    h->tr_ws = nullptr; h->tr_ws_floats = 0;
    ENG_CUDA(h, cudaMalloc(&h->tr_ws, need * sizeof(float)));
    h->tr_ws_floats = need;
  }
  auto al = [](size_t v) { return (v + 63) / 64 * 64; };
  float* x = h->tr_ws;
  float* xn = x + al((size_t)M * D);
  float* att = xn + al((size_t)M * D);
  float* g = att + al((size_t)M * D);
  float* hid = g + al((size_t)M * D);
  float* q = hid + al((size_t)M * wide);
  float* kv = q + al((size_t)M * HD);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Ctx c{h, st};
  const bool tc = cfg.precision == B200ASR_PRECISION_TF32;
  const float eps = cfg.ln_eps;
  h->launches++;
  ENG_TRY(h, launch_embed(ids_dev, h->tr_emb, x, M, D, cfg.tr_inp_classes, st));
  auto ln = [&](const LNW& w, const float* pe) -> int {
    h->launches++;
    return launch_layernorm(x, w.g, w.b, xn, M, D, eps, st, pe, U, tc ? 1 : 0);
  };
  auto ffn = [&](const FFNW& f) -> int {
    if (ln(f.ln, nullptr)) return 1;
    if (gemm(c, xn, D, f.w1, f.b1, nullptr, 0.f, hid, F, M, F, D, EPI_BIAS_SWISH, true)) return 1;
    return gemm(c, hid, F, f.w2, f.b2, x, 0.5f, x, D, M, D, F, EPI_RESID);
  };
  for (const BlockW& w : h->tr_blocks) {
    ENG_TRY(h, ffn(w.ffn1));
    // RMHSAModule: queries from LN(x + positions); keys / values straight from the encoder states; residual on x
    ENG_TRY(h, ln(w.mhsa.ln, h->tr_pe));
    ENG_TRY(h, gemm(c, xn, D, w.mhsa.wqkv, nullptr, nullptr, 0.f, q, HD, M, HD, D, EPI_NONE));                          // rows [0, HD) of wqkv = Wq (pre-scaled)
    ENG_TRY(h, gemm(c, enc_dev, D, w.mhsa.wqkv + (size_t)HD * D, nullptr, nullptr, 0.f, kv, 2 * HD, Mk, 2 * HD, D, EPI_NONE));   // rows [HD, 3 HD) = Wk | Wv
    h->launches++;
    ENG_TRY(h, launch_cross_attention(q, kv, att, B, U, Tp, H, dh, tc ? 1 : 0, st));
    ENG_TRY(h, gemm(c, att, HD, w.mhsa.wo, w.mhsa.bo, x, 1.0f, x, D, M, D, HD, EPI_RESID));
    // ConvModule ('same' padding)
    ENG_TRY(h, ln(w.conv.ln, nullptr));
    ENG_TRY(h, gemm(c, xn, D, w.conv.pw1w, w.conv.pw1b, nullptr, 0.f, g, D, M, 2 * D, D, EPI_GLU));
    DwConvParams dp{};
    dp.x = g; dp.w = w.conv.dww; dp.y = att; dp.B = B; dp.T = U; dp.D = D; dp.K = w.kernel_size; dp.pad_left = same_pad(U, w.kernel_size, 1).before;
    dp.round_tf32 = tc ? 1 : 0;
    h->launches++;
    ENG_TRY(h, launch_dwconv(dp, st));
    ENG_TRY(h, gemm(c, att, D, w.conv.pww, w.conv.pwb, nullptr, 0.f, hid, 2 * D, M, 2 * D, D, EPI_BIAS_SWISH, true));
    ENG_TRY(h, gemm(c, hid, 2 * D, w.conv.pw2w, w.conv.pw2b, x, 1.0f, x, D, M, D, 2 * D, EPI_RESID));
    ENG_TRY(h, ffn(w.ffn2));
    h->launches++;
    ENG_TRY(h, launch_layernorm(x, w.ln.g, w.ln.b, x, M, D, eps, st));
  }
  ENG_TRY(h, gemm(c, x, D, h->tr_fcw, h->tr_fcb, nullptr, 0.f, logits_dev, cfg.tr_vocab, M, cfg.tr_vocab, D, EPI_BIAS));
  return 0;
}

B200ASR_API int b200asr_recognize(b200asr_handle h, const float* wav_dev, int B, int L, int32_t* ids_dev, int32_t* out_len_dev,
                      void* stream) {
  return b200asr_recognize_lengths(h, wav_dev, nullptr, B, L, ids_dev, out_len_dev, stream);
}

B200ASR_API int b200asr_recognize_lengths(b200asr_handle h, const float* wav_dev, const int32_t* frame_lengths_dev, int B, int L,
                                          int32_t* ids_dev, int32_t* out_len_dev, void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (h->cfg.vocab <= 0) return fail(h, "b200asr_recognize: engine was created without a CTC decoder");
  if (B < 0 || L <= 0 || !wav_dev || !ids_dev || !out_len_dev) return fail(h, "b200asr_recognize: bad arguments");
  if (B == 0) return 0;
  const int B0 = B;
  if (effective_batch(h, &B, &L)) return 1;
  Shapes s = shapes_for(h, B, L);
  Buffers b;
  if (ensure_workspace(h, s, &b)) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Tp = s.M / B0;  // frames per original utterance (chunks concatenated in time)
  GraphKey key{};
  key.kind = 3; key.B = B; key.L = L; key.p0 = wav_dev; key.p1 = ids_dev; key.p2 = out_len_dev; key.p3 = frame_lengths_dev;
  return with_graph(h, st, key, [&](cudaStream_t st) -> int {
    Ctx c{h, st};
    ENG_TRY(h, run_encoder(c, wav_dev, s, b));
    // the CTC decoder sees whole utterances: [B0, Tp, D]; its input is the encoder output held in b.x
    // tensor-core path: CTC head fused with the per-frame argmax (no 42.6 MB logits round trip at the benchmark shape)
    GemmParams probe{};
    probe.A = b.x; probe.W = h->ctc_fcw; probe.bias = h->ctc_fcb; probe.C = reinterpret_cast<float*>(b.amp); probe.M = s.M;
    probe.N = h->cfg.vocab; probe.K = h->cfg.dmodel; probe.lda = h->cfg.dmodel; probe.ldc = h->cfg.vocab;
    const bool fused_argmax = h->cfg.precision == B200ASR_PRECISION_TF32 && fused_ln_ok(h) && !h->ctc_blocks.empty() &&
                              tc_gemm_supported(probe, EPI_BIAS_ARGMAX) && h->cfg.vocab / 64 + 2 >= tc_argmax_tiles(h->cfg.vocab);
    h->launches += 2;
    if (fused_argmax) {
      ENG_TRY(h, run_ctc(c, b.x, B0, Tp, b, nullptr));
      ENG_TRY(h, launch_ctc_greedy_partials(b.amp, tc_argmax_tiles(h->cfg.vocab), frame_lengths_dev, B0, Tp, h->cfg.vocab - 1, b.am, ids_dev,
                                            out_len_dev, st));
    } else {
      ENG_TRY(h, run_ctc(c, b.x, B0, Tp, b, b.logits));
      ENG_TRY(h, launch_ctc_greedy(b.logits, frame_lengths_dev, B0, Tp, h->cfg.vocab, h->cfg.vocab - 1, b.am, ids_dev, out_len_dev, st));
    }
    return 0;
  });
}

B200ASR_API int b200asr_recognize_host(b200asr_handle h, const float* wav_host, int B, int L, int32_t* ids_host, int32_t* out_len_host,
                           void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (B < 0 || L <= 0 || !wav_host || !ids_host || !out_len_host) return fail(h, "b200asr_recognize_host: bad arguments");
  if (B == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // staging buffers live at the end of a workspace sized for this shape
  int Be = B, Le = L;
  if (effective_batch(h, &Be, &Le)) return 1;
  Shapes s = shapes_for(h, Be, Le);
  Buffers b;
  if (ensure_workspace(h, s, &b)) return 1;
  const int Tp = s.M / B;
  // the waveform is staged in a buffer of its own that grows on demand (outside the workspace: any shape works)
  const size_t need = (size_t)B * L;
  if (need > h->stage_wav_floats) {
    ENG_CUDA(h, cudaStreamSynchronize(st));
    if (h->stage_wav) ENG_CUDA(h, cudaFree(h->stage_wav));
    h->stage_wav = nullptr; h->stage_wav_floats = 0;
    // cached graphs keyed on the old staging address can never be hit again; they age out of the LRU cache
    ENG_CUDA(h, cudaMalloc(&h->stage_wav, need * sizeof(float)));
    h->stage_wav_floats = need;
  }
  float* wav_dev = h->stage_wav;
  ENG_CUDA(h, cudaMemcpyAsync(wav_dev, wav_host, sizeof(float) * need, cudaMemcpyHostToDevice, st));
  if (b200asr_recognize(h, wav_dev, B, L, b.ids, b.lens, st)) return 1;
  ENG_CUDA(h, cudaMemcpyAsync(ids_host, b.ids, sizeof(int32_t) * (size_t)B * Tp, cudaMemcpyDeviceToHost, st));
  ENG_CUDA(h, cudaMemcpyAsync(out_len_host, b.lens, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, st));
  ENG_CUDA(h, cudaStreamSynchronize(st));
  return 0;
}

B200ASR_API int b200asr_recognize_host_submit(b200asr_handle h, int slot, const float* wav_host, int B, int L, int32_t* ids_host,
                                  int32_t* out_len_host) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (slot < 0 || slot > 1 || B <= 0 || L <= 0 || !wav_host || !ids_host || !out_len_host)
    return fail(h, "b200asr_recognize_host_submit: bad arguments");
  if (h->cfg.vocab <= 0) return fail(h, "b200asr_recognize_host_submit: engine was created without a CTC decoder");
  auto& ps = h->pipe[slot];
  if (ps.busy) return fail(h, "b200asr_recognize_host_submit: slot still in flight (collect it first)");
  if (!h->pipe_copy) {
    ENG_CUDA(h, cudaStreamCreateWithFlags(&h->pipe_copy, cudaStreamNonBlocking));
    ENG_CUDA(h, cudaStreamCreateWithFlags(&h->pipe_compute, cudaStreamNonBlocking));
  }
  if (!ps.h2d) {
    ENG_CUDA(h, cudaEventCreateWithFlags(&ps.h2d, cudaEventDisableTiming));
    ENG_CUDA(h, cudaEventCreateWithFlags(&ps.done, cudaEventDisableTiming));
  }
  int Be = B, Le = L;
  if (effective_batch(h, &Be, &Le)) return 1;
  const Shapes s = shapes_for(h, Be, Le);
  const int Tp = s.M / B;
  const size_t nw = (size_t)B * L, ni = (size_t)B * Tp;
  // staging buffers live outside the workspace: the other slot's graph may be running on it while this slot's waveform lands
  if (nw > ps.wav_floats) {
    if (ps.wav) ENG_CUDA(h, cudaFree(ps.wav));
    ps.wav = nullptr; ps.wav_floats = 0;
    ENG_CUDA(h, cudaMalloc(&ps.wav, nw * sizeof(float)));
    ps.wav_floats = nw;
  }
  if (ni > ps.id_ints) {
    if (ps.ids) ENG_CUDA(h, cudaFree(ps.ids));
    ps.ids = nullptr; ps.id_ints = 0;
    ENG_CUDA(h, cudaMalloc(&ps.ids, ni * sizeof(int32_t)));
    ps.id_ints = ni;
  }
  if ((size_t)B > ps.len_ints) {
    if (ps.lens) ENG_CUDA(h, cudaFree(ps.lens));
    ps.lens = nullptr; ps.len_ints = 0;
    ENG_CUDA(h, cudaMalloc(&ps.lens, (size_t)B * sizeof(int32_t)));
    ps.len_ints = (size_t)B;
  }
  Buffers b;
  if (ensure_workspace(h, s, &b)) return 1;     // (may synchronise the device when it has to grow: only before the first submit)
  ENG_CUDA(h, cudaMemcpyAsync(ps.wav, wav_host, nw * sizeof(float), cudaMemcpyHostToDevice, h->pipe_copy));
  ENG_CUDA(h, cudaEventRecord(ps.h2d, h->pipe_copy));
  ENG_CUDA(h, cudaStreamWaitEvent(h->pipe_compute, ps.h2d, 0));
  if (b200asr_recognize(h, ps.wav, B, L, ps.ids, ps.lens, h->pipe_compute)) return 1;
  ENG_CUDA(h, cudaMemcpyAsync(ids_host, ps.ids, ni * sizeof(int32_t), cudaMemcpyDeviceToHost, h->pipe_compute));
  ENG_CUDA(h, cudaMemcpyAsync(out_len_host, ps.lens, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, h->pipe_compute));
  ENG_CUDA(h, cudaEventRecord(ps.done, h->pipe_compute));
  ps.busy = true;
  return 0;
}

B200ASR_API int b200asr_recognize_host_collect(b200asr_handle h, int slot) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (slot < 0 || slot > 1) return fail(h, "b200asr_recognize_host_collect: bad slot");
  auto& ps = h->pipe[slot];
  if (!ps.busy) return fail(h, "b200asr_recognize_host_collect: nothing was submitted on this slot");
  ENG_CUDA(h, cudaEventSynchronize(ps.done));
  ps.busy = false;
  return 0;
}

// Time ONE stage of the schedule in isolation (bench.py roofline): `iters` back-to-back launches of the named kernel on
// `stream`, bracketed by CUDA events on that stream; operands are whatever the last forward pass left in the workspace
// (run b200asr_recognize with the same (B, L) first).  Also reports the stage's algorithmic FLOPs and HBM bytes per launch.
B200ASR_API int b200asr_time_stage(b200asr_handle h, int stage, int B, int L, int iters, void* stream, float* ms_per_launch,
                       double* flops, double* bytes) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (B <= 0 || L <= 0 || iters <= 0 || !ms_per_launch || !flops || !bytes) return fail(h, "b200asr_time_stage: bad arguments");
  if (effective_batch(h, &B, &L)) return 1;
  Shapes s = shapes_for(h, B, L);
  Buffers b;
  if (ensure_workspace(h, s, &b)) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const b200asr_config& cfg = h->cfg;
  const double D = cfg.dmodel, M = s.M;
  Ctx c{h, st};
  cudaEvent_t e0, e1;
  ENG_CUDA(h, cudaEventCreate(&e0));
  ENG_CUDA(h, cudaEventCreate(&e1));
  int rc = 0;
  for (int it = -1; it < iters && rc == 0; ++it) {   // it == -1: one untimed warm-up launch
    if (it == 0) cudaEventRecord(e0, st);
    switch (stage) {
      case B200ASR_STAGE_CONV2: {
        if (h->use_fused_sub) {   // the fused conv1 + conv2 kernel (mel in, conv2 map out)
          rc = run_subsample_convs(c, b.mel, s, b);
          *flops = 2.0 * (double)s.B * s.T2 * h->F2 * D * 9.0 * D + 2.0 * 9.0 * (double)s.B * s.T1 * h->F1 * D;
          *bytes = 4.0 * ((double)s.B * s.T * cfg.n_mels + 9.0 * D * D + (double)s.B * s.T2 * h->F2 * D);
          break;
        }
        GemmParams g = conv2_params(h, s, b);
        const int mode = sub_f16_mode(h, s, b);     // (b.c1 holds the map the last recognize call left there, in this very format)
        if (mode >= 1) { g.f16 = 1; g.W = h->c2w16; }
        if (mode == 2) g.out_f16 = 1;
        const double in_bytes = mode >= 1 ? 2.0 : 4.0, out_bytes = mode == 2 ? 2.0 : 4.0;
        if (cfg.precision == B200ASR_PRECISION_TF32 && tc_gemm_supported(g, EPI_BIAS_RELU)) rc = launch_gemm_tc(h->tc, g, EPI_BIAS_RELU, st);
        else rc = launch_gemm_simt(g, EPI_BIAS_RELU, st);
        *flops = 2.0 * g.M * D * 9.0 * D;
        *bytes = in_bytes * ((double)s.B * s.T1 * h->F1 * D + 9.0 * D * D) + out_bytes * (double)g.M * D;
        break;
      }
      case B200ASR_STAGE_FFN_W1: {
        const FFNW& f = h->enc_blocks[0].ffn1;
        rc = gemm(c, b.xn, cfg.dmodel, f.w1, f.b1, nullptr, 0.f, b.h, cfg.ff_dim, s.M, cfg.ff_dim, cfg.dmodel, EPI_BIAS_SWISH);
        *flops = 2.0 * M * D * cfg.ff_dim;
        *bytes = 4.0 * (M * D + D * cfg.ff_dim + M * cfg.ff_dim);
        break;
      }
      case B200ASR_STAGE_FFN_W2: {
        const FFNW& f = h->enc_blocks[0].ffn1;
        rc = gemm(c, b.h, cfg.ff_dim, f.w2, f.b2, b.x, 0.5f, b.x, cfg.dmodel, s.M, cfg.dmodel, cfg.ff_dim, EPI_RESID);
        *flops = 2.0 * M * D * cfg.ff_dim;
        *bytes = 4.0 * (M * cfg.ff_dim + D * cfg.ff_dim + 2.0 * M * D);
        break;
      }
      case B200ASR_STAGE_STFT: {
        rc = run_frontend(c, b.c2 /* any readable [B,L] floats */, s, b, b.mel);
        *flops = 0.0;
        *bytes = 4.0 * ((double)s.B * s.L + 2.0 * s.B * s.T * 513.0 + (double)s.B * s.T * cfg.n_mels);
        break;
      }
      case B200ASR_STAGE_SUBLIN: {
        const bool lin16 = sub_f16_mode(h, s, b) == 2;            // (b.c2 holds what conv2 wrote last: fp16 then)
        if (fused_ln_ok(h) && !h->enc_blocks.empty()) {          // the schedule's own launch: bias + LayerNorm fused, writes b.x and b.xn
          GemmParams lp = sublin_params(h, s, b, lin16);
          rc = gemm_p(c, lp, EPI_BIAS_LN);
        } else {
          rc = gemm(c, b.c2, h->F2 * cfg.dmodel, h->linw, h->linb, nullptr, 0.f, b.xn, cfg.dmodel, s.M, cfg.dmodel, h->F2 * cfg.dmodel, EPI_BIAS);
        }
        *flops = 2.0 * M * D * h->F2 * D;
        *bytes = (lin16 ? 2.0 : 4.0) * (M * h->F2 * D + D * h->F2 * D) + 4.0 * 2.0 * M * D;
        break;
      }
      case B200ASR_STAGE_ATTENTION: {
        AttnParams ap{};
        ap.qkv = b.h; ap.out = b.att; ap.B = s.B; ap.T = s.T2; ap.H = cfg.num_heads; ap.dh = cfg.head_size; ap.win_front = -1;
        ap.round_tf32 = 1; ap.async_stage = h->attn_async ? 1 : 0;   // as in run_block_fused (b.h holds the last block's pre-rounded QKV)
        rc = attention(c, ap);
        *flops = 4.0 * s.B * cfg.num_heads * (double)s.T2 * s.T2 * cfg.head_size;
        *bytes = 4.0 * (3.0 * M * cfg.num_heads * cfg.head_size + M * cfg.num_heads * cfg.head_size);
        break;
      }
      case B200ASR_STAGE_CTC_FC: {
        rc = gemm(c, b.x, cfg.dmodel, h->ctc_fcw, h->ctc_fcb, nullptr, 0.f, b.logits, cfg.vocab, s.M, cfg.vocab, cfg.dmodel, EPI_BIAS);
        *flops = 2.0 * M * D * cfg.vocab;
        *bytes = 4.0 * (M * D + D * cfg.vocab + M * cfg.vocab);
        break;
      }
      case B200ASR_STAGE_FFN_CHAIN: {
        const BlockW& w = h->enc_blocks[0];
        rc = chain_resid_ln(c, b.xn, cfg.dmodel, w.ffn1.w1, w.ffn1.b1, cfg.ff_dim, w.ffn1.w2, w.ffn1.b2, 0.5f, b, s.M, cfg.dmodel, w.mhsa.ln,
                            nullptr, cfg.ln_eps);
        *flops = 4.0 * M * D * cfg.ff_dim;
        *bytes = 4.0 * (4.0 * M * D + 2.0 * D * cfg.ff_dim);   // xn + x in, x + xn out, both weight matrices once
        break;
      }
      case B200ASR_STAGE_CONV1: {
        if (h->use_fused_sub) {
          rc = 1;
          snprintf(g_errbuf, sizeof(g_errbuf), "b200asr_time_stage: conv1 is fused into the conv2 kernel on this engine (stage conv2 times both)");
          break;
        }
        Conv1Params c1{};
        c1.mel = b.mel; c1.w = h->c1w; c1.bias = h->c1b; c1.out = b.c1; c1.B = s.B; c1.T = s.T; c1.F = cfg.n_mels; c1.T1 = s.T1;
        c1.F1 = h->F1; c1.D = cfg.dmodel; c1.pad_t = s.pt1; c1.pad_f = s.pf1;
        c1.round_tf32 = (cfg.precision == B200ASR_PRECISION_TF32) ? 1 : 0;
        c1.out_f16 = sub_f16_mode(h, s, b) >= 1 ? 1 : 0;
        rc = launch_conv1(c1, st);
        *flops = 2.0 * 9.0 * s.B * s.T1 * h->F1 * D;
        *bytes = 4.0 * (double)s.B * s.T * cfg.n_mels + (c1.out_f16 ? 2.0 : 4.0) * (double)s.B * s.T1 * h->F1 * D;
        break;
      }
      case B200ASR_STAGE_DWCONV: {
        DwConvParams dp{};
        dp.x = b.g; dp.w = h->enc_blocks[0].conv.dww; dp.y = b.att; dp.B = s.B; dp.T = s.T2; dp.D = cfg.dmodel; dp.K = cfg.kernel_size;
        dp.pad_left = same_pad(s.T2, cfg.kernel_size, 1).before;
        rc = launch_dwconv(dp, st);
        *flops = 2.0 * M * D * cfg.kernel_size;
        *bytes = 4.0 * 2.0 * M * D;
        break;
      }
      case B200ASR_STAGE_QKV: {
        const int HD = cfg.num_heads * cfg.head_size;
        rc = gemm(c, b.xn, cfg.dmodel, h->enc_blocks[0].mhsa.wqkv, nullptr, nullptr, 0.f, b.h, 3 * HD, s.M, 3 * HD, cfg.dmodel, EPI_NONE);
        *flops = 2.0 * M * D * 3.0 * HD;
        *bytes = 4.0 * (M * D + 3.0 * HD * D + M * 3.0 * HD);
        break;
      }
      default:
        rc = 1;
        snprintf(g_errbuf, sizeof(g_errbuf), "b200asr_time_stage: unknown stage %d", stage);
    }
  }
  cudaEventRecord(e1, st);
  cudaError_t e = cudaEventSynchronize(e1);
  float ms = 0.f;
  if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc != 0) return fail_cuda(h);
  if (e != cudaSuccess) {
    snprintf(g_errbuf, sizeof(g_errbuf), "b200asr_time_stage: %s", cudaGetErrorString(e));
    return fail_cuda(h);
  }
  *ms_per_launch = ms / iters;
  return 0;
}

// Test hook: one GEMM through either arithmetic path (tests/test_gpu_parity.py validates tcgen05 against fp32 CUDA cores).
B200ASR_API int b200asr_debug_gemm(b200asr_handle h, const float* A, const float* W, const float* bias, const float* resid, float* C,
                       int M, int N, int K, int lda, int ldc, float alpha, int epilogue, int use_tensor_cores, void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  GemmParams p{};
  p.A = A; p.W = W; p.bias = bias; p.resid = resid; p.C = C; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldc = ldc; p.alpha = alpha;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  h->launches++;
  if (use_tensor_cores) {
    if (!tc_gemm_supported(p, epilogue)) return fail(h, "b200asr_debug_gemm: shape not supported by the tcgen05 path");
    ENG_TRY(h, launch_gemm_tc(h->tc, p, epilogue, st));
  } else {
    ENG_TRY(h, launch_gemm_simt(p, epilogue, st));
  }
  return 0;
}

// Test hook for the fused-LayerNorm epilogues of the tcgen05 kernel (epilogue 6, 7 or 8; C may alias resid).
B200ASR_API int b200asr_debug_gemm_ln(b200asr_handle h, const float* A, const float* W, const float* bias, const float* resid, float* C,
                          float* C2, int M, int N, int K, float alpha, int epilogue, const float* ln1_g, const float* ln1_b,
                          const float* ln2_g, const float* ln2_b, float eps, void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  GemmParams p{};
  p.A = A; p.W = W; p.bias = bias; p.resid = resid; p.C = C; p.C2 = C2; p.M = M; p.N = N; p.K = K; p.lda = K; p.ldc = N;
  p.alpha = alpha; p.ln1_g = ln1_g; p.ln1_b = ln1_b; p.ln2_g = ln2_g; p.ln2_b = ln2_b; p.ln_eps = eps;
  if (!tc_gemm_supported(p, epilogue)) return fail(h, "b200asr_debug_gemm_ln: shape not supported by the tcgen05 path");
  h->launches++;
  ENG_TRY(h, launch_gemm_tc(h->tc, p, epilogue, static_cast<cudaStream_t>(stream)));
  return 0;
}

// Test hook: the depthwise convolution of the conv module alone (y[b,t,c] = sum_j x[b, t + j - pad_left, c] * w[j, c]).
B200ASR_API int b200asr_debug_dwconv(b200asr_handle h, const float* x, const float* w, float* y, int B, int T, int D, int K, int pad_left,
                                     int round_tf32, void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (!x || !w || !y || B <= 0 || T <= 0 || D <= 0 || K <= 0) return fail(h, "b200asr_debug_dwconv: bad arguments");
  DwConvParams dp{};
  dp.x = x; dp.w = w; dp.y = y; dp.B = B; dp.T = T; dp.D = D; dp.K = K; dp.pad_left = pad_left; dp.round_tf32 = round_tf32;
  h->launches++;
  ENG_TRY(h, launch_dwconv(dp, static_cast<cudaStream_t>(stream)));
  return 0;
}

// Test hook: the two subsampling convolutions alone: mel [B, T, n_mels] -> relu(conv2(relu(conv1(mel)))) [B, T2, F2, D]
// (NHWC, the tensor the subsampling linear layer reads), through the engine's own precision path.
B200ASR_API int b200asr_debug_subsample_convs(b200asr_handle h, const float* mel_dev, int B, int T, float* out_dev, void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (!mel_dev || !out_dev || B <= 0 || T <= 0) return fail(h, "b200asr_debug_subsample_convs: bad arguments");
  // workspace sized through an equivalent (B, L) with ceil(L / hop) == T
  Shapes s = shapes_for(h, B, T * h->cfg.hop);
  if (s.T != T) return fail(h, "b200asr_debug_subsample_convs: internal shape inference mismatch");
  Buffers b;
  if (ensure_workspace(h, s, &b)) return 1;
  Ctx c{h, static_cast<cudaStream_t>(stream)};
  const bool keep = h->sub_out_f16;
  h->sub_out_f16 = false;                   // the hook returns conv2's map as fp32 (the conv1 map may still be fp16)
  const int rc_convs = run_subsample_convs(c, mel_dev, s, b);
  h->sub_out_f16 = keep;
  ENG_TRY(h, rc_convs);
  ENG_CUDA(h, cudaMemcpyAsync(out_dev, b.c2, sizeof(float) * (size_t)s.B * s.T2 * h->F2 * h->cfg.dmodel, cudaMemcpyDeviceToDevice, c.s));
  return 0;
}

// Test hook: one multi-head attention call through the tcgen05 kernel (use_tensor_cores = 1) or the fp32 CUDA-core kernel.
B200ASR_API int b200asr_debug_attention(b200asr_handle h, const float* qkv, float* out, int B, int T, int H, int dh, int win_front,
                            int win_back, int use_tensor_cores, void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  AttnParams ap{};
  ap.qkv = qkv; ap.out = out; ap.B = B; ap.T = T; ap.H = H; ap.dh = dh; ap.win_front = win_front; ap.win_back = win_back;
  ap.async_stage = (use_tensor_cores == 2) ? 1 : 0;   // 2: the caller's qkv holds tf32 numbers already (cp.async staging, no rounding)
  h->launches++;
  if (use_tensor_cores) {
    if (!attention_tc_supported(ap)) return fail(h, "b200asr_debug_attention: shape not supported by the tcgen05 kernel");
    ENG_TRY(h, launch_attention_tc(ap, static_cast<cudaStream_t>(stream)));
  } else {
    ENG_TRY(h, launch_attention(ap, static_cast<cudaStream_t>(stream)));
  }
  return 0;
}

// Test hook: the chained two-GEMM kernel.  epilogue 6 or 7 as in b200asr_debug_gemm_ln; C may alias resid.
B200ASR_API int b200asr_debug_chain(b200asr_handle h, const float* X, const float* W1, const float* b1, const float* W2, const float* b2,
                        const float* resid, float* C, float* C2, int M, int K1, int N1, int N2, float alpha, int epilogue,
                        const float* ln1_g, const float* ln1_b, const float* ln2_g, const float* ln2_b, float eps, void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  ChainGemmParams cp{};
  cp.X = X; cp.W1 = W1; cp.bias1 = b1; cp.W2 = W2; cp.bias2 = b2; cp.resid = resid; cp.C = C; cp.C2 = C2; cp.M = M; cp.K1 = K1; cp.N1 = N1;
  cp.N2 = N2; cp.ldx = K1; cp.alpha = alpha; cp.ln1_g = ln1_g; cp.ln1_b = ln1_b; cp.ln2_g = ln2_g; cp.ln2_b = ln2_b; cp.ln_eps = eps;
  if (!tc_chain_supported(cp, epilogue)) return fail(h, "b200asr_debug_chain: shape not supported by the chained kernel");
  h->launches++;
  ENG_TRY(h, launch_gemm_chain(h->tc, cp, epilogue, static_cast<cudaStream_t>(stream)));
  return 0;
}

// Test hook: the cluster-pair variant of the chained kernel (same arguments).
B200ASR_API int b200asr_debug_chain_pair(b200asr_handle h, const float* X, const float* W1, const float* b1, const float* W2, const float* b2,
                             const float* resid, float* C, float* C2, int M, int K1, int N1, int N2, float alpha, int epilogue,
                             const float* ln1_g, const float* ln1_b, const float* ln2_g, const float* ln2_b, float eps, void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  ChainGemmParams cp{};
  cp.X = X; cp.W1 = W1; cp.bias1 = b1; cp.W2 = W2; cp.bias2 = b2; cp.resid = resid; cp.C = C; cp.C2 = C2; cp.M = M; cp.K1 = K1; cp.N1 = N1;
  cp.N2 = N2; cp.ldx = K1; cp.alpha = alpha; cp.ln1_g = ln1_g; cp.ln1_b = ln1_b; cp.ln2_g = ln2_g; cp.ln2_b = ln2_b; cp.ln_eps = eps;
  if (N1 == 0 ? !tc_pair_direct_supported(cp, epilogue) : !tc_chain_pair_supported(cp, epilogue))
    return fail(h, "b200asr_debug_chain_pair: shape not supported by the cluster-pair kernel");
  h->launches++;
  ENG_TRY(h, launch_gemm_chain_pair(h->tc, cp, epilogue, static_cast<cudaStream_t>(stream)));
  return 0;
}

}  // extern "C"
