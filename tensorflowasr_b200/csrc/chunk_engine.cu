// ChunkConformer (asr/models/chunk_conformer_blocks.py): causal chunk streaming with state caches behind the stream-state C ABI of
// include/b200asr.h -- picker step (front end + encoder + picker, :807-824), feature_pick (:913-999), decoder step (helper +
// decoder with look-ahead, :826-837).  The GEMM-shaped work runs through the same tcgen05 kernels as the offline path (chained
// FFModule / conv-tail cluster-pair kernels, fused LayerNorm epilogues, the fused conv1->conv2 subsampler with 'valid' geometry);
// the cache handling lives in chunk_ops.cu.  One streaming block = 8 GEMM-class launches + attention + depthwise conv + two cache rolls.
#include "engine_internal.cuh"

namespace b200asr {

struct ChunkModel {
  b200asr_chunk_config cfg;
  std::vector<BlockW> enc, picker, helper, dec;
  const float *c1w, *c1b, *c2w, *c2b, *linw, *linb;
  const float *pick_projw, *pick_projb, *pick_fcw, *pick_fcb;   // fc rows padded to Vp_pad (zero weights, -1e30 bias)
  const float *dec_projw, *dec_projb, *dec_fcw, *dec_fcb;
  int Vp_pad, Vt_pad;
  int T;          // encoder frames per step = chunk_num / reduction
  int sub;        // mel frames kept in front of a chunk = chunk_num / reduction
  int S;          // samples per step = chunk_num * hop
  int Tcat, T1, F1, T2, F2;   // 'valid' subsampler geometry of one step
};

void chunk_model_free(ChunkModel* m) { delete m; }

}  // namespace b200asr

struct b200asr_stream_state {
  b200asr_handle owner = nullptr;
  int B = 0, Tmax = 0;
  char* base = nullptr;
  size_t bytes = 0;
  float *wavbuf, *sub_cache, *power, *mel_new, *melcat, *c2, *c2s, *x, *xn, *hwide, *att, *g, *qkv, *logits_p, *logits_t, *carry, *cat;
  unsigned int* pmax;
  int* nmax_dev;
  std::vector<float*> kv[2], glu[2];   // [parity][block]: enc 0..E-1, picker, helper, dec
  int par_a = 0, par_b = 0;            // current parity of the (encoder + picker) / (helper + decoder) caches
  int c_enc = 0, c_pick = 0, c_help = 0, c_dec = 0, n_carry = 0;
};

namespace {

using b200asr::ChunkModel;

size_t up256(size_t v) { return (v + 255) / 256 * 256; }

struct BlockGeom { int D, F, H, dh, K, W; float eps; };

// One ChunkConformerBlock.stream_call (:382-389) on Tc new rows per stream.  Pre: b.xn = LN(b.x; w.ffn1.ln).  Post: b.x = block output,
// b.xn = LN(b.x; *next_ln) when given; kv_new / glu_new = the rolled caches.
int stream_block(Ctx& c, const BlockW& w, const Buffers& b, float* qkv, const BlockGeom& g, int B, int Tc, int cache_len, int win_back,
                 const float* kv_old, float* kv_new, const float* glu_old, float* glu_new, const LNW* next_ln) {
  const int M = B * Tc, HD = g.H * g.dh, D = g.D;
  if (chain_resid_ln(c, b.xn, D, w.ffn1.w1, w.ffn1.b1, g.F, w.ffn1.w2, w.ffn1.b2, 0.5f, b, M, D, w.mhsa.ln, nullptr, g.eps)) return 1;
  if (gemm(c, b.xn, D, w.mhsa.wqkv, w.mhsa.bqkv, nullptr, 0.f, qkv, 3 * HD, M, 3 * HD, D, w.mhsa.bqkv ? EPI_BIAS : EPI_NONE)) return 1;
  StreamAttnParams ap{};
  ap.qkv = qkv; ap.kv_cache = kv_old; ap.out = b.att; ap.B = B; ap.Tc = Tc; ap.H = g.H; ap.dh = g.dh; ap.W = g.W; ap.c = cache_len;
  ap.win_front = g.W; ap.win_back = win_back; ap.round_tf32 = 1;
  c.h->launches++;
  if (launch_stream_attention(ap, c.s)) return 1;
  CacheUpdateParams cu{};
  cu.old_cache = kv_old; cu.cur = qkv; cu.new_cache = kv_new; cu.B = B; cu.W = g.W; cu.C = 2 * HD; cu.Tc = Tc; cu.shift = Tc - win_back;
  cu.cur_ld = 3 * HD; cu.cur_col0 = HD;
  c.h->launches++;
  if (launch_stream_cache_update(cu, c.s)) return 1;
  if (gemm_resid_ln(c, b.att, HD, w.mhsa.wo, w.mhsa.bo, 1.0f, b, M, D, w.conv.ln, nullptr, g.eps)) return 1;
  if (gemm(c, b.xn, D, w.conv.pw1w, w.conv.pw1b, nullptr, 0.f, b.g, D, M, 2 * D, D, EPI_GLU)) return 1;
  StreamDwParams dp{};
  dp.cache = glu_old; dp.cur = b.g; dp.w = w.conv.dww; dp.y = b.att; dp.B = B; dp.Tc = Tc; dp.D = D; dp.K = g.K; dp.round_tf32 = 1;
  c.h->launches++;
  if (launch_stream_dwconv(dp, c.s)) return 1;
  CacheUpdateParams gu{};
  gu.old_cache = glu_old; gu.cur = b.g; gu.new_cache = glu_new; gu.B = B; gu.W = g.K - 1; gu.C = D; gu.Tc = Tc; gu.shift = Tc - win_back;
  gu.cur_ld = D; gu.cur_col0 = 0;
  c.h->launches++;
  if (launch_stream_cache_update(gu, c.s)) return 1;
  if (chain_resid_ln(c, b.att, D, w.conv.pww, w.conv.pwb, 2 * D, w.conv.pw2w, w.conv.pw2b, 1.0f, b, M, D, w.ffn2.ln, nullptr, g.eps)) return 1;
  LNW none{nullptr, nullptr};
  if (chain_resid_ln(c, b.xn, D, w.ffn2.w1, w.ffn2.b1, g.F, w.ffn2.w2, w.ffn2.b2, 0.5f, b, M, D, w.ln, next_ln ? next_ln : &none, g.eps)) return 1;
  return 0;
}

BlockGeom geom_of(const ChunkModel& m) {
  return BlockGeom{m.cfg.dmodel, m.cfg.ff_dim, m.cfg.num_heads, m.cfg.head_size, m.cfg.kernel_size, m.cfg.win_front, m.cfg.ln_eps};
}

Buffers buffers_of(b200asr_stream st) {
  Buffers b{};
  b.x = st->x; b.xn = st->xn; b.h = st->hwide; b.att = st->att; b.g = st->g; b.power = st->power; b.pmax = st->pmax;
  return b;
}

int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// strip the class padding: dst [rows, V] <- src [rows, Vpad][:, :V]
int copy_classes(b200asr_handle h, float* dst, const float* src, int rows, int V, int Vpad, cudaStream_t s) {
  ENG_CUDA(h, cudaMemcpy2DAsync(dst, sizeof(float) * V, src, sizeof(float) * Vpad, sizeof(float) * V, rows, cudaMemcpyDeviceToDevice, s));
  return 0;
}

}  // namespace

extern "C" {

B200ASR_API int b200asr_chunk_create(const void* weight_blob, size_t blob_bytes, const b200asr_chunk_config* cfg, int device, b200asr_handle* out) {
  if (!weight_blob || !cfg || !out) return fail(nullptr, "b200asr_chunk_create: null argument");
  if (cfg->abi_version != B200ASR_ABI_VERSION) return fail(nullptr, "b200asr_chunk_create: ABI version mismatch");
  *out = nullptr;
  if (cfg->n_dft != 1024) return fail(nullptr, "b200asr_chunk_create: n_dft must be 1024 (reference hard-codes it)");
  if (cfg->dmodel != 144 || cfg->ff_dim != 4 * cfg->dmodel || cfg->num_heads * cfg->head_size != cfg->dmodel)
    return fail(nullptr, "b200asr_chunk_create: the streaming schedule is built on the dmodel-144 fused kernels (chunk_conformerS.yml geometry)");
  if (cfg->picker_back != 0) return fail(nullptr, "b200asr_chunk_create: picker look-ahead (picker_back != 0) is not supported");
  if (cfg->chunk_num % cfg->reduction != 0 || cfg->reduction != 4 || cfg->chunk_num < 8)
    return fail(nullptr, "b200asr_chunk_create: chunk_num must be a multiple of reduction = 4, >= 8");
  if (cfg->win_front > 96 || cfg->dec_back < 0 || cfg->dec_back > 32) return fail(nullptr, "b200asr_chunk_create: unsupported attention band");
  b200asr_engine* h = nullptr;
  if (b200asr::engine_alloc(weight_blob, blob_bytes, device, "b200asr_chunk_create", &h)) return 1;
  auto bail = [&]() {
    std::string e = g_errbuf;
    b200asr_destroy(h);
    snprintf(g_errbuf, sizeof(g_errbuf), "%s", e.c_str());
    return 1;
  };
  // the shared helpers read the offline config struct: fill the fields they use
  b200asr_config& oc = h->cfg;
  memset(&oc, 0, sizeof(oc));
  oc.abi_version = B200ASR_ABI_VERSION;
  oc.dmodel = cfg->dmodel; oc.num_heads = cfg->num_heads; oc.head_size = cfg->head_size; oc.kernel_size = cfg->kernel_size; oc.ff_dim = cfg->ff_dim;
  oc.n_mels = cfg->n_mels; oc.n_dft = cfg->n_dft; oc.hop = cfg->hop; oc.ln_eps = cfg->ln_eps; oc.precision = B200ASR_PRECISION_TF32;
  oc.use_cuda_graph = cfg->use_cuda_graph;
  ChunkModel* m = new ChunkModel();
  h->chunk = m;
  m->cfg = *cfg;
  const int D = cfg->dmodel;
  m->T = cfg->chunk_num / cfg->reduction;
  m->sub = cfg->chunk_num / cfg->reduction;
  m->S = cfg->chunk_num * cfg->hop;
  m->Tcat = m->sub + cfg->chunk_num;
  m->T1 = (m->Tcat - 3) / 2 + 1;
  m->F1 = (cfg->n_mels + 4 - 3) / 2 + 1;
  m->T2 = (m->T1 - 3) / 2 + 1;
  m->F2 = (m->F1 - 3) / 2 + 1;
  if (m->T2 < m->T) { snprintf(g_errbuf, sizeof(g_errbuf), "b200asr_chunk_create: subsampler yields %d < %d frames per step", m->T2, m->T); return bail(); }
  m->Vp_pad = (cfg->phone_classes + 3) / 4 * 4;
  m->Vt_pad = (cfg->txt_classes + 3) / 4 * 4;
  bool ok = true;
  m->c1w = lookup(h, "sub.conv1.w", 9ull * D, &ok);
  m->c1b = ok ? lookup(h, "sub.conv1.b", D, &ok) : nullptr;
  m->c2w = ok ? lookup(h, "sub.conv2.w", 9ull * D * D, &ok) : nullptr;
  m->c2b = ok ? lookup(h, "sub.conv2.b", D, &ok) : nullptr;
  m->linw = ok ? lookup(h, "sub.lin.w", (uint64_t)m->F2 * D * D, &ok) : nullptr;
  m->linb = ok ? lookup(h, "sub.lin.b", D, &ok) : nullptr;
  m->pick_projw = ok ? lookup(h, "picker.proj.w", (uint64_t)D * D, &ok) : nullptr;
  m->pick_projb = ok ? lookup(h, "picker.proj.b", D, &ok) : nullptr;
  m->pick_fcw = ok ? lookup(h, "picker.fc.w", (uint64_t)m->Vp_pad * D, &ok) : nullptr;
  m->pick_fcb = ok ? lookup(h, "picker.fc.b", m->Vp_pad, &ok) : nullptr;
  m->dec_projw = ok ? lookup(h, "dec.proj.w", (uint64_t)D * D, &ok) : nullptr;
  m->dec_projb = ok ? lookup(h, "dec.proj.b", D, &ok) : nullptr;
  m->dec_fcw = ok ? lookup(h, "dec.fc.w", (uint64_t)m->Vt_pad * D, &ok) : nullptr;
  m->dec_fcb = ok ? lookup(h, "dec.fc.b", m->Vt_pad, &ok) : nullptr;
  if (!ok) return bail();
  struct { std::vector<BlockW>* v; int n; const char* prefix; } stacks[4] = {
      {&m->enc, cfg->enc_blocks, "enc."}, {&m->picker, cfg->picker_blocks, "picker."}, {&m->helper, cfg->helper_blocks, "helper."}, {&m->dec, cfg->dec_blocks, "dec."}};
  for (auto& sk : stacks) {
    sk.v->resize(sk.n);
    for (int i = 0; i < sk.n; ++i) {
      if (!load_block(h, std::string(sk.prefix) + std::to_string(i) + ".", D, cfg->ff_dim, cfg->num_heads, cfg->head_size, cfg->kernel_size, &(*sk.v)[i]))
        return bail();
      if ((*sk.v)[i].mhsa.bqkv == nullptr) {
        snprintf(g_errbuf, sizeof(g_errbuf), "b200asr_chunk_create: block %s%d has no q/k/v bias (mhsa.bqkv)", sk.prefix, i);
        return bail();
      }
    }
  }
  if (cfg->enc_blocks < 1 || cfg->picker_blocks < 1 || cfg->helper_blocks < 1 || cfg->dec_blocks < 1) {
    snprintf(g_errbuf, sizeof(g_errbuf), "b200asr_chunk_create: every stack needs at least one block");
    return bail();
  }
  if (b200asr::engine_init_frontend(h, weight_blob)) return bail();
  ConvSubParams probe{};
  probe.D = D; probe.F2 = m->F2; probe.B = 1; probe.T = 1; probe.w2 = m->c2w;
  if (!h->tc.ready || !conv_subsample_tc_supported(probe) || !fused_ln_ok(h)) {
    snprintf(g_errbuf, sizeof(g_errbuf), "b200asr_chunk_create: the tcgen05 path is unavailable for this geometry");
    return bail();
  }
  *out = h;
  return 0;
}

B200ASR_API int b200asr_stream_state_create(b200asr_handle h, int B, b200asr_stream* out) {
  if (!h || !out) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (!h->chunk) return fail(h, "b200asr_stream_state_create: not a chunk engine (use b200asr_chunk_create)");
  if (B <= 0) return fail(h, "b200asr_stream_state_create: B must be positive");
  const ChunkModel& m = *h->chunk;
  const b200asr_chunk_config& c = m.cfg;
  b200asr_stream st = new b200asr_stream_state();
  st->owner = h;
  st->B = B;
  st->Tmax = c.dec_back + 4 * m.T;
  const int D = c.dmodel, HD = c.num_heads * c.head_size, nblk = c.enc_blocks + c.picker_blocks + c.helper_blocks + c.dec_blocks;
  const size_t Mmax = (size_t)B * st->Tmax;
  size_t off = 0;
  auto take = [&](size_t nfloat) { const size_t o = off; off += up256(nfloat * sizeof(float)); return o; };
  const size_t o_wav = take((size_t)B * 2 * m.S), o_sub = take((size_t)B * m.sub * c.n_mels), o_pow = take((size_t)B * c.chunk_num * kPowerStride),
               o_mel = take((size_t)B * c.chunk_num * c.n_mels), o_cat = take((size_t)B * m.Tcat * c.n_mels),
               o_c2 = take((size_t)B * m.T2 * m.F2 * D), o_c2s = take((size_t)B * m.T * m.F2 * D), o_x = take(Mmax * D), o_xn = take(Mmax * D),
               o_h = take(Mmax * c.ff_dim), o_att = take(Mmax * D), o_g = take(Mmax * D), o_qkv = take(Mmax * 3 * HD),
               o_lp = take((size_t)B * m.T * m.Vp_pad), o_lt = take(Mmax * m.Vt_pad), o_carry = take((size_t)B * (c.dec_back + 1) * D),
               o_dcat = take(Mmax * D), o_pmax = take(B), o_nmax = take(64);
  std::vector<size_t> o_kv[2], o_glu[2];
  for (int p = 0; p < 2; ++p)
    for (int i = 0; i < nblk; ++i) {
      o_kv[p].push_back(take((size_t)B * c.win_front * 2 * HD));
      o_glu[p].push_back(take((size_t)B * (c.kernel_size - 1) * D));
    }
  st->bytes = off;
  if (cudaMalloc(&st->base, off) != cudaSuccess) {
    delete st;
    return fail(h, "b200asr_stream_state_create: cannot allocate the stream state");
  }
  auto F = [&](size_t o) { return reinterpret_cast<float*>(st->base + o); };
  st->wavbuf = F(o_wav); st->sub_cache = F(o_sub); st->power = F(o_pow); st->mel_new = F(o_mel); st->melcat = F(o_cat); st->c2 = F(o_c2);
  st->c2s = F(o_c2s); st->x = F(o_x); st->xn = F(o_xn); st->hwide = F(o_h); st->att = F(o_att); st->g = F(o_g); st->qkv = F(o_qkv);
  st->logits_p = F(o_lp); st->logits_t = F(o_lt); st->carry = F(o_carry); st->cat = F(o_dcat);
  st->pmax = reinterpret_cast<unsigned int*>(st->base + o_pmax);
  st->nmax_dev = reinterpret_cast<int*>(st->base + o_nmax);
  for (int p = 0; p < 2; ++p)
    for (int i = 0; i < nblk; ++i) {
      st->kv[p].push_back(F(o_kv[p][i]));
      st->glu[p].push_back(F(o_glu[p][i]));
    }
  *out = st;
  return b200asr_stream_state_reset(h, st);
}

B200ASR_API int b200asr_stream_state_reset(b200asr_handle h, b200asr_stream st) {
  if (!h || !st) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (st->owner != h) return fail(h, "b200asr_stream_state_reset: the state belongs to another handle");
  ENG_CUDA(h, cudaDeviceSynchronize());
  ENG_CUDA(h, cudaMemset(st->base, 0, st->bytes));
  st->par_a = st->par_b = 0;
  st->c_enc = st->c_pick = st->c_help = st->c_dec = st->n_carry = 0;
  return 0;
}

B200ASR_API int b200asr_stream_state_destroy(b200asr_handle h, b200asr_stream st) {
  if (!st) return 0;
  if (h) {
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    DeviceGuard dev_guard(h->device);
    cudaDeviceSynchronize();
    // graphs captured for this state hold its addresses
    for (auto it = h->graphs.begin(); it != h->graphs.end();) {
      if (it->first.p0 == st) { cudaGraphExecDestroy(it->second.exec); it = h->graphs.erase(it); }
      else ++it;
    }
    if (st->base) cudaFree(st->base);
  }
  delete st;
  return 0;
}

B200ASR_API int b200asr_stream_step(b200asr_handle h, b200asr_stream st, const float* wav_chunk_dev, float* phone_logits_dev, float* hidden_dev,
                        void* stream) {
  if (!h || !st) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (!h->chunk || st->owner != h) return fail(h, "b200asr_stream_step: not a chunk engine / foreign state");
  if (!wav_chunk_dev || !phone_logits_dev || !hidden_dev) return fail(h, "b200asr_stream_step: bad arguments");
  const ChunkModel& m = *h->chunk;
  const b200asr_chunk_config& cf = m.cfg;
  const int B = st->B, D = cf.dmodel, T = m.T, M = B * T;
  const BlockGeom geo = geom_of(m);
  const int par = st->par_a, c_enc = st->c_enc, c_pick = st->c_pick;
  GraphKey key{};
  key.kind = 20; key.B = B; key.L = (c_enc << 8) | (c_pick << 1) | par; key.p0 = st; key.p1 = wav_chunk_dev; key.p2 = phone_logits_dev; key.p3 = hidden_dev;
  const int rc = with_graph(h, static_cast<cudaStream_t>(stream), key, [&](cudaStream_t s) -> int {
    Ctx c{h, s};
    // ---- front end: wav cache roll, 'valid' mel of the newest chunk_num frames, mel cache, 'valid' subsampling (:455-466, :74-91)
    h->launches += 5;
    ENG_TRY(h, launch_stream_wav_shift(st->wavbuf, wav_chunk_dev, B, m.S, s));
    FrontendParams fp{};
    fp.wav = st->wavbuf; fp.window = h->window; fp.twiddle = h->twiddle; fp.melw = h->melw; fp.mel_lo = h->mel_lo; fp.mel_hi = h->mel_hi;
    fp.mel_off = h->mel_off; fp.mel_wc = h->mel_wc; fp.mel_nnz = h->mel_nnz; fp.power = st->power; fp.pmax = st->pmax; fp.mel = st->mel_new;
    fp.B = B; fp.L = 2 * m.S; fp.T = cf.chunk_num; fp.hop = cf.hop; fp.power_stride = kPowerStride; fp.n_mels = cf.n_mels; fp.mode = 1;
    fp.pad_left = (cf.n_dft - 1) - cf.chunk_num * cf.hop;   // frame t of the call = frame chunk_num + t of the 2S buffer: ends at sample 160 (chunk_num + t)
    ENG_TRY(h, launch_frontend(fp, s));
    ENG_TRY(h, launch_stream_mel_cat(st->melcat, st->sub_cache, st->mel_new, B, m.sub, cf.chunk_num, cf.n_mels, s));
    ConvSubParams cp{};
    cp.mel = st->melcat; cp.w1 = m.c1w; cp.b1 = m.c1b; cp.w2 = m.c2w; cp.b2 = m.c2b; cp.out = st->c2;
    cp.B = B; cp.T = m.Tcat; cp.F = cf.n_mels; cp.T1 = m.T1; cp.F1 = m.F1; cp.T2 = m.T2; cp.F2 = m.F2; cp.D = D;
    cp.pt1 = 0; cp.pf1 = 2; cp.pt2 = 0; cp.pf2 = 0; cp.round_out = 1;   // 'valid' convs; the [2, 2] frequency padding of the mel map is conv1's (:60)
    h->launches++;
    ENG_TRY(h, launch_conv_subsample_tc(h->tc, cp, s));
    const float* sub_in = st->c2;
    if (m.T2 != T) {   // keep the last T outputs (:89)
      h->launches++;
      ENG_TRY(h, launch_rows_slice(st->c2, m.T2, m.T2 - T, T, st->c2s, B, m.F2 * D, s));
      sub_in = st->c2s;
    }
    Buffers b = buffers_of(st);
    GemmParams lp{};
    lp.A = sub_in; lp.W = m.linw; lp.bias = m.linb; lp.C = b.x; lp.C2 = b.xn; lp.M = M; lp.N = D; lp.K = m.F2 * D; lp.lda = m.F2 * D; lp.ldc = D;
    lp.ln1_g = m.enc[0].ffn1.ln.g; lp.ln1_b = m.enc[0].ffn1.ln.b; lp.ln_eps = cf.ln_eps;
    ENG_TRY(h, gemm_p(c, lp, EPI_BIAS_LN));
    // ---- encoder (:532-563)
    int blk = 0;
    for (size_t i = 0; i < m.enc.size(); ++i, ++blk) {
      const LNW* next = (i + 1 < m.enc.size()) ? &m.enc[i + 1].ffn1.ln : nullptr;
      ENG_TRY(h, stream_block(c, m.enc[i], b, st->qkv, geo, B, T, c_enc, 0, st->kv[par][blk], st->kv[par ^ 1][blk], st->glu[par][blk],
                              st->glu[par ^ 1][blk], next));
    }
    // ---- picker (:626-658 with win_back = 0): project, block(s), phone head
    Buffers bb = b;
    std::swap(bb.x, bb.g);            // the projection must not alias its input (the encoder output in b.x)
    GemmParams pp{};
    pp.A = b.x; pp.W = m.pick_projw; pp.bias = m.pick_projb; pp.C = bb.x; pp.C2 = bb.xn; pp.M = M; pp.N = D; pp.K = D; pp.lda = D; pp.ldc = D;
    pp.ln1_g = m.picker[0].ffn1.ln.g; pp.ln1_b = m.picker[0].ffn1.ln.b; pp.ln_eps = cf.ln_eps;
    ENG_TRY(h, gemm_p(c, pp, EPI_BIAS_LN));
    for (size_t i = 0; i < m.picker.size(); ++i, ++blk) {
      const LNW* next = (i + 1 < m.picker.size()) ? &m.picker[i + 1].ffn1.ln : nullptr;
      ENG_TRY(h, stream_block(c, m.picker[i], bb, st->qkv, geo, B, T, c_pick, 0, st->kv[par][blk], st->kv[par ^ 1][blk], st->glu[par][blk],
                              st->glu[par ^ 1][blk], next));
    }
    ENG_TRY(h, gemm(c, bb.x, D, m.pick_fcw, m.pick_fcb, nullptr, 0.f, st->logits_p, m.Vp_pad, M, m.Vp_pad, D, EPI_BIAS));
    ENG_CUDA(h, cudaMemcpyAsync(hidden_dev, bb.x, sizeof(float) * (size_t)M * D, cudaMemcpyDeviceToDevice, s));
    ENG_TRY(h, copy_classes(h, phone_logits_dev, st->logits_p, M, cf.phone_classes, m.Vp_pad, s));
    return 0;
  });
  if (rc) return rc;
  st->par_a ^= 1;
  st->c_enc = clampi(c_enc + T, 0, cf.win_front);
  st->c_pick = clampi(c_pick + T, 0, cf.win_front);
  return 0;
}

B200ASR_API int b200asr_stream_feature_pick(b200asr_handle h, const float* hidden_dev, const float* phone_logits_dev, int B, int T, int V, int blank,
                                float* feats_dev, float* picked_logits_dev, int32_t* counts_dev, int32_t* n_max_host, void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (!h->chunk) return fail(h, "b200asr_stream_feature_pick: not a chunk engine");
  if (!hidden_dev || !phone_logits_dev || !feats_dev || !counts_dev || B <= 0 || T <= 0 || V <= 0) return fail(h, "b200asr_stream_feature_pick: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int* nmax_dev = nullptr;
  ENG_CUDA(h, cudaMallocAsync(&nmax_dev, sizeof(int), s));
  PickParams p{};
  p.hidden = hidden_dev; p.logits = phone_logits_dev; p.feats = feats_dev; p.picked = picked_logits_dev; p.counts = counts_dev; p.n_max = nmax_dev;
  p.B = B; p.T = T; p.D = h->chunk->cfg.dmodel; p.V = V; p.ldv = V; p.blank = blank;
  h->launches++;
  int rc = launch_feature_pick(p, s);
  if (rc == 0 && n_max_host) {
    int v = 0;
    cudaError_t e = cudaMemcpyAsync(&v, nmax_dev, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { snprintf(g_errbuf, sizeof(g_errbuf), "b200asr_stream_feature_pick: %s", cudaGetErrorString(e)); rc = 1; }
    *n_max_host = v;
  }
  cudaFreeAsync(nmax_dev, s);
  return rc ? fail_cuda(h) : 0;
}

B200ASR_API int b200asr_stream_decoder_rows(b200asr_handle h, b200asr_stream st, int n) {
  if (!h || !st || n < 0) return -1;
  return st->n_carry + n;
}

B200ASR_API int b200asr_stream_decoder_step(b200asr_handle h, b200asr_stream st, const float* feats_dev, int n, float* txt_logits_dev,
                                int32_t* n_rows_host, int32_t* n_valid_host, void* stream) {
  if (!h || !st) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (!h->chunk || st->owner != h) return fail(h, "b200asr_stream_decoder_step: not a chunk engine / foreign state");
  const ChunkModel& m = *h->chunk;
  const b200asr_chunk_config& cf = m.cfg;
  if (!feats_dev || !txt_logits_dev || n <= 0) return fail(h, "b200asr_stream_decoder_step: bad arguments");
  const int B = st->B, D = cf.dmodel;
  const int Tc2 = st->n_carry + n;
  if (Tc2 > st->Tmax || n > st->Tmax) return fail(h, "b200asr_stream_decoder_step: too many frames in one step for this state");
  const BlockGeom geo = geom_of(m);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Ctx c{h, s};
  Buffers b = buffers_of(st);
  const int par = st->par_b;
  const int blk0 = cf.enc_blocks + cf.picker_blocks;
  // ---- helper (:735-758): streaming blocks without look-ahead on the n picked frames
  ENG_CUDA(h, cudaMemcpyAsync(b.x, feats_dev, sizeof(float) * (size_t)B * n * D, cudaMemcpyDeviceToDevice, s));
  h->launches++;
  ENG_TRY(h, launch_layernorm(b.x, m.helper[0].ffn1.ln.g, m.helper[0].ffn1.ln.b, b.xn, B * n, D, cf.ln_eps, s));
  int blk = blk0;
  for (size_t i = 0; i < m.helper.size(); ++i, ++blk) {
    const LNW* next = (i + 1 < m.helper.size()) ? &m.helper[i + 1].ffn1.ln : nullptr;
    ENG_TRY(h, stream_block(c, m.helper[i], b, st->qkv, geo, B, n, st->c_help, 0, st->kv[par][blk], st->kv[par ^ 1][blk], st->glu[par][blk],
                            st->glu[par ^ 1][blk], next));
  }
  // ---- decoder input = frames still inside the look-ahead of the previous call + the new helper outputs (:829-831)
  h->launches++;
  ENG_TRY(h, launch_rows_cat(st->carry, st->n_carry, b.x, n, st->cat, B, D, s));
  const int M2 = B * Tc2;
  Buffers bb = b;
  GemmParams pp{};
  pp.A = st->cat; pp.W = m.dec_projw; pp.bias = m.dec_projb; pp.C = bb.x; pp.C2 = bb.xn; pp.M = M2; pp.N = D; pp.K = D; pp.lda = D; pp.ldc = D;
  pp.ln1_g = m.dec[0].ffn1.ln.g; pp.ln1_b = m.dec[0].ffn1.ln.b; pp.ln_eps = cf.ln_eps;
  ENG_TRY(h, gemm_p(c, pp, EPI_BIAS_LN));
  for (size_t i = 0; i < m.dec.size(); ++i, ++blk) {
    const LNW* next = (i + 1 < m.dec.size()) ? &m.dec[i + 1].ffn1.ln : nullptr;
    ENG_TRY(h, stream_block(c, m.dec[i], bb, st->qkv, geo, B, Tc2, st->c_dec, cf.dec_back, st->kv[par][blk], st->kv[par ^ 1][blk], st->glu[par][blk],
                            st->glu[par ^ 1][blk], next));
  }
  ENG_TRY(h, gemm(c, bb.x, D, m.dec_fcw, m.dec_fcb, nullptr, 0.f, st->logits_t, m.Vt_pad, M2, m.Vt_pad, D, EPI_BIAS));
  ENG_TRY(h, copy_classes(h, txt_logits_dev, st->logits_t, M2, cf.txt_classes, m.Vt_pad, s));
  // ---- bookkeeping (:646-658): the first Tc2 - dec_back rows are final; the rest is carried and re-fed
  const int n_valid = Tc2 > cf.dec_back ? Tc2 - cf.dec_back : 0;
  const int n_keep = Tc2 - n_valid;
  h->launches++;
  ENG_TRY(h, launch_rows_slice(st->cat, Tc2, n_valid, n_keep, st->carry, B, D, s));
  st->n_carry = n_keep;
  st->par_b ^= 1;
  st->c_help = clampi(st->c_help + n, 0, cf.win_front);
  st->c_dec = clampi(st->c_dec + Tc2 - cf.dec_back, 0, cf.win_front);
  if (n_rows_host) *n_rows_host = Tc2;
  if (n_valid_host) *n_valid_host = n_valid;
  return 0;
}

}  // extern "C"
