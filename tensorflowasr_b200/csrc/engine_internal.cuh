// Internals shared by the engine translation units (engine.cu: offline / block-streaming path; chunk_engine.cu: ChunkConformer
// state-cache streaming): the handle type, weight lookup, the GEMM / chained-GEMM dispatch helpers and the CUDA-graph cache.
// Everything except struct b200asr_engine lives in an anonymous namespace: each translation unit gets its own copy.
#pragma once
#include "../../include/b200asr.h"
#include "kernels.cuh"
#include "gemm_tc.cuh"

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <string>
#include <vector>

using namespace b200asr;

namespace {

struct BlobEntry {
  char name[48];
  uint64_t offset;
  uint64_t numel;
};

struct LNW { const float *g, *b; };
struct FFNW { LNW ln; const float *w1, *b1, *w2, *b2; };
struct MHSAW { LNW ln; const float *wqkv, *wo, *bo; const float* bqkv = nullptr; /* Keras MultiHeadAttention q/k/v biases (ChunkConformer) */ };
struct ConvW { LNW ln; const float *pw1w, *pw1b, *dww, *pww, *pwb, *pw2w, *pw2b; };
struct BlockW { FFNW ffn1, ffn2; MHSAW mhsa; ConvW conv; LNW ln; int kernel_size; };

struct Workspace {
  float* base = nullptr;
  size_t bytes = 0;
};

struct GraphKey {
  int kind = 0, B = 0, L = 0;
  const void *p0 = nullptr, *p1 = nullptr, *p2 = nullptr, *p3 = nullptr;
  bool operator<(const GraphKey& o) const {
    return std::tie(kind, B, L, p0, p1, p2, p3) < std::tie(o.kind, o.B, o.L, o.p0, o.p1, o.p2, o.p3);
  }
};

struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0;     // kernel launches one replay stands for
  uint64_t last_use = 0;    // LRU stamp
};

}  // namespace

namespace b200asr {
struct ChunkModel;                                  // chunk_engine.cu: ChunkConformer weights + geometry
void chunk_model_free(ChunkModel* m);
struct VadModel;                                    // vad_engine.cu: voice-activity model of the session layer
void vad_model_free(VadModel* m);
struct PuncModel;                                   // punc_engine.cu: punctuation model of the session layer
void punc_model_free(PuncModel* m);
int engine_alloc(const void* weight_blob, size_t blob_bytes, int device, const char* who, b200asr_engine** out);   // engine.cu
int engine_init_frontend(b200asr_engine* h, const void* weight_blob);                                            // engine.cu
}

struct b200asr_engine {
  b200asr_config cfg;
  int device = 0;
  char* blob_dev = nullptr;
  std::map<std::string, std::pair<const float*, uint64_t>> tensors;
  // frontend
  const float *window = nullptr, *melw = nullptr;
  float2* twiddle = nullptr;
  int *mel_lo = nullptr, *mel_hi = nullptr, *mel_off = nullptr;
  float* mel_wc = nullptr;
  int mel_nnz = 0;
  // subsampling
  const float *c1w, *c1b, *c2w, *c2b, *linw, *linb;
  const float *c2w16 = nullptr, *linw16 = nullptr;   // conv2 / subsampling-linear weights as IEEE fp16 (optional blob entries "sub.conv2.w16", "sub.lin.w16")
  bool conv_f16 = false;          // tf32 mode: conv1 writes its map in fp16 and conv2 runs kind::f16 (B200ASR_NO_CONV_F16=1 turns it off)
  bool sub_out_f16 = false;       // ... and conv2 writes ITS output in fp16 for an fp16-operand subsampling linear layer (B200ASR_NO_CONV_F16=2 turns only this off)
  std::vector<BlockW> enc_blocks, ctc_blocks;
  const float *ctc_projw, *ctc_projb, *ctc_fcw, *ctc_fcb;
  int F1 = 0, F2 = 0;  // mel bins after conv1 / conv2
  Workspace ws;
  std::map<GraphKey, GraphEntry> graphs;   // at most kMaxGraphs entries, least recently used evicted first
  uint64_t graph_clock = 0;
  std::recursive_mutex mu;                 // every C-ABI entry point locks it: a handle may be shared between host threads
  float* stage_wav = nullptr;              // b200asr_recognize_host: device staging of the waveform (grows on demand)
  size_t stage_wav_floats = 0;
  int64_t launches = 0;
  std::string err;
  bool use_chain = true;   // chained FFN / conv-tail kernel (B200ASR_NO_CHAIN=1 in the environment turns it off)
  bool use_pair = true;    // ... with the hidden dimension split across a 2-CTA cluster (B200ASR_NO_PAIR=1 turns it off)
  bool attn_async = false;      // attention stages Q / K / V^T with cp.async from the pre-rounded QKV (B200ASR_ATTN_ASYNC=1; measured: no gain, 18.2 vs 18.3 us)
  bool use_fused_sub = false;   // conv1 computed inside conv2's kernel (B200ASR_FUSED_SUB=1; always on in the chunk engine): no conv1 map in HBM
  void* beam_ws = nullptr;
  size_t beam_ws_bytes = 0;
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  TcContext tc;  // tcgen05 GEMM state (tensor-map encoder entry point etc.)
  // two-deep host pipeline (b200asr_recognize_host_submit / _collect)
  struct PipeSlot {
    float* wav = nullptr;
    int32_t *ids = nullptr, *lens = nullptr;
    size_t wav_floats = 0, id_ints = 0, len_ints = 0;
    cudaEvent_t h2d = nullptr, done = nullptr;
    bool busy = false;
  } pipe[2];
  cudaStream_t pipe_copy = nullptr, pipe_compute = nullptr;
  // b200asr_debug_encode_taps: when set, run_encoder copies the residual stream after the subsampler and after every block
  // translator (b200asr_translate)
  std::vector<BlockW> tr_blocks;
  const float *tr_emb = nullptr, *tr_fcw = nullptr, *tr_fcb = nullptr, *tr_pe = nullptr;
  int tr_pe_rows = 0;
  float* tr_ws = nullptr;
  size_t tr_ws_floats = 0;
  b200asr::ChunkModel* chunk = nullptr;   // set by b200asr_chunk_create: this handle is a ChunkConformer (state-cache streaming) engine
  b200asr::VadModel* vad = nullptr;       // set by b200asr_vad_create: this handle is the session layer's voice-activity model
  b200asr::PuncModel* punc = nullptr;     // set by b200asr_punc_create: this handle is the session layer's punctuation model
  float* tap_dst = nullptr;
  int tap_count = 0, tap_max = 0;
};

namespace {

// every entry point runs on the engine's device whatever the caller's current device is, and puts the caller's device back
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

int fail(b200asr_handle h, const char* msg) {
  if (h) h->err = msg;
  snprintf(g_errbuf, sizeof(g_errbuf), "%s", msg);
  return 1;
}
int fail_cuda(b200asr_handle h) {
  if (h) h->err = g_errbuf;
  return 1;
}

#define ENG_CUDA(h, expr)                                                                                  \
  do {                                                                                                     \
    cudaError_t _e = (expr);                                                                               \
    if (_e != cudaSuccess) {                                                                               \
      snprintf(g_errbuf, sizeof(g_errbuf), "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return fail_cuda(h);                                                                                 \
    }                                                                                                      \
  } while (0)
#define ENG_TRY(h, expr)           \
  do {                             \
    if ((expr) != 0) return fail_cuda(h); \
  } while (0)

const float* lookup(b200asr_handle h, const std::string& name, uint64_t expect_numel, bool* ok) {
  auto it = h->tensors.find(name);
  if (it == h->tensors.end()) {
    snprintf(g_errbuf, sizeof(g_errbuf), "weight blob: tensor '%s' missing", name.c_str());
    *ok = false;
    return nullptr;
  }
  if (expect_numel && it->second.second != expect_numel) {
    snprintf(g_errbuf, sizeof(g_errbuf), "weight blob: tensor '%s' has %llu elements, expected %llu", name.c_str(),
             (unsigned long long)it->second.second, (unsigned long long)expect_numel);
    *ok = false;
    return nullptr;
  }
  return it->second.first;
}

bool load_block(b200asr_handle h, const std::string& p, int D, int F, int H, int dh, int K, BlockW* w) {
  bool ok = true;
  auto L = [&](const std::string& n, uint64_t numel) { return lookup(h, p + n, numel, &ok); };
  const uint64_t uD = D, uF = F;
  FFNW* ff[2] = {&w->ffn1, &w->ffn2};
  for (int i = 0; i < 2 && ok; ++i) {
    const std::string q = std::string("ffn") + char('1' + i);
    ff[i]->ln.g = L(q + ".ln.g", uD);
    ff[i]->ln.b = L(q + ".ln.b", uD);
    ff[i]->w1 = L(q + ".w1", uF * uD);
    ff[i]->b1 = L(q + ".b1", uF);
    ff[i]->w2 = L(q + ".w2", uD * uF);
    ff[i]->b2 = L(q + ".b2", uD);
  }
  w->mhsa.ln.g = L("mhsa.ln.g", uD);
  w->mhsa.ln.b = L("mhsa.ln.b", uD);
  w->mhsa.wqkv = L("mhsa.wqkv", 3ull * H * dh * uD);
  w->mhsa.wo = L("mhsa.wo", uD * H * dh);
  w->mhsa.bo = L("mhsa.bo", uD);
  {
    auto it = h->tensors.find(p + "mhsa.bqkv");
    w->mhsa.bqkv = (it != h->tensors.end() && it->second.second == 3ull * H * dh) ? it->second.first : nullptr;
  }
  w->conv.ln.g = L("conv.ln.g", uD);
  w->conv.ln.b = L("conv.ln.b", uD);
  w->conv.pw1w = L("conv.pw1.w", 2 * uD * uD);
  w->conv.pw1b = L("conv.pw1.b", 2 * uD);
  w->conv.dww = L("conv.dw.w", (uint64_t)K * uD);
  w->conv.pww = L("conv.pw.w", 2 * uD * uD);
  w->conv.pwb = L("conv.pw.b", 2 * uD);
  w->conv.pw2w = L("conv.pw2.w", 2 * uD * uD);
  w->conv.pw2b = L("conv.pw2.b", uD);
  w->ln.g = L("ln.g", uD);
  w->ln.b = L("ln.b", uD);
  w->kernel_size = K;
  return ok;
}

constexpr int kPowerStride = 520;

struct Buffers {
  float *power, *mel, *c1, *c2, *x, *xn, *h, *att, *g, *logits;
  unsigned int* pmax;
  int *am, *ids, *lens;
  float2* amp;   // per-(frame, N tile) (max, argmax) partials of the fused CTC head
};

// ------------------------------------------------------------------------------------------------ schedule
struct Ctx {
  b200asr_handle h;
  cudaStream_t s;
};

// round_out: the output is only ever read as a tensor-core operand again (store it rounded to nearest tf32)
int gemm(Ctx& c, const float* A, int lda, const float* W, const float* bias, const float* resid, float alpha, float* C,
         int ldc, int M, int N, int K, int epi, bool round_out = false) {
  GemmParams p{};
  p.A = A; p.W = W; p.bias = bias; p.resid = resid; p.C = C;
  p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldc = ldc; p.alpha = alpha; p.a_mode = 0;
  p.round_out = round_out ? 1 : 0;
  c.h->launches++;
  if (c.h->cfg.precision == B200ASR_PRECISION_TF32 && tc_gemm_supported(p, epi)) return launch_gemm_tc(c.h->tc, p, epi, c.s);
  return launch_gemm_simt(p, epi, c.s);
}

int attention(Ctx& c, const AttnParams& ap) {
  c.h->launches++;
  if (c.h->cfg.precision == B200ASR_PRECISION_TF32 && c.h->tc.ready && attention_tc_supported(ap)) return launch_attention_tc(ap, c.s);
  return launch_attention(ap, c.s);
}

int gemm_p(Ctx& c, const GemmParams& p, int epi) {
  c.h->launches++;
  if (c.h->cfg.precision == B200ASR_PRECISION_TF32 && tc_gemm_supported(p, epi)) return launch_gemm_tc(c.h->tc, p, epi, c.s);
  if (epi >= EPI_RESID_LN) {
    snprintf(g_errbuf, sizeof(g_errbuf), "internal: fused-LayerNorm epilogue requested on an unsupported shape");
    return 1;
  }
  return launch_gemm_simt(p, epi, c.s);
}

// Can every residual GEMM of this model carry its LayerNorm(s) in the tcgen05 epilogue?
bool fused_ln_ok(b200asr_handle h) {
  const int D = h->cfg.dmodel;
  return h->cfg.precision == B200ASR_PRECISION_TF32 && h->tc.ready && (D == 64 || D == 128 || D == 144 || D == 192 || D == 256);
}

// x = resid(x) + alpha * (A.W^T + bias);  then LayerNorm(s) fused in the epilogue:
//   ln2 == null:  C = x, C2 = LN(x; ln1)                       (EPI_RESID_LN)
//   ln2 != null:  C = LN(x; ln1), C2 = LN(C; ln2) (if ln2->g)   (EPI_RESID_LN2)
int gemm_resid_ln(Ctx& c, const float* A, int K, const float* W, const float* bias, float alpha, const Buffers& b, int M, int D,
                  const LNW& ln1, const LNW* ln2, float eps) {
  GemmParams p{};
  p.A = A; p.W = W; p.bias = bias; p.resid = b.x; p.C = b.x; p.C2 = b.xn; p.M = M; p.N = D; p.K = K; p.lda = K; p.ldc = D;
  p.alpha = alpha; p.ln1_g = ln1.g; p.ln1_b = ln1.b; p.ln_eps = eps;
  if (ln2) { p.ln2_g = ln2->g; p.ln2_b = ln2->b; }
  const int epi = ln2 ? EPI_RESID_LN2 : EPI_RESID_LN;
  if (c.h->use_chain && c.h->use_pair && c.h->cfg.precision == B200ASR_PRECISION_TF32 && c.h->tc.ready) {
    // 144 -> 144 projections (attention output): cluster-pair kernel with K split across the pair
    ChainGemmParams cp{};
    cp.X = A; cp.W2 = W; cp.bias2 = bias; cp.resid = b.x; cp.C = b.x; cp.C2 = b.xn; cp.M = M; cp.K1 = K; cp.N1 = 0; cp.N2 = D; cp.ldx = K;
    cp.alpha = alpha; cp.ln1_g = ln1.g; cp.ln1_b = ln1.b; cp.ln_eps = eps;
    if (ln2) { cp.ln2_g = ln2->g; cp.ln2_b = ln2->b; }
    if (tc_pair_direct_supported(cp, epi)) {
      c.h->launches++;
      return launch_gemm_chain_pair(c.h->tc, cp, epi, c.s);
    }
  }
  return gemm_p(c, p, epi);
}

// x = x + alpha * (swish(X.W1^T + b1).W2^T + b2) with the LayerNorm epilogue(s), as ONE chained kernel when supported,
// else as two GEMMs through the wide scratch buffer b.h.
int chain_resid_ln(Ctx& c, const float* X, int K1, const float* W1, const float* b1, int N1, const float* W2, const float* b2, float alpha,
                   const Buffers& b, int M, int D, const LNW& ln1, const LNW* ln2, float eps, bool round_c = false) {
  ChainGemmParams cp{};
  cp.round_c = round_c ? 1 : 0;
  cp.X = X; cp.W1 = W1; cp.bias1 = b1; cp.W2 = W2; cp.bias2 = b2; cp.resid = b.x; cp.C = b.x; cp.C2 = b.xn; cp.M = M; cp.K1 = K1;
  cp.N1 = N1; cp.N2 = D; cp.ldx = K1; cp.alpha = alpha; cp.ln1_g = ln1.g; cp.ln1_b = ln1.b; cp.ln_eps = eps;
  if (ln2) { cp.ln2_g = ln2->g; cp.ln2_b = ln2->b; }
  const int epi = ln2 ? EPI_RESID_LN2 : EPI_RESID_LN;
  if (c.h->use_chain && c.h->use_pair && tc_chain_pair_supported(cp, epi)) {
    c.h->launches++;
    return launch_gemm_chain_pair(c.h->tc, cp, epi, c.s);
  }
  if (c.h->use_chain && tc_chain_supported(cp, epi)) {
    c.h->launches++;
    return launch_gemm_chain(c.h->tc, cp, epi, c.s);
  }
  if (gemm(c, X, K1, W1, b1, nullptr, 0.f, b.h, N1, M, N1, K1, EPI_BIAS_SWISH, true)) return 1;
  return gemm_resid_ln(c, b.h, N1, W2, b2, alpha, b, M, D, ln1, ln2, eps);
}

// Run `body(stream)` either directly or through a cached CUDA graph keyed on shapes + pointers.  Graphs cannot be
// captured on the legacy default stream, so calls that arrive on it are bridged (event in / event out) onto a
// private stream; ordering with respect to the caller's stream is preserved.
template <class Body>
int with_graph(b200asr_handle h, cudaStream_t s, const GraphKey& key, Body body) {
  if (!h->cfg.use_cuda_graph) return body(s);
  cudaStream_t rs = s;
  const bool bridged = (s == nullptr || s == cudaStreamLegacy || s == cudaStreamPerThread);
  if (bridged) {
    if (!h->own_stream) {
      ENG_CUDA(h, cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
      ENG_CUDA(h, cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
      ENG_CUDA(h, cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
    }
    rs = h->own_stream;
    ENG_CUDA(h, cudaEventRecord(h->ev_in, s));
    ENG_CUDA(h, cudaStreamWaitEvent(rs, h->ev_in, 0));
  }
  auto it = h->graphs.find(key);
  if (it == h->graphs.end()) {
    constexpr size_t kMaxGraphs = 32;
    if (h->graphs.size() >= kMaxGraphs) {   // evict the least recently used graph only (a serving loop that rotates through a few
      auto victim = h->graphs.begin();      // buffer sets keeps its hot graphs)
      for (auto g = h->graphs.begin(); g != h->graphs.end(); ++g)
        if (g->second.last_use < victim->second.last_use) victim = g;
      // the graph may still be executing on a stream: destroying an exec graph is deferred by the runtime until it has finished
      cudaGraphExecDestroy(victim->second.exec);
      h->graphs.erase(victim);
    }
    cudaGraph_t graph = nullptr;
    const int64_t before = h->launches;
    ENG_CUDA(h, cudaStreamBeginCapture(rs, cudaStreamCaptureModeThreadLocal));
    int rc = body(rs);
    const int64_t graph_launches = h->launches - before;
    h->launches = before;
    cudaError_t e = cudaStreamEndCapture(rs, &graph);
    if (rc != 0) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (e != cudaSuccess) {
      snprintf(g_errbuf, sizeof(g_errbuf), "cudaStreamEndCapture: %s", cudaGetErrorString(e));
      return fail_cuda(h);
    }
    if (getenv("B200ASR_GRAPH_DBG")) {   // how many kernel->kernel edges were captured as programmatic (PDL) dependencies?
      size_t ne = 0;
      if (cudaGraphGetEdges_v2(graph, nullptr, nullptr, nullptr, &ne) == cudaSuccess && ne > 0) {
        std::vector<cudaGraphNode_t> from(ne), to(ne);
        std::vector<cudaGraphEdgeData> ed(ne);
        size_t nprog = 0;
        if (cudaGraphGetEdges_v2(graph, from.data(), to.data(), ed.data(), &ne) == cudaSuccess)
          for (size_t i = 0; i < ne; ++i) nprog += (ed[i].type == cudaGraphDependencyTypeProgrammatic);
        fprintf(stderr, "b200asr graph: %zu edges, %zu programmatic\n", ne, nprog);
      }
    }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
      snprintf(g_errbuf, sizeof(g_errbuf), "cudaGraphInstantiate: %s", cudaGetErrorString(e));
      return fail_cuda(h);
    }
    GraphEntry ge;
    ge.exec = exec;
    ge.launches = graph_launches;
    it = h->graphs.emplace(key, ge).first;
  }
  it->second.last_use = ++h->graph_clock;
  ENG_CUDA(h, cudaGraphLaunch(it->second.exec, rs));
  h->launches += it->second.launches;
  if (bridged) {
    ENG_CUDA(h, cudaEventRecord(h->ev_out, rs));
    ENG_CUDA(h, cudaStreamWaitEvent(s, h->ev_out, 0));
  }
  return 0;
}

}  // namespace
