// Punctuation model of the reference's session layer (Inference/PythonInference/punc_recover/src/punc_recover.py:44-62 runs
// punc_recover/models/punc.onnx: token ids [1, U] -> class probabilities [1, U, 32]).  "PuncTransformer", d_model 64, 8 heads of 8:
//   x = embedding[ids] * 8 + PE;  x = ELU(Dense(x))
//   3 x { y = EncoderLayer(x);  x = ReLU(causal Conv1D_k3(y)) + x }
//   h = Dense_64(Dense_768(x));  h = EncoderLayer(h);  h = EncoderLayer(h);  softmax(Dense_32(h))
//   EncoderLayer(x): x1 = LN(x + MHA(x));  LN(x1 + Dense(ReLU(Dense(x1))))
// One sentence per call (as the reference calls it: no padding, so its mask is empty).  Exact fp32: every Dense is a launch of the
// CUDA-core GEMM (q | k | v as one N = 192 GEMM with 1/sqrt(8) folded into the q rows; residual adds in the GEMM epilogue), attention is
// the fp32 attention kernel, the causal convolution a GEMM over overlapping rows of a front-padded buffer (lda 64, K 192) that the
// layer's closing LayerNorm writes straight into.  ~47 launches for a sentence: launch-latency bound by construction.
#include "engine_internal.cuh"

namespace b200asr {

struct PuncLayer {
  const float *qkv_w, *qkv_b, *o_w, *o_b, *f1_w, *f1_b, *f2_w, *f2_b, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
};

struct PuncModel {
  const float *emb, *pe, *in_w, *in_b, *up_w, *up_b, *down_w, *down_b, *out_w, *out_b;
  const float *c_w[3], *c_b[3];
  PuncLayer layer[5];
  int vocab = 0, pe_rows = 0;
  float eps = 1e-6f;
  float* ws = nullptr;
  size_t ws_rows = 0;
};

void punc_model_free(PuncModel* m) {
  if (!m) return;
  if (m->ws) cudaFree(m->ws);
  delete m;
}

}  // namespace b200asr

namespace {

constexpr int kD = 64, kH = 8, kDh = 8, kUp = 768, kCls = 32, kTaps = 3;

// x[u, :] = emb[ids[u], :] (pre-scaled by sqrt(d_model) on the host) + pe[u, :]; an id outside the table raises the flag
__global__ void punc_embed_kernel(const int* __restrict__ ids, const float* __restrict__ emb, const float* __restrict__ pe, float* __restrict__ x,
                                  int U, int vocab, int* __restrict__ bad) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= U * kD) return;
  const int u = i / kD, c = i % kD;
  const int id = ids[u];
  if (id < 0 || id >= vocab) {
    if (c == 0) atomicExch(bad, 1);
    x[i] = 0.f;
    return;
  }
  x[i] = emb[(size_t)id * kD + c] + pe[(size_t)u * kD + c];
}

__global__ void punc_elu_kernel(float* __restrict__ x, int n) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = x[i];
    x[i] = v > 0.f ? v : expm1f(v);
  }
}

// y = a + b (y may alias a or b)
__global__ void punc_add_kernel(const float* a, const float* b, float* y, int n) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] + b[i];
}

// one warp per row of 32 classes
__global__ void punc_softmax_kernel(const float* __restrict__ z, float* __restrict__ p, int U) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (row >= U) return;
  const float v = z[(size_t)row * kCls + lane];
  float m = v;
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float e = expf(v - m);
  float s = e;
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  p[(size_t)row * kCls + lane] = e / s;
}

}  // namespace

extern "C" {

B200ASR_API int b200asr_punc_create(const void* weight_blob, size_t blob_bytes, float ln_eps, int device, b200asr_handle* out) {
  if (!weight_blob || !out) return fail(nullptr, "b200asr_punc_create: null argument");
  *out = nullptr;
  b200asr_engine* h = nullptr;
  if (b200asr::engine_alloc(weight_blob, blob_bytes, device, "b200asr_punc_create", &h)) return 1;
  memset(&h->cfg, 0, sizeof(h->cfg));
  h->cfg.abi_version = B200ASR_ABI_VERSION;
  h->cfg.precision = B200ASR_PRECISION_FP32;
  b200asr::PuncModel* m = new b200asr::PuncModel();
  h->punc = m;
  m->eps = ln_eps;
  bool ok = true;
  auto bail = [&]() {
    std::string e = g_errbuf;
    b200asr_destroy(h);
    snprintf(g_errbuf, sizeof(g_errbuf), "%s", e.c_str());
    return 1;
  };
  auto it = h->tensors.find("emb");
  auto ip = h->tensors.find("pe");
  if (it == h->tensors.end() || ip == h->tensors.end() || it->second.second % kD != 0 || ip->second.second % kD != 0) {
    snprintf(g_errbuf, sizeof(g_errbuf), "b200asr_punc_create: 'emb' / 'pe' tables missing or not %d wide", kD);
    return bail();
  }
  m->emb = it->second.first; m->vocab = (int)(it->second.second / kD);
  m->pe = ip->second.first; m->pe_rows = (int)(ip->second.second / kD);
  auto L = [&](const std::string& n, uint64_t numel) { return ok ? lookup(h, n, numel, &ok) : nullptr; };
  m->in_w = L("in.w", kD * kD); m->in_b = L("in.b", kD);
  m->up_w = L("up.w", (uint64_t)kUp * kD); m->up_b = L("up.b", kUp);
  m->down_w = L("down.w", (uint64_t)kD * kUp); m->down_b = L("down.b", kD);
  m->out_w = L("out.w", kCls * kD); m->out_b = L("out.b", kCls);
  for (int i = 0; i < 3; ++i) {
    m->c_w[i] = L("c" + std::to_string(i) + ".w", (uint64_t)kD * kTaps * kD);
    m->c_b[i] = L("c" + std::to_string(i) + ".b", kD);
  }
  for (int i = 0; i < 5; ++i) {
    const std::string p = "l" + std::to_string(i) + ".";
    b200asr::PuncLayer& l = m->layer[i];
    l.qkv_w = L(p + "qkv.w", 3ull * kD * kD); l.qkv_b = L(p + "qkv.b", 3 * kD);
    l.o_w = L(p + "o.w", kD * kD); l.o_b = L(p + "o.b", kD);
    l.f1_w = L(p + "f1.w", kD * kD); l.f1_b = L(p + "f1.b", kD);
    l.f2_w = L(p + "f2.w", kD * kD); l.f2_b = L(p + "f2.b", kD);
    l.ln1_g = L(p + "ln1.g", kD); l.ln1_b = L(p + "ln1.b", kD);
    l.ln2_g = L(p + "ln2.g", kD); l.ln2_b = L(p + "ln2.b", kD);
  }
  if (!ok) return bail();
  *out = h;
  return 0;
}

B200ASR_API int b200asr_punc_infer(b200asr_handle h, const int32_t* ids_dev, int U, float* probs_dev, void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (!h->punc) return fail(h, "b200asr_punc_infer: not a punctuation handle (use b200asr_punc_create)");
  if (!ids_dev || !probs_dev) return fail(h, "b200asr_punc_infer: null buffer");
  b200asr::PuncModel& m = *h->punc;
  if (U <= 0 || U > m.pe_rows) return fail(h, "b200asr_punc_infer: the sentence must hold 1 .. pe_rows tokens");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if ((size_t)U > m.ws_rows) {
    ENG_CUDA(h, cudaDeviceSynchronize());
    if (m.ws) ENG_CUDA(h, cudaFree(m.ws));
    m.ws = nullptr;
    m.ws_rows = 0;
    const size_t rows = (size_t)U + U / 2 + 64;
    // x, a, b [rows, 64] each, padded conv input [rows + 2, 64], wide [rows, 768], one flag word
    ENG_CUDA(h, cudaMalloc(&m.ws, sizeof(float) * (rows * (3 * kD + kUp) + (rows + 2) * kD + 64)));
    m.ws_rows = rows;
  }
  const size_t R = m.ws_rows;
  float* x = m.ws;                          // the residual stream of the convolutional encoder
  float* a = x + R * kD;
  float* b = a + R * kD;
  float* cpad = b + R * kD;                 // [2 + U, 64]: two zero frames in front (the graph's causal Pad)
  float* wide = cpad + (R + 2) * kD;        // qkv [U, 192] / FFN hidden [U, 64] / bottleneck [U, 768] / logits [U, 32]
  int* bad = reinterpret_cast<int*>(wide + R * kUp);
  Ctx c{h, s};
  const int n = U * kD;
  auto blocks = [](int total) { return dim3((unsigned)((total + 255) / 256)); };
  ENG_CUDA(h, cudaMemsetAsync(cpad, 0, sizeof(float) * 2 * kD, s));
  ENG_CUDA(h, cudaMemsetAsync(bad, 0, sizeof(int), s));
  ENG_CUDA(h, launch_k(punc_embed_kernel, blocks(n), dim3(256), 0, s, ids_dev, m.emb, m.pe, a, U, m.vocab, bad));
  ENG_TRY(h, gemm(c, a, kD, m.in_w, m.in_b, nullptr, 1.f, x, kD, U, kD, kD, EPI_BIAS));
  ENG_CUDA(h, launch_k(punc_elu_kernel, blocks(n), dim3(256), 0, s, x, n));
  h->launches += 2;
  // EncoderLayer(src) -> dst; no operand of a launch aliases its output (scratch: a, b, wide)
  auto layer = [&](const b200asr::PuncLayer& l, const float* src, float* dst) -> int {
    if (gemm(c, src, kD, l.qkv_w, l.qkv_b, nullptr, 1.f, wide, 3 * kD, U, 3 * kD, kD, EPI_BIAS)) return 1;
    AttnParams ap{};
    ap.qkv = wide; ap.out = a; ap.B = 1; ap.T = U; ap.H = kH; ap.dh = kDh; ap.win_front = -1; ap.win_back = 0;
    if (attention(c, ap)) return 1;
    if (gemm(c, a, kD, l.o_w, l.o_b, src, 1.f, b, kD, U, kD, kD, EPI_RESID)) return 1;
    h->launches++;
    if (launch_layernorm(b, l.ln1_g, l.ln1_b, a, U, kD, m.eps, s)) return 1;
    if (gemm(c, a, kD, l.f1_w, l.f1_b, nullptr, 1.f, wide, kD, U, kD, kD, EPI_BIAS_RELU)) return 1;
    if (gemm(c, wide, kD, l.f2_w, l.f2_b, a, 1.f, b, kD, U, kD, kD, EPI_RESID)) return 1;
    h->launches++;
    return launch_layernorm(b, l.ln2_g, l.ln2_b, dst, U, kD, m.eps, s);
  };
  for (int i = 0; i < 3; ++i) {
    ENG_TRY(h, layer(m.layer[i], x, cpad + 2 * kD));
    ENG_TRY(h, gemm(c, cpad, kD, m.c_w[i], m.c_b[i], nullptr, 1.f, a, kD, U, kD, kTaps * kD, EPI_BIAS_RELU));
    ENG_CUDA(h, launch_k(punc_add_kernel, blocks(n), dim3(256), 0, s, (const float*)a, (const float*)x, x, n));
    h->launches++;
  }
  ENG_TRY(h, gemm(c, x, kD, m.up_w, m.up_b, nullptr, 1.f, wide, kUp, U, kUp, kD, EPI_BIAS));
  ENG_TRY(h, gemm(c, wide, kUp, m.down_w, m.down_b, nullptr, 1.f, x, kD, U, kD, kUp, EPI_BIAS));
  ENG_TRY(h, layer(m.layer[3], x, cpad + 2 * kD));
  ENG_TRY(h, layer(m.layer[4], cpad + 2 * kD, x));
  ENG_TRY(h, gemm(c, x, kD, m.out_w, m.out_b, nullptr, 1.f, wide, kCls, U, kCls, kD, EPI_BIAS));
  ENG_CUDA(h, launch_k(punc_softmax_kernel, dim3((U + 7) / 8), dim3(256), 0, s, (const float*)wide, probs_dev, U));
  h->launches++;
  int bad_host = 0;
  ENG_CUDA(h, cudaMemcpyAsync(&bad_host, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
  ENG_CUDA(h, cudaStreamSynchronize(s));
  if (bad_host) return fail(h, "b200asr_punc_infer: a token id lies outside the embedding table");
  return 0;
}

}  // extern "C"
