// Voice-activity model of the reference's session layer (Inference/PythonInference/vad/src/vad.py:24-28 runs
// vad/models/vad.onnx: frames [1, N, 80] of the 8 kHz signal -> logits [1, N, 1]).  90 K parameters:
//   Dense 80 -> Dense 80 + ReLU -> causal Conv1D(k = 5) + ReLU -> Dense 80 + ReLU -> LayerNorm -> causal Conv1D(k = 5) + ReLU -> Dense 80 + ReLU -> Dense 1
// Every layer is one launch of the exact-fp32 CUDA-core GEMM (the session layer thresholds the logits at 0: no reduced precision here).
// Frames live in a front-padded buffer [B, 4 + N, 80]; a causal convolution is then a GEMM over OVERLAPPING rows (lda = 80, K = 5 * 80:
// row m of A is frames m .. m+4 back to back), written 4 rows further down so that its output is again front-padded.  The 4 rows per
// session that such a GEMM computes across a session boundary land exactly in the next session's pad rows, which are re-zeroed before
// the next convolution reads them.
#include "engine_internal.cuh"

namespace b200asr {

struct VadModel {
  const float *d_w[5], *d_b[5], *c_w[2], *c_b[2], *ln_g, *ln_b;
  float eps = 1e-3f;
  float* ws = nullptr;       // two activation buffers [R + 4, 80] + the padded output column block [R, 4]
  size_t ws_rows = 0;
};

void vad_model_free(VadModel* m) {
  if (!m) return;
  if (m->ws) cudaFree(m->ws);
  delete m;
}

}  // namespace b200asr

namespace {

constexpr int kF = 80, kPad = 4, kTaps = 5;

// X[b, 4 + n, c] = wav[b, (n * 80 + c) * stride]; rows 0..3 of every session = 0
__global__ void vad_pack_kernel(const float* __restrict__ wav, float* __restrict__ X, int B, int N, int stride) {
  pdl_trigger();
  pdl_wait();      // (launched with the programmatic-dependent-launch attribute like every kernel here: wait for the producer's writes)
  const size_t total = (size_t)B * (N + kPad) * kF;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % kF);
    const size_t r = i / kF;
    const int j = (int)(r % (N + kPad));
    const size_t b = r / (N + kPad);
    X[i] = j < kPad ? 0.f : wav[(b * (size_t)N * kF + (size_t)(j - kPad) * kF + c) * stride];
  }
}

__global__ void vad_zero_pad_kernel(float* __restrict__ X, int B, int N) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * kPad * kF) return;
  const int b = i / (kPad * kF), k = i % (kPad * kF);
  X[(size_t)b * (N + kPad) * kF + k] = 0.f;
}

// logits[b, n] = Y[(b * (N + 4) + 4 + n) * 4]
__global__ void vad_unpack_kernel(const float* __restrict__ Y, float* __restrict__ logits, int B, int N) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const int b = i / N, n = i % N;
  logits[i] = Y[((size_t)b * (N + kPad) + kPad + n) * 4];
}

}  // namespace

extern "C" {

B200ASR_API int b200asr_vad_create(const void* weight_blob, size_t blob_bytes, float ln_eps, int device, b200asr_handle* out) {
  if (!weight_blob || !out) return fail(nullptr, "b200asr_vad_create: null argument");
  *out = nullptr;
  b200asr_engine* h = nullptr;
  if (b200asr::engine_alloc(weight_blob, blob_bytes, device, "b200asr_vad_create", &h)) return 1;
  memset(&h->cfg, 0, sizeof(h->cfg));
  h->cfg.abi_version = B200ASR_ABI_VERSION;
  h->cfg.precision = B200ASR_PRECISION_FP32;
  b200asr::VadModel* m = new b200asr::VadModel();
  h->vad = m;
  m->eps = ln_eps;
  bool ok = true;
  for (int i = 0; i < 5 && ok; ++i) {
    const int n_out = i == 4 ? 4 : kF;
    m->d_w[i] = lookup(h, ("d" + std::to_string(i) + ".w").c_str(), (uint64_t)n_out * kF, &ok);
    m->d_b[i] = ok ? lookup(h, ("d" + std::to_string(i) + ".b").c_str(), n_out, &ok) : nullptr;
  }
  for (int i = 0; i < 2 && ok; ++i) {
    m->c_w[i] = lookup(h, ("c" + std::to_string(i) + ".w").c_str(), (uint64_t)kF * kTaps * kF, &ok);
    m->c_b[i] = ok ? lookup(h, ("c" + std::to_string(i) + ".b").c_str(), kF, &ok) : nullptr;
  }
  m->ln_g = ok ? lookup(h, "ln.g", kF, &ok) : nullptr;
  m->ln_b = ok ? lookup(h, "ln.b", kF, &ok) : nullptr;
  if (!ok) {
    std::string e = g_errbuf;
    b200asr_destroy(h);
    snprintf(g_errbuf, sizeof(g_errbuf), "%s", e.c_str());
    return 1;
  }
  *out = h;
  return 0;
}

B200ASR_API int b200asr_vad_infer(b200asr_handle h, const float* wav_dev, int B, int N, int stride, float* logits_dev, void* stream) {
  if (!h) return 1;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard dev_guard(h->device);
  if (!h->vad) return fail(h, "b200asr_vad_infer: not a VAD handle (use b200asr_vad_create)");
  if (!wav_dev || !logits_dev) return fail(h, "b200asr_vad_infer: null buffer");
  if (B <= 0 || N <= 0 || (stride != 1 && stride != 2)) return fail(h, "b200asr_vad_infer: B, N must be positive and stride 1 or 2");
  if ((size_t)B * (N + kPad) > (size_t)1 << 26) return fail(h, "b200asr_vad_infer: too many frames in one call");
  b200asr::VadModel& m = *h->vad;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int P = N + kPad, R = B * P;
  if ((size_t)R > m.ws_rows) {
    ENG_CUDA(h, cudaDeviceSynchronize());
    if (m.ws) ENG_CUDA(h, cudaFree(m.ws));
    m.ws = nullptr;
    m.ws_rows = 0;
    const size_t rows = (size_t)R + R / 4 + 64;
    ENG_CUDA(h, cudaMalloc(&m.ws, sizeof(float) * (2 * (rows + kPad) * kF + rows * 4)));
    m.ws_rows = rows;
  }
  float* a = m.ws;
  float* b = a + (m.ws_rows + kPad) * kF;
  float* y = b + (m.ws_rows + kPad) * kF;
  Ctx c{h, s};
  const size_t total = (size_t)R * kF;
  ENG_CUDA(h, launch_k(vad_pack_kernel, dim3((unsigned)std::min<size_t>((total + 255) / 256, 4096)), dim3(256), 0, s, wav_dev, a, B, N, stride));
  h->launches++;
  auto zero_pad = [&](float* x) {
    h->launches++;
    return launch_k(vad_zero_pad_kernel, dim3((B * kPad * kF + 255) / 256), dim3(256), 0, s, x, B, N);
  };
  // (pad rows pick up biases in the dense layers: harmless, they are zeroed again before a convolution reads them)
  ENG_TRY(h, gemm(c, a, kF, m.d_w[0], m.d_b[0], nullptr, 1.f, b, kF, R, kF, kF, EPI_BIAS));
  ENG_TRY(h, gemm(c, b, kF, m.d_w[1], m.d_b[1], nullptr, 1.f, a, kF, R, kF, kF, EPI_BIAS_RELU));
  ENG_CUDA(h, zero_pad(a));
  ENG_TRY(h, gemm(c, a, kF, m.c_w[0], m.c_b[0], nullptr, 1.f, b + kPad * kF, kF, R - kPad, kF, kTaps * kF, EPI_BIAS_RELU));
  ENG_TRY(h, gemm(c, b, kF, m.d_w[2], m.d_b[2], nullptr, 1.f, a, kF, R, kF, kF, EPI_BIAS_RELU));
  h->launches++;
  ENG_TRY(h, launch_layernorm(a, m.ln_g, m.ln_b, b, R, kF, m.eps, s));
  ENG_CUDA(h, zero_pad(b));
  ENG_TRY(h, gemm(c, b, kF, m.c_w[1], m.c_b[1], nullptr, 1.f, a + kPad * kF, kF, R - kPad, kF, kTaps * kF, EPI_BIAS_RELU));
  ENG_TRY(h, gemm(c, a, kF, m.d_w[3], m.d_b[3], nullptr, 1.f, b, kF, R, kF, kF, EPI_BIAS_RELU));
  ENG_TRY(h, gemm(c, b, kF, m.d_w[4], m.d_b[4], nullptr, 1.f, y, 4, R, 4, kF, EPI_BIAS));
  h->launches++;
  ENG_CUDA(h, launch_k(vad_unpack_kernel, dim3((B * N + 255) / 256), dim3(256), 0, s, (const float*)y, logits_dev, B, N));
  return 0;
}

}  // extern "C"
