// TEST INFRASTRUCTURE ONLY (oracle).  Never linked into, imported by or called from the product path.
//
// Thin C shim over the reference's *own* vendored ONNX Runtime 1.10.0
// (/root/reference/Inference/CppInference/onnx/ext/onnxruntime) so Python tests / bench.py can run the
// reference's shipped ONNX graphs exactly the way the reference's deployment code does
// (Inference/CppInference/onnx/src/core/asr_session.cpp:77-122 -- one float input "inputs", one output).
// Extra graph outputs ("taps") can be requested by name when the model file was patched to expose them.
#include <onnxruntime_cxx_api.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {
struct OrtRef {
  Ort::Env env{ORT_LOGGING_LEVEL_ERROR, "ortref"};
  Ort::SessionOptions opts;
  Ort::Session* session = nullptr;
  std::string err;
};
}  // namespace

extern "C" {

void* ortref_open(const char* model_path, int intra_threads) {
  auto* h = new OrtRef();
  try {
    h->opts.SetIntraOpNumThreads(intra_threads > 0 ? intra_threads : 1);
    h->opts.SetInterOpNumThreads(1);
    h->opts.SetGraphOptimizationLevel(GraphOptimizationLevel::ORT_ENABLE_ALL);
    h->session = new Ort::Session(h->env, model_path, h->opts);
  } catch (const std::exception& e) {
    fprintf(stderr, "ortref_open(%s): %s\n", model_path, e.what());
    delete h;
    return nullptr;
  }
  return h;
}

void ortref_close(void* hv) {
  auto* h = static_cast<OrtRef*>(hv);
  if (!h) return;
  delete h->session;
  delete h;
}

const char* ortref_error(void* hv) { return static_cast<OrtRef*>(hv)->err.c_str(); }

// Run with up to two inputs.  dtype: 0 = float32, 1 = int32.  The output tensor (float32) is copied into a
// malloc'ed buffer (*out, free with ortref_free); its dims go to out_dims[0..*out_nd).
int ortref_run(void* hv, int n_inputs, const char** in_names, const void** in_data, const int64_t** in_dims,
               const int* in_nd, const int* in_dtype, const char* out_name, float** out, int64_t* out_dims,
               int* out_nd) {
  auto* h = static_cast<OrtRef*>(hv);
  try {
    auto mem = Ort::MemoryInfo::CreateCpu(OrtArenaAllocator, OrtMemTypeDefault);
    std::vector<Ort::Value> inputs;
    for (int i = 0; i < n_inputs; ++i) {
      size_t count = 1;
      for (int d = 0; d < in_nd[i]; ++d) count *= static_cast<size_t>(in_dims[i][d]);
      if (in_dtype[i] == 0) {
        inputs.push_back(Ort::Value::CreateTensor<float>(mem, const_cast<float*>(static_cast<const float*>(in_data[i])),
                                                         count, in_dims[i], in_nd[i]));
      } else {
        inputs.push_back(Ort::Value::CreateTensor<int32_t>(
            mem, const_cast<int32_t*>(static_cast<const int32_t*>(in_data[i])), count, in_dims[i], in_nd[i]));
      }
    }
    const char* out_names[1] = {out_name};
    auto outs = h->session->Run(Ort::RunOptions{nullptr}, in_names, inputs.data(), inputs.size(), out_names, 1);
    auto info = outs[0].GetTensorTypeAndShapeInfo();
    auto shape = info.GetShape();
    size_t count = info.GetElementCount();
    *out_nd = static_cast<int>(shape.size());
    for (size_t d = 0; d < shape.size(); ++d) out_dims[d] = shape[d];
    *out = static_cast<float*>(malloc(count * sizeof(float) + 16));
    memcpy(*out, outs[0].GetTensorData<float>(), count * sizeof(float));
  } catch (const std::exception& e) {
    h->err = e.what();
    return 1;
  }
  return 0;
}

void ortref_free(float* p) { free(p); }

}  // extern "C"
