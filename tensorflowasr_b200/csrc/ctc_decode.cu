// CTC greedy decode on the device.  Replaces the reference's host-side greedy decoders
// (Inference/CppInference/onnx/src/core/ctc_greedy_decoder.h:4-43, Inference/PythonInference/asr/src/asr.py:41-61,
// externals/ctc_decoders/ctc_greedy_decoder.cpp:4-45 and tf.keras.backend.ctc_decode(greedy=True) at test_asr.py:198):
// per-frame argmax (first maximum wins), merge repeated symbols, drop blank.
#include "kernels.cuh"

namespace b200asr {

namespace {

// warp per frame
__global__ void __launch_bounds__(256) frame_argmax_kernel(const float* __restrict__ logits, int rows, int V,
                                                           int* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* r = logits + (size_t)row * V;
  float best = -INFINITY;
  int idx = 0x7fffffff;
  for (int v = lane; v < V; v += 32) {
    const float x = r[v];
    if (x > best) {  // strictly greater: lowest index among equals stays
      best = x;
      idx = v;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ob > best || (ob == best && oi < idx)) {
      best = ob;
      idx = oi;
    }
  }
  if (lane == 0) out[row] = (idx == 0x7fffffff) ? 0 : idx;
}

// thread per frame: first maximum over the N tiles' partial (max, argmax) pairs (tiles ascend with the class index, so "strictly
// greater" keeps the lowest class among equals, as ctc_greedy_decoder.h:11-18 does)
__global__ void __launch_bounds__(256) argmax_combine_kernel(const float2* __restrict__ part, int rows, int n_tiles, int* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * 256 + threadIdx.x;
  if (row >= rows) return;
  float best = -INFINITY;
  int idx = 0;
  for (int t = 0; t < n_tiles; ++t) {
    const float2 v = part[(size_t)row * n_tiles + t];
    if (v.x > best) {
      best = v.x;
      idx = __float_as_int(v.y);
    }
  }
  out[row] = idx;
}

// warp per utterance: ids[b, :] = collapsed sequence padded with -1, out_len[b] = its length
__global__ void __launch_bounds__(32) ctc_collapse_kernel(const int* __restrict__ am, const int* __restrict__ lengths,
                                                          int T, int blank, int* __restrict__ ids,
                                                          int* __restrict__ out_len) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, lane = threadIdx.x;
  const int len = lengths ? min(lengths[b], T) : T;
  const int* a = am + (size_t)b * T;
  int* o = ids + (size_t)b * T;
  int count = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    bool keep = false;
    int sym = -1;
    if (t < len) {
      sym = a[t];
      const int prev = (t > 0) ? a[t - 1] : -1;
      keep = (sym != prev) && (sym != blank);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (keep) o[count + __popc(mask & ((1u << lane) - 1u))] = sym;
    count += __popc(mask);
  }
  for (int t = count + lane; t < T; t += 32) o[t] = -1;
  if (lane == 0) out_len[b] = count;
}

}  // namespace

int launch_ctc_greedy(const float* logits, const int* lengths, int B, int T, int V, int blank, int* frame_argmax, int* ids,
                      int* out_len, cudaStream_t stream) {
  if (B == 0) return 0;
  if (T > 0) {
    B200_CUDA_OK(launch_k(frame_argmax_kernel, dim3(ceil_div(B * T, 8)), dim3(256), 0, stream, logits, B * T, V, frame_argmax));
  }
  B200_CUDA_OK(launch_k(ctc_collapse_kernel, dim3(B), dim3(32), 0, stream, (const int*)frame_argmax, lengths, T, blank, ids, out_len));
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_ctc_greedy_partials(const float2* part, int n_tiles, const int* lengths, int B, int T, int blank, int* frame_argmax, int* ids,
                               int* out_len, cudaStream_t stream) {
  if (B == 0) return 0;
  if (T > 0) B200_CUDA_OK(launch_k(argmax_combine_kernel, dim3(ceil_div(B * T, 256)), dim3(256), 0, stream, part, B * T, n_tiles, frame_argmax));
  B200_CUDA_OK(launch_k(ctc_collapse_kernel, dim3(B), dim3(32), 0, stream, (const int*)frame_argmax, lengths, T, blank, ids, out_len));
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace b200asr
