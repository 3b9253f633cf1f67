// tcgen05 (5th-gen tensor core) tf32 GEMM path -- interface.  Implementation in gemm_tc.cu.
#pragma once
#include "kernels.cuh"

namespace b200asr {

struct TcContext {
  void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled entry point (resolved through the runtime, no -lcuda)
  int num_sms = 148;
  bool ready = false;
};

int tc_init(TcContext* ctx);
bool tc_gemm_supported(const GemmParams& p, int epilogue);
int launch_gemm_tc(TcContext& ctx, const GemmParams& p, int epilogue, cudaStream_t stream);

// Y = LN-epilogue( resid + alpha * ( swish(X . W1^T + b1) . W2^T + b2 ) )  in ONE kernel (gemm_chain.cu); the hidden
// activations stay in TMEM.  X [M, K1] (row stride ldx), W1 [N1, K1], W2 [N2, N1], outputs C / C2 [M, N2].
struct ChainGemmParams {
  const float *X, *W1, *bias1, *W2, *bias2, *resid;
  float *C, *C2;
  int M, K1, N1, N2, ldx;
  float alpha;
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  float ln_eps;
  int round_c = 0;   // EPI_RESID_LN2 without a second LayerNorm: store C rounded to nearest tf32 (it only feeds another GEMM)
};
bool tc_chain_supported(const ChainGemmParams& p, int epilogue);
int launch_gemm_chain(TcContext& ctx, const ChainGemmParams& p, int epilogue, cudaStream_t stream);
// same result with the hidden dimension split across a cluster of two CTAs (gemm_chain_pair.cu)
bool tc_chain_pair_supported(const ChainGemmParams& p, int epilogue);
// N1 == 0: no hidden layer, C/C2 = LN-epilogue(resid + alpha * (X . W2^T + bias2)) with K split across the pair
bool tc_pair_direct_supported(const ChainGemmParams& p, int epilogue);
int launch_gemm_chain_pair(TcContext& ctx, const ChainGemmParams& p, int epilogue, cudaStream_t stream);

// ConvSubsampling's two convolutions as one kernel (conv_sub_tc.cu): conv1 computed on the fly into conv2's implicit-GEMM A tile.
struct ConvSubParams {
  const float* mel;   // [B, T, F]
  const float* w1;    // [9, D] tap-major (conv1: 1 -> D)
  const float* b1;    // [D]
  const float* w2;    // [D, 9 D] K-major: [cout][(kh, kw, cin)]
  const float* b2;    // [D]
  float* out;         // [B, T2, F2, D]
  int B, T, F, T1, F1, T2, F2, D;
  int pt1, pf1, pt2, pf2;   // 'same' padding before (time / frequency) of conv1 and conv2
  int round_out;            // store the output rounded to nearest tf32 (it feeds the subsampling linear GEMM only)
};
bool conv_subsample_tc_supported(const ConvSubParams& p);
int launch_conv_subsample_tc(TcContext& ctx, const ConvSubParams& p, cudaStream_t stream);

}  // namespace b200asr
