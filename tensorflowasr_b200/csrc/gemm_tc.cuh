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

}  // namespace b200asr
