// placeholder until the tcgen05 kernel lands (next commit)
#include "ops.cuh"
namespace ymt3 {
int gemm_bf16_tc(const GemmParams&, int, cudaStream_t) {
  ymt3_set_error("gemm_bf16_tc: not built yet");
  return YMT3_ERR_UNSUPPORTED;
}
}  // namespace ymt3
