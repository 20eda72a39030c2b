// Sparse MoE feed-forward launcher (moe.cu).
#pragma once
#include "common.cuh"

namespace ymt3 {

struct MoEWeights {
  const float* gate = nullptr;  // (E, D) router, always fp32
  const void* w13 = nullptr;    // (E, 2*I, D) compute dtype, rows interleaved: 2j = w1[j] (activated), 2j+1 = w3[j]
  const void* w2 = nullptr;     // (E, D, I)
  int D = 0, I = 0, E = 0, topk = 0, act = 0;
};

size_t moe_workspace_bytes(int64_t N, int D, int I, int E, int topk, int dtype);
// out = residual + moe(x); x/out/residual: (N, D) in `precision`
int moe_forward(int precision, const void* x, const void* residual, void* out, int64_t N, const MoEWeights& w,
                void* workspace, cudaStream_t stream);

}  // namespace ymt3
