// Host-side plumbing shared by the model stages: named-tensor lookup (state-dict keys),
// device buffers owned by a handle, weight packing into the handle's compute dtype.
#pragma once
#include "ops.cuh"
#include "../../include/ymt3_b200.h"
#include <string>
#include <string.h>
#include <vector>

namespace ymt3 {

inline size_t dtype_size(int dtype) { return dtype == YMT3_F32 ? 4 : 2; }

struct TensorTable {
  const ymt3_tensor_t* t;
  int n;
  const ymt3_tensor_t* find(const std::string& name) const {
    for (int i = 0; i < n; ++i)
      if (t[i].name && name == t[i].name) return &t[i];
    return nullptr;
  }
  // returns nullptr and sets the error string if missing or mis-shaped (d1 < 0: 1-D tensor)
  const ymt3_tensor_t* require(const std::string& name, int64_t d0, int64_t d1 = -1) const {
    const ymt3_tensor_t* x = find(name);
    if (!x) {
      ymt3_set_error("missing tensor '%s'", name.c_str());
      return nullptr;
    }
    const int nd = d1 < 0 ? 1 : 2;
    int64_t numel = 1;
    for (int i = 0; i < x->ndim; ++i) numel *= x->shape[i];
    const int64_t want = d0 * (d1 < 0 ? 1 : d1);
    if (x->dtype != YMT3_F32 || !x->data || numel != want || x->ndim < nd || x->shape[0] != d0) {
      ymt3_set_error("tensor '%s': expected f32 shape (%lld%s%lld), got ndim %d [%lld, %lld, ...]", name.c_str(),
                     (long long)d0, d1 < 0 ? "" : ", ", (long long)(d1 < 0 ? 0 : d1), x->ndim,
                     (long long)x->shape[0], (long long)(x->ndim > 1 ? x->shape[1] : 0));
      return nullptr;
    }
    return x;
  }
};

// cudaMalloc'ed buffers owned by a handle; freed together
struct DevicePool {
  std::vector<void*> ptrs;
  size_t bytes = 0;
  void* alloc(size_t n) {
    void* p = nullptr;
    if (n == 0) n = 16;
    if (cudaMalloc(&p, n) != cudaSuccess) {
      ymt3_set_error("cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(cudaGetLastError()));
      return nullptr;
    }
    ptrs.push_back(p);
    bytes += n;
    return p;
  }
  void release() {
    for (void* p : ptrs) cudaFree(p);
    ptrs.clear();
    bytes = 0;
  }
};

// A linear layer packed in the compute dtype: W is (N, K) row-major, bias fp32 or null.
struct Linear {
  void* W = nullptr;
  float* bias = nullptr;
  int N = 0, K = 0;
};

// Stack the rows of several (n_i, K) f32 tensors into one (sum n_i, K) weight in `dtype`.
// interleave2: exactly two sources of equal shape; output row 2j = src0[j], 2j+1 = src1[j].
// col_scale (device fp32 (K) or null): every row is multiplied element-wise by it BEFORE the conversion (RMSNorm
// weight folded into the consumer GEMM, see GemmParams::norm_ss_in)
int pack_rows(DevicePool& pool, const std::vector<const ymt3_tensor_t*>& srcs, int K, int dtype,
              bool interleave2, Linear* out, cudaStream_t stream, const float* col_scale = nullptr);
// write one (rows, K) f32 source into an existing packed weight: dst row = row0 + r * row_stride
int pack_rows_at(void* W, int64_t row0, int row_stride, const ymt3_tensor_t* src, int K, int dtype, cudaStream_t stream,
                 const float* col_scale = nullptr);
// concatenate 1-D f32 tensors (biases / norm scales) into one fp32 device vector
int pack_vec(DevicePool& pool, const std::vector<const ymt3_tensor_t*>& srcs, bool interleave2, float** out,
             cudaStream_t stream);
// copy an f32 (rows, cols) table into `dtype`
int pack_table(DevicePool& pool, const float* src_dev_or_host, bool src_on_host, int64_t numel, int dtype,
               void** out, cudaStream_t stream);

// y = epi(x @ lin.W^T + bias) dispatched on the handle's precision
// nf: fused RMSNorm hooks of the bf16 GEMM (consumer: row scale from sum-of-squares partials; producer: emit them)
struct NormFuse {
  const float* ss_in = nullptr; int chunks = 0; float eps = 0.f;
  float* ss_out = nullptr;
  unsigned long long* argmax_out = nullptr; int argmax_n = 0;   // fused greedy selection (GemmParams::argmax_out)
};
int linear_fwd(int precision, const void* x, int64_t ldx, const Linear& lin, void* y, int64_t ldy, int M,
               int act, int gated, const void* residual, int64_t ldr, float out_scale, int out_dtype,
               cudaStream_t stream, const NormFuse& nf = NormFuse());

}  // namespace ymt3
