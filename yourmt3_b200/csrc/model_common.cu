#include "model_common.cuh"

namespace ymt3 {

// dst[(row_map(r)), :] = convert(src[r, :])
template <typename D>
__global__ void __launch_bounds__(256)
pack_rows_kernel(const float* __restrict__ src, D* __restrict__ dst, int64_t rows, int K, int64_t row_offset,
                 int row_stride, const float* __restrict__ col_scale = nullptr) {
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= rows * K) return;
  int64_t r = i / K;
  int c = (int)(i - r * K);
  const float v = col_scale ? src[i] * col_scale[c] : src[i];
  D* d = dst + (row_offset + r * row_stride) * K + c;
  if constexpr (sizeof(D) == 4) *d = v; else *d = __float2bfloat16(v);
}

int pack_rows(DevicePool& pool, const std::vector<const ymt3_tensor_t*>& srcs, int K, int dtype,
              bool interleave2, Linear* out, cudaStream_t stream, const float* col_scale) {
  int64_t total = 0;
  for (auto* s : srcs) {
    if (!s) return YMT3_ERR_INVALID;  // error already set by TensorTable::require
    total += s->shape[0];
  }
  YMT3_REQUIRE(!interleave2 || (srcs.size() == 2 && srcs[0]->shape[0] == srcs[1]->shape[0]),
               "pack_rows: interleave needs two equal sources");
  void* W = pool.alloc((size_t)total * K * dtype_size(dtype));
  if (!W) return YMT3_ERR_CUDA;
  int64_t off = 0;
  for (size_t si = 0; si < srcs.size(); ++si) {
    const int64_t rows = srcs[si]->shape[0];
    const int64_t n = rows * K;
    const unsigned grid = (unsigned)((n + 255) / 256);
    const int64_t row_offset = interleave2 ? (int64_t)si : off;
    const int row_stride = interleave2 ? 2 : 1;
    if (dtype == YMT3_F32)
      pack_rows_kernel<float><<<grid, 256, 0, stream>>>((const float*)srcs[si]->data, (float*)W, rows, K,
                                                        row_offset, row_stride, col_scale);
    else
      pack_rows_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const float*)srcs[si]->data, (__nv_bfloat16*)W,
                                                                rows, K, row_offset, row_stride, col_scale);
    off += rows;
  }
  YMT3_CUDA_CHECK(cudaGetLastError());
  out->W = W;
  out->N = (int)total;
  out->K = K;
  return YMT3_OK;
}

int pack_rows_at(void* W, int64_t row0, int row_stride, const ymt3_tensor_t* src, int K, int dtype, cudaStream_t stream,
                 const float* col_scale) {
  if (!src) return YMT3_ERR_INVALID;
  const int64_t rows = src->shape[0];
  const int64_t n = rows * K;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (dtype == YMT3_F32)
    pack_rows_kernel<float><<<grid, 256, 0, stream>>>((const float*)src->data, (float*)W, rows, K, row0, row_stride,
                                                      col_scale);
  else
    pack_rows_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const float*)src->data, (__nv_bfloat16*)W, rows, K, row0,
                                                              row_stride, col_scale);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

int pack_vec(DevicePool& pool, const std::vector<const ymt3_tensor_t*>& srcs, bool interleave2, float** out,
             cudaStream_t stream) {
  int64_t total = 0;
  for (auto* s : srcs) {
    if (!s) return YMT3_ERR_INVALID;
    total += s->shape[0];
  }
  float* v = (float*)pool.alloc((size_t)total * 4);
  if (!v) return YMT3_ERR_CUDA;
  int64_t off = 0;
  for (size_t si = 0; si < srcs.size(); ++si) {
    const int64_t rows = srcs[si]->shape[0];
    const unsigned grid = (unsigned)((rows + 255) / 256);
    pack_rows_kernel<float><<<grid, 256, 0, stream>>>((const float*)srcs[si]->data, v, rows, 1,
                                                      interleave2 ? (int64_t)si : off, interleave2 ? 2 : 1);
    off += rows;
  }
  YMT3_CUDA_CHECK(cudaGetLastError());
  *out = v;
  return YMT3_OK;
}

int pack_table(DevicePool& pool, const float* src, bool src_on_host, int64_t numel, int dtype, void** out,
               cudaStream_t stream) {
  void* d = pool.alloc((size_t)numel * dtype_size(dtype));
  if (!d) return YMT3_ERR_CUDA;
  const float* dev_src = src;
  if (src_on_host) {
    float* tmp = (float*)pool.alloc((size_t)numel * 4);
    if (!tmp) return YMT3_ERR_CUDA;
    YMT3_CUDA_CHECK(cudaMemcpyAsync(tmp, src, (size_t)numel * 4, cudaMemcpyHostToDevice, stream));
    YMT3_CUDA_CHECK(cudaStreamSynchronize(stream));  // src may be a temporary host vector
    dev_src = tmp;
  }
  int rc = convert(dev_src, YMT3_F32, d, dtype, numel, stream);
  if (rc) return rc;
  *out = d;
  return YMT3_OK;
}

int linear_fwd(int precision, const void* x, int64_t ldx, const Linear& lin, void* y, int64_t ldy, int M,
               int act, int gated, const void* residual, int64_t ldr, float out_scale, int out_dtype,
               cudaStream_t stream, const NormFuse& nf) {
  GemmParams p{};
  p.norm_ss_in = nf.ss_in; p.norm_ss_chunks = nf.chunks; p.norm_eps = nf.eps; p.ss_out = nf.ss_out;
  p.argmax_out = nf.argmax_out; p.argmax_n = nf.argmax_n;
  p.A = x; p.lda = ldx;
  p.W = lin.W; p.ldw = lin.K;
  p.C = y; p.ldc = ldy;
  p.bias = lin.bias;
  p.residual = residual; p.ldr = ldr;
  p.M = M; p.N = lin.N; p.K = lin.K;
  p.act = act; p.gated = gated;
  p.out_scale = out_scale;
  if (precision == YMT3_F32) {
    YMT3_REQUIRE(out_dtype == YMT3_F32, "linear_fwd: f32 path writes f32");
    return gemm_f32(p, stream);
  }
  return gemm_bf16_tc(p, out_dtype, stream);
}

}  // namespace ymt3
