// Mixtral-style sparse MoE feed-forward (HF modeling_mixtral.py:62-135 semantics):
//   probs = softmax_fp32(x @ Wg^T); (w, idx) = topk(probs, k); w /= sum(w)
//   out   = sum_k w_k * W2[idx_k]( act(W1[idx_k] x) * W3[idx_k] x )
// The reference loops over experts in Python with gather / index_add_.  Here: one fused
// router + top-k kernel, a device-side scan, one scatter that sorts token copies by expert
// (no host sync: offsets stay on the device), two GROUPED GEMMs over the expert-sorted rows
// (gated-activation epilogue; per-row routing weight applied in the second epilogue), and a
// gather-combine that also adds the residual.  Results are independent of the atomic slot
// order, so the op is deterministic.
#include "model_common.cuh"
#include "moe.cuh"

namespace ymt3 {

template <typename T> __device__ __forceinline__ float moe_to_f(T v);
template <> __device__ __forceinline__ float moe_to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float moe_to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// one warp per token; E <= 32, topk <= 4
template <typename T>
__global__ void __launch_bounds__(256)
moe_route_kernel(const T* __restrict__ x, const float* __restrict__ Wg, int N, int D, int E, int topk,
                 int* __restrict__ top_idx, float* __restrict__ top_w, int* __restrict__ counts) {
  const int tok = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (tok >= N) return;
  const T* xr = x + (int64_t)tok * D;
  float my_logit = -INFINITY;  // lane e holds logit e
  for (int e = 0; e < E; ++e) {
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc = fmaf(moe_to_f(xr[d]), Wg[e * D + d], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == e) my_logit = acc;
  }
  // softmax over the E lanes
  float mx = my_logit;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float p = lane < E ? expf(my_logit - mx) : 0.f;
  float sum = p;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  p /= sum;
  // top-k by repeated arg-max (ties -> lowest expert index)
  float sel_w[4];
  int sel_i[4];
  float wsum = 0.f;
  float cur = lane < E ? p : -1.f;
  for (int k = 0; k < topk; ++k) {
    float bv = cur;
    int bi = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    sel_w[k] = bv;
    sel_i[k] = bi;
    wsum += bv;
    if (lane == bi) cur = -1.f;
  }
  if (lane == 0) {
    for (int k = 0; k < topk; ++k) {
      top_idx[tok * topk + k] = sel_i[k];
      top_w[tok * topk + k] = sel_w[k] / wsum;
      atomicAdd(counts + sel_i[k], 1);
    }
  }
}

// offsets[0..E] = exclusive scan of counts; cursor[e] = 0
__global__ void moe_offsets_kernel(const int* __restrict__ counts, int* __restrict__ offsets, int* __restrict__ cursor,
                                   int E) {
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int e = 0; e < E; ++e) {
      offsets[e] = acc;
      acc += counts[e];
      cursor[e] = 0;
    }
    offsets[E] = acc;
  }
}

// one warp per (token, k): claim a slot in the expert's segment and copy the token row there
template <typename T>
__global__ void __launch_bounds__(256)
moe_scatter_kernel(const T* __restrict__ x, int N, int D, int topk, const int* __restrict__ top_idx,
                   const float* __restrict__ top_w, const int* __restrict__ offsets, int* __restrict__ cursor,
                   T* __restrict__ xs, float* __restrict__ slot_w, int* __restrict__ tok_slot) {
  const int item = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (item >= N * topk) return;
  const int tok = item / topk;
  int slot = 0;
  if (lane == 0) {
    const int e = top_idx[item];
    slot = offsets[e] + atomicAdd(cursor + e, 1);
    slot_w[slot] = top_w[item];
    tok_slot[item] = slot;
  }
  slot = __shfl_sync(0xffffffffu, slot, 0);
  const T* src = x + (int64_t)tok * D;
  T* dst = xs + (int64_t)slot * D;
  for (int d = lane; d < D; d += 32) dst[d] = src[d];
}

// out[tok, :] = residual[tok, :] + sum_k ys[tok_slot[tok, k], :]
template <typename T>
__global__ void __launch_bounds__(256)
moe_combine_kernel(const T* __restrict__ ys, const int* __restrict__ tok_slot, const T* __restrict__ residual,
                   T* __restrict__ out, int N, int D, int topk) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (int64_t)N * D) return;
  const int tok = (int)(i / D), d = (int)(i % D);
  float acc = residual ? moe_to_f(residual[i]) : 0.f;
  for (int k = 0; k < topk; ++k) acc += moe_to_f(ys[(int64_t)tok_slot[tok * topk + k] * D + d]);
  if constexpr (sizeof(T) == 4) out[i] = acc; else out[i] = __float2bfloat16(acc);
}

size_t moe_workspace_bytes(int64_t N, int D, int I, int E, int topk, int dtype) {
  const size_t es = dtype_size(dtype);
  const int64_t S = N * topk;
  size_t b = 0;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  b += al(S * D * es);        // xs
  b += al(S * I * es);        // hs
  b += al(S * D * es);        // ys
  b += al(S * 4) * 3;         // top_idx, top_w, tok_slot
  b += al(S * 4);             // slot_w
  b += al((3 * E + 8) * 4);   // counts, offsets, cursor
  return b;
}

int moe_forward(int precision, const void* x, const void* residual, void* out, int64_t N64, const MoEWeights& w,
                void* workspace, cudaStream_t s) {
  if (N64 <= 0) return YMT3_OK;
  YMT3_REQUIRE(w.E >= 1 && w.E <= 32 && w.topk >= 1 && w.topk <= 4 && w.topk <= w.E, "moe: bad E/topk");
  YMT3_REQUIRE(N64 * w.topk < (1ll << 31), "moe: too many tokens");
  const int N = (int)N64, D = w.D, I = w.I, E = w.E, topk = w.topk;
  const size_t es = dtype_size(precision);
  const int64_t S = (int64_t)N * topk;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  char* p = (char*)workspace;
  void* xs = p; p += al(S * D * es);
  void* hs = p; p += al(S * I * es);
  void* ys = p; p += al(S * D * es);
  int* top_idx = (int*)p; p += al(S * 4);
  float* top_w = (float*)p; p += al(S * 4);
  int* tok_slot = (int*)p; p += al(S * 4);
  float* slot_w = (float*)p; p += al(S * 4);
  int* counts = (int*)p;
  int* offsets = counts + E;
  int* cursor = offsets + E + 1;
  YMT3_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)E * 4, s));
  const unsigned gt = (unsigned)ymt3_div_up(N, 8), gs = (unsigned)ymt3_div_up(S, 8);
  if (precision == YMT3_F32) {
    moe_route_kernel<float><<<gt, 256, 0, s>>>((const float*)x, w.gate, N, D, E, topk, top_idx, top_w, counts);
    moe_offsets_kernel<<<1, 32, 0, s>>>(counts, offsets, cursor, E);
    moe_scatter_kernel<float><<<gs, 256, 0, s>>>((const float*)x, N, D, topk, top_idx, top_w, offsets, cursor,
                                                 (float*)xs, slot_w, tok_slot);
  } else {
    moe_route_kernel<__nv_bfloat16><<<gt, 256, 0, s>>>((const __nv_bfloat16*)x, w.gate, N, D, E, topk, top_idx, top_w, counts);
    moe_offsets_kernel<<<1, 32, 0, s>>>(counts, offsets, cursor, E);
    moe_scatter_kernel<__nv_bfloat16><<<gs, 256, 0, s>>>((const __nv_bfloat16*)x, N, D, topk, top_idx, top_w, offsets,
                                                         cursor, (__nv_bfloat16*)xs, slot_w, tok_slot);
  }
  YMT3_CUDA_CHECK(cudaGetLastError());
  // grouped GEMM 1: hs = act(xs W1^T) * (xs W3^T)   (rows of w13 interleaved per expert)
  GemmParams g{};
  g.A = xs; g.lda = D; g.W = w.w13; g.ldw = D; g.C = hs; g.ldc = I;
  g.M = (int)S; g.N = 2 * I; g.K = D; g.act = w.act; g.gated = 1; g.out_scale = 1.f;
  g.group_offsets = offsets; g.num_groups = E; g.strideW = (int64_t)2 * I * D;
  int rc = precision == YMT3_F32 ? gemm_f32(g, s) : gemm_bf16_tc(g, precision, s);
  if (rc) return rc;
  // grouped GEMM 2: ys = slot_w * (hs W2^T)
  GemmParams h{};
  h.A = hs; h.lda = I; h.W = w.w2; h.ldw = I; h.C = ys; h.ldc = D;
  h.M = (int)S; h.N = D; h.K = I; h.out_scale = 1.f; h.row_scale = slot_w;
  h.group_offsets = offsets; h.num_groups = E; h.strideW = (int64_t)D * I;
  rc = precision == YMT3_F32 ? gemm_f32(h, s) : gemm_bf16_tc(h, precision, s);
  if (rc) return rc;
  const unsigned gc = (unsigned)(((int64_t)N * D + 255) / 256);
  if (precision == YMT3_F32)
    moe_combine_kernel<float><<<gc, 256, 0, s>>>((const float*)ys, tok_slot, (const float*)residual, (float*)out, N, D, topk);
  else
    moe_combine_kernel<__nv_bfloat16><<<gc, 256, 0, s>>>((const __nv_bfloat16*)ys, tok_slot, (const __nv_bfloat16*)residual,
                                                         (__nv_bfloat16*)out, N, D, topk);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

}  // namespace ymt3
