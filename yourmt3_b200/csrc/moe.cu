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

// 8 warps x 8 tokens per block; E <= 8 * EG experts, topk <= 4. Router weights live in REGISTERS (lane holds
// Wg[e][lane + 32 i]), the token row is read once; expert counts are aggregated in shared memory so that only E
// global atomics are issued per 64 tokens.
#define MOE_TOK_PER_BLOCK 64
template <typename T, int NV, int EMAX>   // NV = D / 32 values per lane, EMAX >= E
__global__ void __launch_bounds__(256)
moe_route_kernel(const T* __restrict__ x, const float* __restrict__ Wg, int N, int D, int E, int topk,
                 int* __restrict__ top_idx, float* __restrict__ top_w, int* __restrict__ counts) {
  __shared__ int hist[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 32) hist[threadIdx.x] = 0;
  float wg[EMAX][NV];
#pragma unroll
  for (int e = 0; e < EMAX; ++e)
#pragma unroll
    for (int i = 0; i < NV; ++i) wg[e][i] = e < E ? Wg[e * D + i * 32 + lane] : 0.f;
  __syncthreads();
  for (int it = 0; it < MOE_TOK_PER_BLOCK / 8; ++it) {
    const int tok = blockIdx.x * MOE_TOK_PER_BLOCK + it * 8 + warp;
    if (tok >= N) break;
    const T* xr = x + (int64_t)tok * D;
    float xv[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) xv[i] = moe_to_f(xr[i * 32 + lane]);
    float my_logit = -INFINITY;  // lane e holds logit e
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) acc = fmaf(xv[i], wg[e][i], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == e && e < E) my_logit = acc;
    }
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
        atomicAdd(&hist[sel_i[k]], 1);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < E && hist[threadIdx.x]) atomicAdd(counts + threadIdx.x, hist[threadIdx.x]);
}

template <typename T>
static int launch_route(const T* x, const float* Wg, int N, int D, int E, int topk, int* top_idx, float* top_w, int* counts,
                        cudaStream_t s) {
  const unsigned gt = (unsigned)ymt3_div_up(N, MOE_TOK_PER_BLOCK);
#define ROUTE(NV, EM) moe_route_kernel<T, NV, EM><<<gt, 256, 0, s>>>(x, Wg, N, D, E, topk, top_idx, top_w, counts)
  if (E <= 8) {
    if (D == 128) ROUTE(4, 8); else if (D == 256) ROUTE(8, 8); else if (D == 512) ROUTE(16, 8); else if (D == 64) ROUTE(2, 8);
    else { ymt3_set_error("moe: d_model must be 64/128/256/512 (got %d)", D); return YMT3_ERR_UNSUPPORTED; }
  } else if (E <= 16) {
    if (D == 128) ROUTE(4, 16); else if (D == 256) ROUTE(8, 16); else if (D == 512) ROUTE(16, 16); else if (D == 64) ROUTE(2, 16);
    else { ymt3_set_error("moe: d_model must be 64/128/256/512 (got %d)", D); return YMT3_ERR_UNSUPPORTED; }
  } else {
    ymt3_set_error("moe: at most 16 experts are supported by the fused router (got %d)", E);
    return YMT3_ERR_UNSUPPORTED;
  }
#undef ROUTE
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
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

// Block = 64 tokens (<= 256 (token, k) items): local ranks via shared-memory atomics, ONE global atomic per
// expert per block reserves the block's range in the expert segment, then warps copy the rows.
template <typename T>
__global__ void __launch_bounds__(256)
moe_scatter_kernel(const T* __restrict__ x, int N, int D, int topk, const int* __restrict__ top_idx,
                   const float* __restrict__ top_w, const int* __restrict__ offsets, int* __restrict__ cursor,
                   T* __restrict__ xs, float* __restrict__ slot_w, int* __restrict__ tok_slot) {
  __shared__ int hist[32], base[32];
  __shared__ int slots[MOE_TOK_PER_BLOCK * 4];
  const int tid = threadIdx.x;
  const int item0 = blockIdx.x * MOE_TOK_PER_BLOCK * topk;
  const int n_items = min(MOE_TOK_PER_BLOCK * topk, N * topk - item0);
  if (tid < 32) hist[tid] = 0;
  __syncthreads();
  int e = 0, rank = 0;
  if (tid < n_items) {
    e = top_idx[item0 + tid];
    rank = atomicAdd(&hist[e], 1);
  }
  __syncthreads();
  if (tid < 32 && hist[tid]) base[tid] = atomicAdd(cursor + tid, hist[tid]);
  __syncthreads();
  if (tid < n_items) {
    const int slot = offsets[e] + base[e] + rank;
    slots[tid] = slot;
    slot_w[slot] = top_w[item0 + tid];
    tok_slot[item0 + tid] = slot;
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  for (int i = warp; i < n_items; i += 8) {
    const int tok = (item0 + i) / topk;
    const T* src = x + (int64_t)tok * D;
    T* dst = xs + (int64_t)slots[i] * D;
    for (int d = lane; d < D; d += 32) dst[d] = src[d];
  }
}

// out[tok, :] = residual[tok, :] + sum_k ys[tok_slot[tok, k], :]     (4 fp32 / 8 bf16 elements per thread)
template <typename T>
__global__ void __launch_bounds__(256)
moe_combine_kernel(const T* __restrict__ ys, const int* __restrict__ tok_slot, const T* __restrict__ residual,
                   T* __restrict__ out, int N, int D, int topk) {
  constexpr int V = 16 / sizeof(T);
  const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * V;
  if (i >= (int64_t)N * D) return;
  const int tok = (int)(i / D), d = (int)(i % D);
  float acc[V];
  auto add = [&](const T* p, bool init) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    if constexpr (sizeof(T) == 4) {
      const float* f = reinterpret_cast<const float*>(&u);
#pragma unroll
      for (int q = 0; q < V; ++q) acc[q] = init ? f[q] : acc[q] + f[q];
    } else {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int q = 0; q < V / 2; ++q) {
        acc[2 * q] = (init ? 0.f : acc[2 * q]) + __bfloat162float(h[q].x);
        acc[2 * q + 1] = (init ? 0.f : acc[2 * q + 1]) + __bfloat162float(h[q].y);
      }
    }
  };
  if (residual) add(residual + i, true);
  else {
#pragma unroll
    for (int q = 0; q < V; ++q) acc[q] = 0.f;
  }
  for (int k = 0; k < topk; ++k) add(ys + (int64_t)tok_slot[tok * topk + k] * D + d, false);
  uint4 o;
  if constexpr (sizeof(T) == 4) {
    float* f = reinterpret_cast<float*>(&o);
#pragma unroll
    for (int q = 0; q < V; ++q) f[q] = acc[q];
  } else {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int q = 0; q < V / 2; ++q) h[q] = __floats2bfloat162_rn(acc[2 * q], acc[2 * q + 1]);
  }
  *reinterpret_cast<uint4*>(out + i) = o;
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
  YMT3_REQUIRE(w.E >= 1 && w.E <= 16 && w.topk >= 1 && w.topk <= 4 && w.topk <= w.E, "moe: bad E/topk");
  YMT3_REQUIRE(w.D <= 512, "moe: router supports d_model <= 512 (got %d)", w.D);
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
  const unsigned gt = (unsigned)ymt3_div_up(N, MOE_TOK_PER_BLOCK), gs = gt;
  if (precision == YMT3_F32) {
    if (int rr = launch_route<float>((const float*)x, w.gate, N, D, E, topk, top_idx, top_w, counts, s)) return rr;
    moe_offsets_kernel<<<1, 32, 0, s>>>(counts, offsets, cursor, E);
    moe_scatter_kernel<float><<<gs, 256, 0, s>>>((const float*)x, N, D, topk, top_idx, top_w, offsets, cursor,
                                                 (float*)xs, slot_w, tok_slot);
  } else {
    if (int rr = launch_route<__nv_bfloat16>((const __nv_bfloat16*)x, w.gate, N, D, E, topk, top_idx, top_w, counts, s)) return rr;
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
  YMT3_REQUIRE(D % 8 == 0, "moe: d_model must be a multiple of 8");
  const int64_t vec = 16 / (int64_t)es;
  const unsigned gc = (unsigned)(((int64_t)N * D / vec + 255) / 256);
  if (precision == YMT3_F32)
    moe_combine_kernel<float><<<gc, 256, 0, s>>>((const float*)ys, tok_slot, (const float*)residual, (float*)out, N, D, topk);
  else
    moe_combine_kernel<__nv_bfloat16><<<gc, 256, 0, s>>>((const __nv_bfloat16*)ys, tok_slot, (const __nv_bfloat16*)residual,
                                                         (__nv_bfloat16*)out, N, D, topk);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

}  // namespace ymt3
