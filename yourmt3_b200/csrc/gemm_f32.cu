// SIMT fp32 GEMM with fused epilogue (bias / activation / gating / scale / residual).
// This is the exact-arithmetic (fp32 FFMA, fp32 accumulate, fixed k order) path used when
// the model handle is created with precision f32 -- the path on which decoded tokens must
// equal the reference's. The throughput path is gemm_bf16_tc.cu (tcgen05).
#include "ops.cuh"

namespace ymt3 {

template <int BM, int BN>
__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmParams p) {
  constexpr int BK = 16;
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
  constexpr int A_LD = BM / 64, B_LD = BN / 64;  // float4 loads per thread per tile
  __shared__ __align__(16) float As[2][BK][LDA_S];
  __shared__ __align__(16) float Bs[2][BK][LDB_S];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const float* __restrict__ A = static_cast<const float*>(p.A);
  const float* __restrict__ W = static_cast<const float*>(p.W);
  const float* __restrict__ bias = p.bias;
  int row_end = p.M;
  int m0 = blockIdx.x * BM;   // M tiles on grid.x (2^31 limit), N tiles on grid.y
  if (p.group_offsets) {
    const int g = blockIdx.z;
    const int row_begin = p.group_offsets[g];
    row_end = p.group_offsets[g + 1];
    m0 += row_begin;
    W += (int64_t)g * p.strideW;
    if (bias) bias += (int64_t)g * p.N;
  }
  if (m0 >= row_end) return;
  const int n0 = blockIdx.y * BN;
  const int K = p.K, N = p.N;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[A_LD], rb[B_LD];
  auto load_global = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      int idx = tid + 256 * i;
      int r = m0 + (idx >> 2), k = k0 + (idx & 3) * 4;
      ra[i] = (r < row_end && k < K) ? __ldg(reinterpret_cast<const float4*>(A + (int64_t)r * p.lda + k))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      int idx = tid + 256 * i;
      int r = n0 + (idx >> 2), k = k0 + (idx & 3) * 4;
      rb[i] = (r < N && k < K) ? __ldg(reinterpret_cast<const float4*>(W + (int64_t)r * p.ldw + k))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      int idx = tid + 256 * i;
      int r = idx >> 2, k = (idx & 3) * 4;
      As[buf][k + 0][r] = ra[i].x;
      As[buf][k + 1][r] = ra[i].y;
      As[buf][k + 2][r] = ra[i].z;
      As[buf][k + 3][r] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      int idx = tid + 256 * i;
      int r = idx >> 2, k = (idx & 3) * 4;
      Bs[buf][k + 0][r] = rb[i].x;
      Bs[buf][k + 1][r] = rb[i].y;
      Bs[buf][k + 2][r] = rb[i].z;
      Bs[buf][k + 3][r] = rb[i].w;
    }
  };

  const int nk = (K + BK - 1) / BK;
  load_global(0);
  store_smem(0);
  __syncthreads();
  int buf = 0;
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) load_global((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int bi = 0; bi < TM / 4; ++bi) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][k][bi * 64 + ty * 4]);
        a[bi * 4 + 0] = v.x; a[bi * 4 + 1] = v.y; a[bi * 4 + 2] = v.z; a[bi * 4 + 3] = v.w;
      }
#pragma unroll
      for (int bj = 0; bj < TN / 4; ++bj) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][bj * 64 + tx * 4]);
        b[bj * 4 + 0] = v.x; b[bj * 4 + 1] = v.y; b[bj * 4 + 2] = v.z; b[bj * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_smem(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  // ---- epilogue ----
  float* C = static_cast<float*>(p.C);  // may alias residual (in-place x += ...)
  const float* R = static_cast<const float*>(p.residual);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
    if (r >= row_end) continue;
    const float rs = p.out_scale * (p.row_scale ? p.row_scale[r] : 1.0f);
#pragma unroll
    for (int bj = 0; bj < TN / 4; ++bj) {
      const int c = n0 + bj * 64 + tx * 4;
      if (c >= N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = acc[i][bj * 4 + j] + (bias ? bias[c + j] : 0.f);
      if (p.gated) {
        float o0 = act_apply(v[0], p.act) * v[1] * rs;
        float o1 = act_apply(v[2], p.act) * v[3] * rs;
        const int co = c >> 1;
        if (R) {
          o0 += R[(int64_t)r * p.ldr + co];
          o1 += R[(int64_t)r * p.ldr + co + 1];
        }
        *reinterpret_cast<float2*>(C + (int64_t)r * p.ldc + co) = make_float2(o0, o1);
      } else {
        float4 o;
        o.x = act_apply(v[0], p.act) * rs;
        o.y = act_apply(v[1], p.act) * rs;
        o.z = act_apply(v[2], p.act) * rs;
        o.w = act_apply(v[3], p.act) * rs;
        if (p.argmax_out) {   // fused greedy selection (ops.cuh); no residual in this mode
          const float ov[4] = {o.x, o.y, o.z, o.w};
          unsigned long long best_key = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c + j < p.argmax_n) {
              const unsigned long long key = argmax_key(ov[j], c + j);
              best_key = key > best_key ? key : best_key;
            }
          if (best_key) atomicMax(p.argmax_out + r, best_key);
        }
        if (R) {
          const float4 q = *reinterpret_cast<const float4*>(R + (int64_t)r * p.ldr + c);
          o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
        }
        *reinterpret_cast<float4*>(C + (int64_t)r * p.ldc + c) = o;
      }
    }
  }
}

int gemm_f32(const GemmParams& p, cudaStream_t stream) {
  YMT3_REQUIRE(!p.norm_ss_in && !p.ss_out, "gemm_f32: fused RMSNorm is a bf16 tensor-core path feature");
  YMT3_REQUIRE(p.A && p.W && p.C, "gemm_f32: null pointer");
  YMT3_REQUIRE(!p.argmax_out || (!p.gated && !p.residual && !p.group_offsets && p.argmax_n > 0 && p.argmax_n <= p.N),
               "gemm_f32: fused arg-max needs no gate / residual / groups, 0 < argmax_n <= N");
  if (p.M <= 0 || p.N <= 0) return YMT3_OK;
  YMT3_REQUIRE(p.K > 0 && p.K % 4 == 0 && p.lda % 4 == 0 && p.ldw % 4 == 0,
               "gemm_f32: K, lda, ldw must be multiples of 4 (K=%d lda=%lld ldw=%lld)", p.K,
               (long long)p.lda, (long long)p.ldw);
  YMT3_REQUIRE(p.N % 4 == 0, "gemm_f32: N must be a multiple of 4 (N=%d)", p.N);
  YMT3_REQUIRE(p.gated ? (p.ldc % 2 == 0) : (p.ldc % 4 == 0), "gemm_f32: ldc alignment (ldc=%lld)",
               (long long)p.ldc);
  YMT3_REQUIRE(!p.residual || (p.gated ? p.ldr % 2 == 0 : p.ldr % 4 == 0), "gemm_f32: ldr alignment");
  YMT3_REQUIRE((((uintptr_t)p.A | (uintptr_t)p.W | (uintptr_t)p.C | (uintptr_t)p.residual) & 15) == 0,
               "gemm_f32: pointers must be 16-byte aligned");
  const int groups = p.group_offsets ? p.num_groups : 1;
  const int64_t big_tiles = (int64_t)ymt3_div_up(p.M, 128) * ymt3_div_up(p.N, 128);
  if (big_tiles >= 2 * (int64_t)ymt3_num_sms() && !p.group_offsets) {
    dim3 grid(ymt3_div_up(p.M, 128), ymt3_div_up(p.N, 128), 1);
    gemm_f32_kernel<128, 128><<<grid, 256, 0, stream>>>(p);
  } else {
    dim3 grid(ymt3_div_up(p.M, 64), ymt3_div_up(p.N, 64), groups);
    gemm_f32_kernel<64, 64><<<grid, 256, 0, stream>>>(p);
  }
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

}  // namespace ymt3
