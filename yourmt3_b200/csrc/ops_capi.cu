// Per-op C-ABI entry points (include/ymt3_b200.h, "Per-op entry points").
#include "ops.cuh"
#include "decode.cuh"
#include "moe.cuh"
#include "../../include/ymt3_b200.h"

using namespace ymt3;

extern "C" int ymt3_op_linear(int32_t dtype, const void* A, int64_t lda, const void* W, int64_t ldw,
                              const float* bias, void* C, int64_t ldc, const void* residual, int64_t ldr, int64_t M,
                              int64_t N, int64_t K, int32_t act, int32_t gated, float out_scale, int32_t out_dtype,
                              void* stream) {
  YMT3_REQUIRE(M >= 0 && M < (1ll << 31) && N > 0 && N < (1ll << 31) && K > 0 && K < (1ll << 31), "op_linear: bad shape");
  GemmParams p{};
  p.A = A; p.lda = lda; p.W = W; p.ldw = ldw; p.C = C; p.ldc = ldc; p.bias = bias;
  p.residual = residual; p.ldr = ldr; p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.act = act; p.gated = gated; p.out_scale = out_scale;
  if (dtype == YMT3_F32) {
    YMT3_REQUIRE(out_dtype == YMT3_F32, "op_linear: f32 path writes f32");
    return gemm_f32(p, (cudaStream_t)stream);
  }
  return gemm_bf16_tc(p, out_dtype, (cudaStream_t)stream);
}

extern "C" int ymt3_op_rmsnorm(int32_t dtype, const void* x, const float* w, void* y, int64_t rows, int64_t dim,
                               float eps, void* stream) {
  return rmsnorm(x, w, y, rows, (int)dim, eps, dtype, (cudaStream_t)stream);
}

extern "C" int ymt3_op_layernorm(int32_t dtype, const void* x, const float* w, const float* b, void* y, int64_t rows,
                                 int64_t dim, float eps, void* stream) {
  return layernorm(x, w, b, y, rows, (int)dim, eps, dtype, (cudaStream_t)stream);
}

extern "C" int ymt3_op_attention(int32_t dtype, const void* q, const void* k, const void* v, void* o, int64_t B,
                                 int64_t H, int64_t Sq, int64_t Sk, int64_t dk, float scale, int32_t causal,
                                 void* stream) {
  AttnParams a{};
  a.Q = q; a.K = k; a.V = v; a.O = o;
  a.q_sb = Sq * H * dk; a.q_sh = dk; a.q_ss = H * dk;
  a.k_sb = Sk * H * dk; a.k_sh = dk; a.k_ss = H * dk;
  a.v_sb = Sk * H * dk; a.v_sh = dk; a.v_ss = H * dk;
  a.o_sb = Sq * H * dk; a.o_sh = dk; a.o_ss = H * dk;
  a.B = (int)B; a.H = (int)H; a.Sq = (int)Sq; a.Sk = (int)Sk; a.dk = (int)dk;
  a.scale = scale; a.causal = causal;
  return attention(a, dtype, (cudaStream_t)stream);
}

extern "C" int ymt3_op_cross_attn_absorbed(const void* q, const void* z, void* out, int64_t N, int64_t H, int64_t T,
                                           int64_t Tp, void* stream) {
  YMT3_REQUIRE(q && z && out, "op_cross_attn_absorbed: null argument");
  return cross_attn_absorbed(q, H * 256, z, out, H * 256, N, (int)H, (int)T, (int)Tp, 256, (cudaStream_t)stream);
}

extern "C" int ymt3_op_decode_attention(int32_t dtype, const void* q, const void* knew, const void* vnew, void* Kc,
                                        void* Vc, const int32_t* step_dev, int64_t fixed_len, void* out, int64_t N,
                                        int64_t H, int64_t Lcap, void* stream) {
  YMT3_REQUIRE(q && Kc && Vc && out && (knew == nullptr) == (vnew == nullptr), "op_decode_attention: null argument");
  YMT3_REQUIRE(knew ? step_dev != nullptr : (fixed_len > 0 && fixed_len <= Lcap), "op_decode_attention: bad length");
  return decode_attention(q, H * 64, knew, vnew, H * 64, Kc, Vc, H * Lcap * 64, Lcap * 64, 64, (int)Lcap, step_dev,
                          (int)fixed_len, 1.0f, out, H * 64, (int)N, (int)H, 64, dtype, (cudaStream_t)stream);
}

extern "C" int ymt3_op_linear_normfused(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                        const float* ss_in, int64_t chunks, float eps, void* C, int64_t ldc,
                                        const void* residual, int64_t ldr, float* ss_out, int64_t M, int64_t N, int64_t K,
                                        int32_t act, int32_t gated, float out_scale, int32_t out_dtype, void* stream) {
  YMT3_REQUIRE(M >= 0 && M < (1ll << 31) && N > 0 && N < (1ll << 31) && K > 0 && K < (1ll << 31),
               "op_linear_normfused: bad shape");
  GemmParams p{};
  p.A = A; p.lda = lda; p.W = W; p.ldw = ldw; p.C = C; p.ldc = ldc; p.bias = bias;
  p.residual = residual; p.ldr = ldr; p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.act = act; p.gated = gated; p.out_scale = out_scale;
  p.norm_ss_in = ss_in; p.norm_ss_chunks = (int)chunks; p.norm_eps = eps; p.ss_out = ss_out;
  return gemm_bf16_tc(p, out_dtype, (cudaStream_t)stream);
}

extern "C" int ymt3_debug_chain_trace(uint64_t* device_buf) {
  gemm_chain_set_trace(reinterpret_cast<unsigned long long*>(device_buf));
  return YMT3_OK;
}

extern "C" int64_t ymt3_op_linear_chain_counters(int64_t M) {
  return M > 0 && M < (1ll << 31) ? gemm_chain_counters((int)M) : 0;
}

extern "C" int ymt3_op_linear_chain(const ymt3_chain_phase_t* phases, int32_t n_phases, int64_t M, int32_t* counters,
                                    int32_t ordinal, void* stream) {
  YMT3_REQUIRE(phases && n_phases >= 1 && n_phases <= 4 && counters && ordinal >= 0, "op_linear_chain: bad arguments");
  YMT3_REQUIRE(M >= 0 && M < (1ll << 31), "op_linear_chain: bad shape");
  GemmParams g[4];
  for (int i = 0; i < n_phases; ++i) {
    const ymt3_chain_phase_t& q = phases[i];
    YMT3_REQUIRE(q.N > 0 && q.N < (1ll << 31) && q.K > 0 && q.K < (1ll << 31), "op_linear_chain: phase %d: bad shape", i);
    GemmParams p{};
    p.A = q.A; p.lda = q.lda; p.W = q.W; p.ldw = q.ldw; p.C = q.C; p.ldc = q.ldc; p.bias = q.bias;
    p.residual = q.residual; p.ldr = q.ldr; p.M = (int)M; p.N = (int)q.N; p.K = (int)q.K;
    p.act = q.act; p.gated = q.gated; p.out_scale = q.out_scale;
    p.norm_ss_in = q.ss_in; p.norm_ss_chunks = (int)q.chunks; p.norm_eps = q.eps; p.ss_out = q.ss_out;
    g[i] = p;
  }
  return gemm_chain_bf16(g, n_phases, counters, nullptr, 0, ordinal, (cudaStream_t)stream);
}

extern "C" int ymt3_op_linear_argmax(int32_t dtype, const void* A, int64_t lda, const void* W, int64_t ldw,
                                     const float* bias, float* logits, int64_t ldc, int64_t M, int64_t N, int64_t K,
                                     int64_t V, float out_scale, uint64_t* keys, void* stream) {
  YMT3_REQUIRE(M >= 0 && M < (1ll << 31) && N > 0 && N < (1ll << 31) && K > 0 && K < (1ll << 31) && V > 0 && V <= N,
               "op_linear_argmax: bad shape");
  YMT3_REQUIRE(keys, "op_linear_argmax: null keys");
  GemmParams p{};
  p.A = A; p.lda = lda; p.W = W; p.ldw = ldw; p.C = logits; p.ldc = ldc; p.bias = bias;
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.out_scale = out_scale;
  p.argmax_out = reinterpret_cast<unsigned long long*>(keys); p.argmax_n = (int)V;
  if (dtype == YMT3_F32) return gemm_f32(p, (cudaStream_t)stream);
  return gemm_bf16_tc(p, YMT3_F32, (cudaStream_t)stream);
}

extern "C" int64_t ymt3_op_moe_workspace_bytes(int64_t N, int32_t D, int32_t I, int32_t E, int32_t topk, int32_t dtype) {
  if (N <= 0 || D <= 0 || I <= 0 || E <= 0 || topk <= 0) return 0;
  return (int64_t)moe_workspace_bytes(N, D, I, E, topk, dtype);
}

extern "C" int ymt3_op_moe_ff(int32_t dtype, const void* x, const void* residual, void* out, int64_t N,
                              const float* gate, const void* w13, const void* w2, int32_t D, int32_t I, int32_t E,
                              int32_t topk, int32_t act, void* workspace, void* stream) {
  YMT3_REQUIRE(dtype == YMT3_F32 || dtype == YMT3_BF16, "op_moe_ff: bad dtype %d", dtype);
  YMT3_REQUIRE(N >= 0 && D > 0 && I > 0, "op_moe_ff: bad shape");
  if (N == 0) return YMT3_OK;
  YMT3_REQUIRE(x && out && gate && w13 && w2 && workspace, "op_moe_ff: null argument");
  MoEWeights w;
  w.gate = gate; w.w13 = w13; w.w2 = w2; w.D = D; w.I = I; w.E = E; w.topk = topk; w.act = act;
  return moe_forward(dtype, x, residual, out, N, w, workspace, (cudaStream_t)stream);
}
