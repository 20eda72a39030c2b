// Internal op launchers shared by the model orchestration (model.cu) and the
// per-op C-ABI test entry points (ops_capi.cu). All launch on `stream`, never
// synchronise, never allocate. dtype: 0 = f32, 1 = bf16 (activations/weights);
// accumulation, softmax and normalisation statistics are always fp32.
#pragma once
#include "common.cuh"
#include "index_maps.h"   // argmax_key, TMA staging maps, tile-width heuristic (shared with tests/host_emu)

#define YMT3_F32 0
#define YMT3_BF16 1

#define YMT3_ACT_NONE 0
#define YMT3_ACT_GELU_NEW 1   // tanh approximation (HF activations.py:59-66), T5 "gated-gelu"
#define YMT3_ACT_RELU 2
#define YMT3_ACT_SILU 3
#define YMT3_ACT_GELU 4       // erf GELU (torch.nn.functional.gelu default)

namespace ymt3 {

__device__ __forceinline__ float act_apply(float x, int act) {
  switch (act) {
    case YMT3_ACT_GELU_NEW: {
      const float k = 0.7978845608028654f;  // sqrt(2/pi)
      return 0.5f * x * (1.0f + tanhf(k * (x + 0.044715f * x * x * x)));
    }
    case YMT3_ACT_RELU: return fmaxf(x, 0.f);
    case YMT3_ACT_SILU: return x / (1.0f + expf(-x));
    case YMT3_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    default: return x;
  }
}

// C = residual + out_scale * epi(A @ W^T + bias)
//   A: (M, K) row-major, lda.  W: (N, K) row-major (nn.Linear.weight), ldw.
//   gated == 0: C is (M, N);           epi(x) = act(x)
//   gated == 1: C is (M, N/2); rows of W are interleaved (2j = gate/act branch,
//               2j+1 = linear branch): C[:, j] = act(x[:, 2j]) * x[:, 2j+1]
//   bias: fp32 (N) or null. residual: same dtype/ld as C or null (may alias C).
//   Grouped mode (group_offsets != null): rows [off[g], off[g+1]) of A/C use weight
//   W + g*strideW (and bias + g*N); num_groups groups; M = total rows upper bound.
struct GemmParams {
  const void* A; int64_t lda;
  const void* W; int64_t ldw;
  void* C; int64_t ldc;
  const float* bias;
  const void* residual; int64_t ldr;
  int M, N, K;
  int act, gated;
  float out_scale;
  const int* group_offsets; int num_groups; int64_t strideW;
  // optional per-row scale applied with out_scale (MoE combine weights), fp32 (M) or null
  const float* row_scale;
  // Fused RMSNorm (bf16 tensor-core path only; decode step).  rmsnorm(x) W^T = diag(r) x (W . w_ln)^T with
  // r = rsqrt(mean(x^2) + eps): the consumer GEMM reads x itself, has w_ln folded into its weight columns at pack
  // time and multiplies every accumulator row by r BEFORE bias / activation; r comes from per-row partial sums of
  // squares (norm_ss_in: (M, norm_ss_chunks) fp32, one partial per 32 columns of x, summed in a fixed order ->
  // deterministic).  The producer GEMM (the one that writes x = residual + ...) emits those partials from the
  // bf16-ROUNDED values it stores (ss_out: (M, ceil(N/32)) fp32), i.e. exactly the numbers a norm kernel would read.
  const float* norm_ss_in; int norm_ss_chunks; float norm_eps;
  float* ss_out;
  // Fused greedy selection (vocab projection of the decode step; fp32 output, no residual, not gated): every
  // epilogue thread folds the columns < argmax_n it produced into one packed key (argmax_key, index_maps.h) and issues one
  // atomicMax per row and tile into argmax_out (M keys, zero before the launch).  The maximum key is the largest
  // logit with the SMALLEST column index among equals = torch.argmax's first-maximum rule.
  unsigned long long* argmax_out; int argmax_n;
};

int gemm_f32(const GemmParams& p, cudaStream_t stream);            // SIMT fp32 FFMA
int gemm_bf16_tc(const GemmParams& p, int out_dtype, cudaStream_t stream);  // tcgen05/TMEM/TMA
// Several dependent bf16 GEMMs in ONE persistent tcgen05 launch (gemm_bf16_tc.cu, gemm_chain_kernel): phase i > 0 reads
// what phase i - 1 wrote, dependencies tracked per 128-row tile through `done` (gemm_chain_counters(M) device ints,
// zeroed by the caller at the start of a run); launch ordinal = *epoch_ptr * epoch_mul + epoch_add (0, 1, 2, ... over
// the launches that share `done`).  Same arithmetic, tile by tile, as the separate launches.
int gemm_chain_bf16(const GemmParams* phases, int n_phases, int* done, const int* epoch_ptr, int epoch_mul, int epoch_add,
                    cudaStream_t stream);
int gemm_chain_counters(int M);
// Fused MoE expert MLP (gemm_bf16_tc.cu, moe_expert_fused_kernel): both grouped GEMMs of the experts in one kernel, the
// hidden tile stays in shared memory.  d_model 128, hidden 512, SiLU / gelu_new, bf16 only.
int moe_expert_fused(const void* xs, int64_t S, const void* w13, const void* w2, const int* offsets, int E,
                     const float* slot_w, void* ys, int act, cudaStream_t stream);
void gemm_chain_set_trace(unsigned long long* buf);   // debug: per-CTA event timestamps, (SMs x 16 x 32) words or null
// 3x3 same-padding convolution on NHWC bf16 input as an implicit GEMM (tcgen05; TMA 4-D tiles with
// zero-filled halo); `epi` carries W (Cout, 9*Cin), N=Cout, C/ldc, bias, act, residual/ldr, out_scale.
int conv3x3_bf16_tc(const void* x, int B, int T, int F, int Cin, const GemmParams& epi, int out_dtype,
                    cudaStream_t stream);

// y = x * rsqrt(mean(x^2) + eps) * w          (T5LayerNorm / RMSNorm)
// y = (x - mean) * rsqrt(var + eps) * w + b   (nn.LayerNorm), w/b fp32
int rmsnorm(const void* x, const float* w, void* y, int64_t rows, int dim, float eps, int dtype,
            cudaStream_t stream);
int layernorm(const void* x, const float* w, const float* b, void* y, int64_t rows, int dim,
              float eps, int dtype, cudaStream_t stream);
// mixed precision RMSNorm: x fp32 (the residual stream of the bf16 T5 encoder), y bf16 (tensor-core operand)
int rmsnorm_f32_bf16(const float* x, const float* w, void* y, int64_t rows, int dim, float eps, cudaStream_t stream);

// Generic small-sequence attention: O[b,h,i,:] = softmax_j(scale * Q[b,h,i,:].K[b,h,j,:] (+causal)) V[b,h,j,:]
// element strides are given in ELEMENTS; head dim dk in {16, 32, 64, 128}.
struct AttnParams {
  const void* Q; int64_t q_sb, q_sh, q_ss;   // batch, head, seq strides
  const void* K; int64_t k_sb, k_sh, k_ss;
  const void* V; int64_t v_sb, v_sh, v_ss;
  void* O; int64_t o_sb, o_sh, o_ss;
  int B, H, Sq, Sk, dk;
  // optional two-level batch: batch index b -> (b / inner, b % inner), offset = (b / inner) * sb + (b % inner) * sb2
  // (temporal attention over a (B, T, K, D) layout without transposing); inner <= 1 disables it
  int inner;
  int64_t q_sb2, k_sb2, v_sb2, o_sb2;
  float scale;
  // optional rotate-half RoPE applied to q and k on load (tiny-sequence kernel only): rot dims per head,
  // position = index along the attention sequence; tables (n_pos, rot/2) fp32
  const float* rope_cos; const float* rope_sin; int rope_dim;
  int causal;        // key j allowed iff j <= i + (Sk - Sq)
  const int* kv_len; // optional per-batch valid key count (device), else Sk
  // optional T5 relative attention bias folded per distance (HF modeling_t5.py:248-268 compute_bias): score(i, j) +=
  // rel_bias[h * rel_stride + rel_center + (j - i)]; fp32 table; generic kernel only
  const float* rel_bias; int rel_stride, rel_center;
};
int attention(const AttnParams& p, int dtype, cudaStream_t stream);

// elementwise helpers
int add_rows(const void* x, const void* table, void* y, int64_t rows, int period, int dim, int dtype,
             cudaStream_t stream, int64_t div = 1);  // y[r,:] = x[r,:] + table[(r / div) % period,:]
int split_kv_heads(const void* kv, void* Kout, void* Vout, int64_t N, int Tn, int H, int dk, int dtype,
                   cudaStream_t stream);
int permute_btcd_bctd(const void* x, void* y, int64_t B, int64_t T, int64_t C, int64_t D, int dtype,
                      cudaStream_t stream);
int convert(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, cudaStream_t stream);
// y[r, :] = table[r % period, :] (+ table2[r % period, :])   (latent array broadcast)
int tile_rows(const void* table, const void* table2, void* y, int64_t rows, int period, int dim, int dtype,
              cudaStream_t stream);
// in-place rotate-half RoPE on `heads` heads of width dh starting at column col0 of a (rows, ld) matrix;
// position of row r = (r / pos_div) % pos_mod; cos/sin tables are (n_pos, rot/2) fp32
int rope_inplace(void* x, int64_t rows, int64_t ld, int col0, int heads, int dh, int rot, int64_t pos_div, int pos_mod,
                 const float* cos_t, const float* sin_t, int dtype, cudaStream_t stream);

}  // namespace ymt3
