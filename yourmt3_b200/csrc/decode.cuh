// Decode-step op launchers (decode.cu).
#pragma once
#include "common.cuh"

namespace ymt3 {

// ss_out (optional, (N, dim/32) fp32): per-32-column sums of squares of the stored row (fused RMSNorm producer)
int embed_pos(const int* tok, const void* E, const void* pos, const int* step, void* x, int N, int dim,
              int dtype, cudaStream_t stream, float* ss_out = nullptr);

// q: (N, H*dk) rows with leading dim q_ld. knew/vnew (self mode) rows with leading dim new_ld or
// null (cross mode). Kc/Vc: caches addressed as base + n*c_sn + h*c_sh + j*c_ss (+ d), element strides.
// step: device int (self mode attends keys [0, *step]; cross mode keys [0, fixed_len)). out: (N, H*dk).
// rel_bias (self mode, optional): T5 relative position bias per distance, (H, rel_stride) fp32: score of key j +=
// rel_bias[h][*step - j] (HF modeling_t5.py:248-268, unidirectional buckets).
int decode_attention(const void* q, int64_t q_ld, const void* knew, const void* vnew, int64_t new_ld, void* Kc,
                     void* Vc, int64_t c_sn, int64_t c_sh, int64_t c_ss, int Lmax, const int* step, int fixed_len,
                     float scale, void* out, int64_t out_ld, int N, int H, int dk, int dtype, cudaStream_t stream,
                     const float* rel_bias = nullptr, int rel_stride = 0);

// The vocab-projection GEMM leaves one packed arg-max key per row (GemmParams::argmax_out, ops.cuh); this single-block
// launch turns the keys into tokens (finished mask, EOS, forced (N, n_forced) task prefix or null), zeroes them and
// advances *step.  score_out (optional, (N, n_forced)): the arg-max of every FORCED step (teacher-forced scoring).
int select_advance(unsigned long long* keys, int N, int* step, int* cur_tok, int* finished, int* tokens_out,
                   int max_len, int eos_id, int pad_id, int stop_at_eos, int* unfinished_count, const int* forced,
                   int n_forced, cudaStream_t stream, int* score_out = nullptr);
// Absorbed cross-attention (cross_absorbed.cu): q (N, H*zdim) bf16 latent-space queries, z (N, Tp, zdim) bf16 latents
// (rows >= T zero), out (N, H*zdim) bf16 = softmax_t(q_h . z_t) z_t per head.  zdim must be 256, H <= 8, Tp % 16 == 0.
int cross_attn_absorbed(const void* q, int64_t q_ld, const void* z, void* out, int64_t out_ld, int64_t N, int H, int T,
                        int Tp, int zdim, cudaStream_t stream);
// (B, T, C, zdim) bf16 -> (B*C, Tp, zdim); destination rows >= T are left untouched
int gather_latents(const void* src, void* dst, int64_t B, int T, int C, int Tp, int zdim, cudaStream_t stream);

int fill_i32(int* p, int v, int64_t n, cudaStream_t stream);

}  // namespace ymt3
