/*
 * ymt3_b200.h -- C ABI of the B200-native YourMT3 inference hot path.
 *
 * Every entry point is plain C: pointers + sizes, an int status (0 = ok) and a
 * thread-local error string (ymt3_last_error).  No torch / C++ types cross the
 * boundary.  All *_dev pointers are CUDA device pointers owned by the caller;
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 * No entry point synchronises the device unless its name ends in `_host`.
 *
 * The reference (richhiey/YourMT3 -> upstream mimbres/YourMT3, amt/src/model/)
 * is pure Python and has no FFI; each function below names the reference
 * Python interface it replaces.  The mounted fork holds only README.md, so
 * upstream paths are given by name ([RECALL]) and the
 * arithmetic is cited into the installed dependencies the upstream path calls
 * (torchaudio 2.11 / transformers 5.5, abbreviated TA/ and HF/).
 */
#ifndef YMT3_B200_H
#define YMT3_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YMT3_ABI_VERSION 1

#if defined(__GNUC__)
#define YMT3_API __attribute__((visibility("default")))
#else
#define YMT3_API
#endif

/* status codes */
#define YMT3_STATUS_OK 0
#define YMT3_STATUS_INVALID 1
#define YMT3_STATUS_CUDA 2
#define YMT3_STATUS_UNSUPPORTED 3

YMT3_API const char* ymt3_last_error(void);
YMT3_API int ymt3_abi_version(void);
/* fills name (<= cap bytes), sm count, compute capability major/minor of the current device */
YMT3_API int ymt3_device_info(char* name, int cap, int* num_sms, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------- *
 * Frontend: waveform -> log-(mel)spectrogram.
 * Replaces upstream model/spectrogram.py: get_spectrogram_layer_from_audio_cfg(),
 * Melspectrogram.forward / Spectrogram.forward, i.e.
 *   TA/transforms/_transforms.py:621-631 (MelSpectrogram.forward)
 *   TA/functional/functional.py:54-145   (spectrogram: pad, stft, abs/pow)
 *   TA/transforms/_transforms.py:417     (MelScale matmul)
 *   + log(clamp(x, min=eps)).
 * ------------------------------------------------------------------------- */
#define YMT3_CODEC_MELSPEC 0
#define YMT3_CODEC_SPEC 1

typedef struct ymt3_audio_cfg {
  int32_t n_fft;       /* 2048 (only value supported by the fused kernel)           */
  int32_t hop_length;  /* 128 (melspec / MT3) or 300 (spec / Perceiver-TF)          */
  int32_t codec;       /* YMT3_CODEC_MELSPEC | YMT3_CODEC_SPEC                      */
  int32_t n_mels;      /* melspec: number of mel filters (512)                      */
  int32_t spec_bin0;   /* spec: first linear bin kept (1 = drop DC)                 */
  int32_t spec_bins;   /* spec: number of bins kept (1024)                          */
  int32_t power_mode;  /* 1 -> |X|, 2 -> |X|^2                                      */
  float log_eps;       /* clamp floor before log                                    */
} ymt3_audio_cfg_t;

typedef struct ymt3_frontend ymt3_frontend_t;

/* window_host: n_fft floats (module buffer `spectrogram.window`).
 * fb_host: melspec only, dense (n_fft/2+1, n_mels) row-major filterbank (module
 * buffer `mel_scale.fb`, TA/functional/functional.py:518-587); nonzeros of each
 * column must be contiguous in frequency (true for triangular filters). */
YMT3_API int ymt3_frontend_create(const ymt3_audio_cfg_t* cfg, const float* window_host,
                         const float* fb_host, ymt3_frontend_t** out);
YMT3_API int ymt3_frontend_destroy(ymt3_frontend_t* fe);
/* number of frames for L samples (center=True): 1 + L / hop */
YMT3_API int64_t ymt3_frontend_num_frames(const ymt3_frontend_t* fe, int64_t L);
/* feature width F of the output */
YMT3_API int64_t ymt3_frontend_num_features(const ymt3_frontend_t* fe);

/* audio_dev: (B, L) f32 contiguous. out_dev: (B, T, F) f32 contiguous. */
YMT3_API int ymt3_logmel_f32(ymt3_frontend_t* fe, const float* audio_dev, int64_t B, int64_t L,
                    float* out_dev, void* stream);
/* Whole-waveform entry: replaces upstream utils/audio.py slice_padded_array(audio, L, L) +
 * the spectrogram layer.  wave_dev: (n_samples) f32 mono 16 kHz; segments are [b*seg_len, (b+1)*seg_len) with the
 * tail zero-padded; out_dev: (ymt3_num_segments(n_samples, seg_len), T, F).  No sliced copy is materialised. */
YMT3_API int64_t ymt3_num_segments(int64_t n_samples, int64_t seg_len);
YMT3_API int ymt3_logmel_waveform_f32(ymt3_frontend_t* fe, const float* wave_dev, int64_t n_samples, int64_t seg_len,
                                      float* out_dev, void* stream);
/* Same with HOST buffers: H2D + kernel + D2H + stream sync (the end-to-end call). */
YMT3_API int ymt3_logmel_host_f32(ymt3_frontend_t* fe, const float* audio_host, int64_t B, int64_t L,
                         float* out_host, void* stream);


/* ------------------------------------------------------------------------- *
 * Named tensors: how weights cross the boundary.  `name` is the state-dict key
 * of the reference module (HF/upstream naming), `data` a DEVICE pointer to a
 * contiguous fp32 tensor.  Handles copy/pack what they need at create time
 * (into fp32 or bf16 according to `precision`); the caller may free or reuse
 * its tensors afterwards.
 * ------------------------------------------------------------------------- */
#define YMT3_DTYPE_F32 0
#define YMT3_DTYPE_BF16 1

typedef struct ymt3_tensor {
  const char* name;
  const void* data; /* device pointer */
  int32_t dtype;    /* YMT3_DTYPE_F32 */
  int32_t ndim;
  int64_t shape[4];
} ymt3_tensor_t;

/* ------------------------------------------------------------------------- *
 * T5 encoder stack (MT3 encoder).  Replaces upstream model/t5mod.py
 * T5EncoderYMT3.forward(inputs_embeds=...) -> last_hidden_state, a modified copy of
 *   HF/models/t5/modeling_t5.py:617-792 (T5Stack), :411-498 (T5Block),
 *   :153-344 (T5Attention, no 1/sqrt(d) scale, fp32 softmax), :106-131 (gated-GELU FF),
 *   :46-68 (T5LayerNorm = RMSNorm).
 * Tensor names: block.{i}.layer.0.SelfAttention.{q,k,v,o}.weight, block.{i}.layer.0.layer_norm.weight,
 * block.{i}.layer.1.DenseReluDense.{wi_0,wi_1,wo}.weight, block.{i}.layer.1.layer_norm.weight,
 * final_layer_norm.weight, and (optional) pos_table (n_pos, d_model) added to inputs_embeds, and (optional)
 * relative_bias_by_distance (num_heads, 2P-1) fp32: T5's relative attention bias (modeling_t5.py:189-268; computed in
 * block 0, shared by all layers, :755-760) folded per distance by the host module from
 * block.0.layer.0.SelfAttention.relative_attention_bias.weight: score(i, j) += table[h][P - 1 + (j - i)], S <= P.
 * ------------------------------------------------------------------------- */
typedef struct ymt3_t5_cfg {
  int32_t precision;   /* YMT3_DTYPE_F32: fp32 FFMA path (token-exact); YMT3_DTYPE_BF16: tcgen05 path */
  int32_t d_model, num_heads, d_kv, d_ff, num_layers;
  float layer_norm_eps;
  /* decoder only */
  int32_t vocab_size, max_length, tie_word_embeddings;
  int32_t eos_id, pad_id, start_id;
} ymt3_t5_cfg_t;

typedef struct ymt3_t5enc ymt3_t5enc_t;
YMT3_API int ymt3_t5enc_create(const ymt3_t5_cfg_t* cfg, const ymt3_tensor_t* tensors, int n_tensors,
                               ymt3_t5enc_t** out);
YMT3_API int ymt3_t5enc_destroy(ymt3_t5enc_t* enc);
/* x_dev: (B, S, d_model) fp32 inputs_embeds. out_dev: (B, S, d_model) in the handle's precision
 * (fp32, or bf16 when precision = BF16). Workspace grows on demand (cudaMalloc) when B*S exceeds
 * the largest size seen so far. */
YMT3_API int ymt3_t5enc_forward(ymt3_t5enc_t* enc, const float* x_dev, int64_t B, int64_t S, void* out_dev,
                                void* stream);

/* ------------------------------------------------------------------------- *
 * T5 decoder + greedy segment-batched generation.  Replaces upstream
 * model/t5mod.py T5DecoderYMT3 / MultiChannelT5Decoder (channels folded into the batch),
 * model/t5mod_helper.py task_cond_dec_generate(), model/lm_head.py LMHead:
 *   HF/models/t5/modeling_t5.py:380-408 (cross-attention layer), :269-305 (KV cache update),
 *   :1105-1110 (d_model**-0.5 logit scale when embeddings are tied).
 * Extra tensor names: embed_tokens.weight (V, d_model), lm_head.weight (V, d_model) [absent = tied],
 * block.{i}.layer.1.EncDecAttention.{q,k,v,o}.weight, block.{i}.layer.2.*, pos_table, and (optional)
 * relative_bias_by_distance (num_heads, P >= max_length) fp32: self-attention bias of the query at position s over key
 * j = table[h][s - j] (unidirectional buckets; cross-attention has no position bias, modeling_t5.py:387-408).
 * The whole loop runs on the device: KV cache, finished mask, step counter and the token
 * matrix never leave HBM; there is no host synchronisation inside ymt3_t5dec_generate.
 * ------------------------------------------------------------------------- */
typedef struct ymt3_t5dec ymt3_t5dec_t;
YMT3_API int ymt3_t5dec_create(const ymt3_t5_cfg_t* cfg, const ymt3_tensor_t* tensors, int n_tensors,
                               ymt3_t5dec_t** out);
YMT3_API int ymt3_t5dec_destroy(ymt3_t5dec_t* dec);
/* enc_hs_dev: (N, T_enc, d_model) in the handle's precision; N = B*C sequences.
 * tokens_out_dev: (N, max_len) int32, fully written (pad after EOS).
 * stop_at_eos: rows that emitted EOS emit pad afterwards (reference semantics).
 * early_stop_interval > 0: every that many steps the host polls a device counter (one 4-byte
 * async D2H + event) and stops launching once every row has finished; 0 = run max_len steps. */
YMT3_API int ymt3_t5dec_generate(ymt3_t5dec_t* dec, const void* enc_hs_dev, int64_t N, int64_t T_enc,
                                 int32_t max_len, int32_t stop_at_eos, int32_t early_stop_interval,
                                 int32_t* tokens_out_dev, void* stream);
/* Task-conditioned variant (reference `prefix_ids` / `task_tokens`): prefix_ids_dev (N, P) int32.  The decoder
 * input sequence is [start, prefix..., generated...]: the P prefix tokens are teacher-forced (nothing is emitted
 * while they are consumed), then max_len tokens are generated.  P + max_len <= cfg.max_length. */
YMT3_API int ymt3_t5dec_generate_prefixed(ymt3_t5dec_t* dec, const void* enc_hs_dev, int64_t N, int64_t T_enc,
                                          const int32_t* prefix_ids_dev, int32_t P, int32_t max_len,
                                          int32_t stop_at_eos, int32_t early_stop_interval, int32_t* tokens_out_dev,
                                          void* stream);
/* Absorbed ("latent") cross-attention, bf16 handles only.  When the encoder->decoder projection is an affine map of a
 * narrow latent (upstream projection_layer.py `mc_shared_linear` [RECALL]: enc_hs[b,c,t,:] = Wp z[b,t,c,:] + bp with
 * z = the channel's k2 x d_latent = 256 latent values), the K/V projections of every decoder layer are linear in the same
 * z, so HF modeling_t5.py:269-305 (q K^T, softmax, V, o) can be evaluated without materialising K/V:
 *     block.{i}.layer.1.EncDecAttention.q_absorbed.weight (H*zdim, d_model): rows h*zdim.. = (Wk_h Wp)^T Wq_h
 *     block.{i}.layer.1.EncDecAttention.o_absorbed.weight (d_model, H*zdim): cols h*zdim.. = Wo_h Wv_h Wp
 *     block.{i}.layer.1.EncDecAttention.o_absorbed.bias   (d_model)        : Wo Wv bp
 * (the Wk bp term is constant over keys and cancels in the softmax).  If these tensors are present in the table
 * given to ymt3_t5dec_create the handle also accepts ymt3_t5dec_generate_latent, which takes the ENCODER LATENTS
 * latents_dev (B, T_enc, C, zdim) instead of the projected hidden states (N = B*C sequences, channel c of segment b
 * is sequence b*C + c) and reads one zdim-wide row per encoder token per layer instead of 2*H*d_kv: 24x fewer
 * cross-attention bytes per decode step for YPTF.MoE+Multi.  Same function, different association order, hence
 * bf16 only; the fp32 token-exact path keeps the reference order.  zdim must be 256, H <= 8. */
YMT3_API int ymt3_t5dec_generate_latent(ymt3_t5dec_t* dec, const void* latents_dev, int64_t B, int64_t T_enc, int32_t C,
                                        const int32_t* prefix_ids_dev, int32_t P, int32_t max_len, int32_t stop_at_eos,
                                        int32_t early_stop_interval, int32_t* tokens_out_dev, void* stream);
/* Teacher-forced scoring.  Replaces upstream model/ymt3.py YourMT3.forward(x, target_tokens) on the decoder side
 * (decoder(inputs_embeds = embed(shift_right(labels)), encoder_hidden_states) -> lm_head logits; HF
 * modeling_t5.py:637-792 with a causal mask, evaluated here incrementally over the KV cache with the SAME kernels and
 * CUDA graph as generation).  forced_ids_dev (N, L) int32: the decoder inputs are [start, forced[:, 0..L-2]]; step t
 * predicts forced[:, t].  argmax_out_dev (N, L) int32 (optional): the model's greedy choice at every step;
 * logit_steps_host (n_logit_steps HOST ints in [0, L)) + logits_out_dev (n_logit_steps, N, vocab) fp32 (optional):
 * the logits of the selected steps.  C == 0: enc_dev is (N = B, T_enc, d_model) hidden states; C > 0: enc_dev is
 * the latent array (B, T_enc, C, zdim) of ymt3_t5dec_generate_latent and N = B*C.  L <= cfg.max_length. */
YMT3_API int ymt3_t5dec_score_forced(ymt3_t5dec_t* dec, const void* enc_dev, int64_t B, int64_t T_enc, int32_t C,
                                     const int32_t* forced_ids_dev, int32_t L, int32_t* argmax_out_dev,
                                     const int32_t* logit_steps_host, int32_t n_logit_steps, float* logits_out_dev,
                                     void* stream);
/* fp32 logits of the LAST executed step, (N, vocab) (for logit-tolerance tests) */
YMT3_API int ymt3_t5dec_last_logits(ymt3_t5dec_t* dec, float* logits_out_dev, int64_t N, void* stream);


/* ------------------------------------------------------------------------- *
 * Pre-encoder "res3b" of the Perceiver-TF models.  Replaces upstream model/conv_block.py
 * PreEncoderBlockRes3B.forward: (B, T, F) -> (B, T, F/8, C): three pre-activation residual
 * 3x3 conv blocks (BatchNorm eval -> ReLU -> conv, twice, + 1x1 shortcut when the channel
 * count changes) each followed by AvgPool2d((1, 2)) over frequency [RECALL; arithmetic =
 * torch.nn.functional.conv2d / batch_norm / avg_pool2d].
 * Tensor names: blocks.{i}.bn{1,2}.{weight,bias,running_mean,running_var},
 * blocks.{i}.conv{1,2}.weight (Cout,Cin,3,3), blocks.{i}.shortcut.{weight (Cout,Cin,1,1),bias}.
 * ------------------------------------------------------------------------- */
typedef struct ymt3_res3b_cfg {
  int32_t precision;     /* F32: im2col + fp32 FFMA GEMM;  BF16: implicit GEMM on tcgen05 */
  int32_t in_freq;       /* F (1024); must be a multiple of 1024 in BF16 mode (128-row M tiles) */
  int32_t channels[3];   /* {64, 128, 128} */
  float bn_eps;
} ymt3_res3b_cfg_t;
typedef struct ymt3_res3b ymt3_res3b_t;
YMT3_API int ymt3_res3b_create(const ymt3_res3b_cfg_t* cfg, const ymt3_tensor_t* tensors, int n_tensors,
                               ymt3_res3b_t** out);
YMT3_API int ymt3_res3b_destroy(ymt3_res3b_t* h);
/* spec_dev: (B, T, F) fp32. out_dev: (B, T, F/8, channels[2]) in the handle's precision. */
YMT3_API int ymt3_res3b_forward(ymt3_res3b_t* h, const float* spec_dev, int64_t B, int64_t T, void* out_dev,
                                void* stream);


/* ------------------------------------------------------------------------- *
 * Perceiver-TF encoder (YPTF).  Replaces upstream model/perceiver_mod.py PerceiverTFEncoder.forward:
 * (B, T, F', C) conv features -> (B, T, K, D) latents; per block: spectral cross-attention
 * (latents <- F' frequency tokens, per time step) -> N latent self-attention layers (over K, per
 * time step) -> M temporal self-attention layers (over T, per latent); final norm.
 * Layer arithmetic = HF/models/perceiver/modeling_perceiver.py:135-242, 255-331, 334-350, 353-414;
 * MoE feed-forward = HF/models/mixtral/modeling_mixtral.py:62-135; RoPE = :208-254.
 * Tensor names: latent_array.latents, [latent_pos_emb, temporal_pos_emb], layernorm.{weight,bias},
 * block.{b}.{sca | local.{n} | temporal.{m}}.attention.self.{layernorm1,layernorm2,query,key,value}.*,
 * ....attention.output.dense.*, ....layernorm.*, ....mlp.{dense1,dense2}.* or
 * ....moe.gate.weight + ....moe.experts.{e}.{w1,w2,w3}.weight.
 * ------------------------------------------------------------------------- */
typedef struct ymt3_ptf_cfg {
  int32_t precision;
  int32_t num_latents, d_latent, kv_dim, num_blocks, num_local, num_temporal;
  int32_t cross_heads, self_heads;
  int32_t sca_query_residual;
  int32_t norm_type;     /* 0 = LayerNorm, 1 = RMSNorm (weight only) */
  int32_t ff_type;       /* 0 = MLP, 1 = MoE */
  int32_t ff_widening, moe_experts, moe_topk;
  int32_t act;           /* activation code as in ymt3_op_linear */
  int32_t pos_type;      /* 0 none, 1 trainable tables, 2 rotary */
  int32_t rope_dim;      /* rotated dims per self-attention head (pos_type 2) */
  int32_t max_time;      /* largest T (rotary / temporal table length) */
  float norm_eps;
} ymt3_ptf_cfg_t;
typedef struct ymt3_ptf ymt3_ptf_t;
YMT3_API int ymt3_ptf_create(const ymt3_ptf_cfg_t* cfg, const ymt3_tensor_t* tensors, int n_tensors, ymt3_ptf_t** out);
YMT3_API int ymt3_ptf_destroy(ymt3_ptf_t* h);
/* x_dev: (B, T, Fp, kv_dim) in the handle's precision. out_dev: (B, T, num_latents, d_latent) same precision. */
YMT3_API int ymt3_ptf_forward(ymt3_ptf_t* h, const void* x_dev, int64_t B, int64_t T, int64_t Fp, void* out_dev,
                              void* stream);
/* (B, T, C, D) -> (B, C, T, D) in `dtype` (multi-channel pre-decoder layout change) */
YMT3_API int ymt3_op_permute_btcd_bctd(int32_t dtype, const void* x, void* y, int64_t B, int64_t T, int64_t C, int64_t D,
                                       void* stream);
/* dtype conversion of a flat buffer (f32 <-> bf16) */
YMT3_API int ymt3_op_convert(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t n, void* stream);

/* ------------------------------------------------------------------------- *
 * Per-op entry points (unit-parity tests of individual kernels; also usable as building
 * blocks).  dtype: YMT3_DTYPE_F32 | YMT3_DTYPE_BF16 for A/W/C; bias always fp32.
 * ------------------------------------------------------------------------- */
/* C = residual + out_scale * epi(A @ W^T + bias);  A (M,K) lda, W (N,K) ldw (nn.Linear.weight),
 * act: 0 none, 1 gelu_new(tanh), 2 relu, 3 silu, 4 gelu(erf); gated: W rows interleaved (act, lin) pairs,
 * C is (M, N/2). */
YMT3_API int ymt3_op_linear(int32_t dtype, const void* A, int64_t lda, const void* W, int64_t ldw,
                            const float* bias, void* C, int64_t ldc, const void* residual, int64_t ldr,
                            int64_t M, int64_t N, int64_t K, int32_t act, int32_t gated, float out_scale,
                            int32_t out_dtype, void* stream);
YMT3_API int ymt3_op_rmsnorm(int32_t dtype, const void* x, const float* w, void* y, int64_t rows, int64_t dim,
                             float eps, void* stream);
YMT3_API int ymt3_op_layernorm(int32_t dtype, const void* x, const float* w, const float* b, void* y,
                               int64_t rows, int64_t dim, float eps, void* stream);
/* q/k/v/o: (B, S, H, dk) contiguous; causal: key j visible iff j <= i + (Sk - Sq) */
YMT3_API int ymt3_op_attention(int32_t dtype, const void* q, const void* k, const void* v, void* o, int64_t B,
                               int64_t H, int64_t Sq, int64_t Sk, int64_t dk, float scale, int32_t causal,
                               void* stream);
/* bf16 tensor-core linear with the fused-RMSNorm hooks of the decode step (T5LayerNorm, HF modeling_t5.py:46-68, fused
 * into the GEMMs around it).  Consumer side (ss_in != NULL): A is the UN-normalised x (M, K) bf16, W has the norm
 * weight folded into its columns (W[n,k] * w_ln[k]), and every accumulator row is scaled by
 * rsqrt(sum_c ss_in[m, c] / K + eps) before bias / activation; ss_in is (M, chunks) fp32, chunks % 4 == 0.
 * Producer side (ss_out != NULL, bf16 non-gated output, N % 32 == 0): ss_out[m, n/32] = sum of squares of the
 * bf16-rounded outputs C[m, 32*(n/32) .. +31] (after the residual add). */
YMT3_API int ymt3_op_linear_normfused(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                      const float* ss_in, int64_t chunks, float eps, void* C, int64_t ldc,
                                      const void* residual, int64_t ldr, float* ss_out, int64_t M, int64_t N, int64_t K,
                                      int32_t act, int32_t gated, float out_scale, int32_t out_dtype, void* stream);
/* Several DEPENDENT normfused linears in ONE persistent launch (the GEMMs between two attention kernels of a decode
 * step: [o-proj + residual -> cross-q] and [cross-o + residual -> wi (gated) -> wo + residual -> next layer's qkv]).
 * Phase i > 0 reads what phase i - 1 wrote (its A is an earlier phase's C); every phase has the same M, bf16 in / out,
 * fields as in ymt3_op_linear_normfused.  counters: ymt3_op_linear_chain_counters(M) device int32, zero before the
 * first launch that uses them; `ordinal` = number of earlier launches on the same counters since they were zeroed.
 * Result bit-identical to the same phases issued one by one through ymt3_op_linear_normfused. */
typedef struct ymt3_chain_phase {
  const void* A; int64_t lda;
  const void* W; int64_t ldw;
  const float* bias;
  const float* ss_in; int64_t chunks; float eps;
  void* C; int64_t ldc;
  const void* residual; int64_t ldr;
  float* ss_out;
  int64_t N, K;
  int32_t act, gated;
  float out_scale;
} ymt3_chain_phase_t;
YMT3_API int64_t ymt3_op_linear_chain_counters(int64_t M);
/* debug aid (tools/trace_chain.py): later chain launches write globaltimer stamps [CTA][16 tiles][32 events] into
 * device_buf (>= SMs * 512 uint64, zeroed by the caller); NULL switches it off. */
YMT3_API int ymt3_debug_chain_trace(uint64_t* device_buf);
YMT3_API int ymt3_op_linear_chain(const ymt3_chain_phase_t* phases, int32_t n_phases, int64_t M, int32_t* counters,
                                  int32_t ordinal, void* stream);
/* Vocab projection with the greedy selection fused into its epilogue (the LM head of the decode step; HF
 * modeling_t5.py:1105-1110 followed by torch.argmax): logits (M, N) fp32 = out_scale * (A @ W^T + bias) are stored AND
 * every row's arg-max over columns [0, V) is left in keys[m] as  (ordered(logit) << 32) | (0xFFFFFFFF - column)  via
 * atomicMax - largest logit, smallest column among equals (torch.argmax's first-maximum rule); column =
 * 0xFFFFFFFF - (keys[m] & 0xFFFFFFFF).  keys (M) must be zero before the call. */
YMT3_API int ymt3_op_linear_argmax(int32_t dtype, const void* A, int64_t lda, const void* W, int64_t ldw,
                                   const float* bias, float* logits, int64_t ldc, int64_t M, int64_t N, int64_t K,
                                   int64_t V, float out_scale, uint64_t* keys, void* stream);
/* Single-query attention over a device-resident KV cache alone (the decode step's dominant kernel; HF
 * modeling_t5.py:269-305 with use_cache): q (N, H*64); cache K/V (N, H, Lcap, 64).  knew/vnew (N, H*64) non-null:
 * self mode - the row is appended at index *step_dev, then keys [0, *step_dev] are attended.  knew == NULL: cross
 * mode over keys [0, fixed_len).  out (N, H*64).  No scaling (T5).  Contract: *step_dev < Lcap; a launch whose step
 * counter is at or beyond the cache capacity writes nothing (no append, `out` untouched) instead of running out of
 * bounds. */
YMT3_API int ymt3_op_decode_attention(int32_t dtype, const void* q, const void* knew, const void* vnew, void* Kc,
                                      void* Vc, const int32_t* step_dev, int64_t fixed_len, void* out, int64_t N,
                                      int64_t H, int64_t Lcap, void* stream);
/* Absorbed cross-attention kernel alone (see ymt3_t5dec_generate_latent): q (N, H*256) bf16, z (N, Tp, 256) bf16 with
 * rows >= T zero, out (N, H*256) bf16 = per head softmax_t(q_h . z_t) z_t;  Tp % 16 == 0, H <= 8. */
YMT3_API int ymt3_op_cross_attn_absorbed(const void* q, const void* z, void* out, int64_t N, int64_t H, int64_t T,
                                         int64_t Tp, void* stream);

/* Sparse MoE feed-forward alone (upstream model/ff_layer.py MoE [RECALL] = HF modeling_mixtral.py:62-135): out =
 * residual + sum_k w_k * W2[e_k](act(W1[e_k] x) * W3[e_k] x) with (w, e) = renormalised top-k of the fp32 router softmax.
 * x / residual (or NULL) / out: (N, D) in `dtype`; gate (E, D) fp32; w13 (E, 2*I, D) in `dtype`, rows interleaved
 * (2j = w1[j] = the activated half, 2j+1 = w3[j]); w2 (E, D, I) in `dtype`; act as in ymt3_op_linear.
 * workspace: ymt3_op_moe_workspace_bytes(...) device bytes, caller-owned.  Fused router + top-k, device-side expert
 * sort (no host sync), two grouped GEMMs (fp32 FFMA or tcgen05), gather-combine; deterministic. */
YMT3_API int64_t ymt3_op_moe_workspace_bytes(int64_t N, int32_t D, int32_t I, int32_t E, int32_t topk, int32_t dtype);
YMT3_API int ymt3_op_moe_ff(int32_t dtype, const void* x, const void* residual, void* out, int64_t N,
                            const float* gate, const void* w13, const void* w2, int32_t D, int32_t I, int32_t E,
                            int32_t topk, int32_t act, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YMT3_B200_H */
