// T5 encoder stack and T5 decoder + greedy generation, orchestrated natively (C++) over the
// kernels in gemm_*.cu / attention.cu / ops_misc.cu / decode.cu. Reference semantics:
// HF transformers modeling_t5.py (cited per step below) as copied by upstream model/t5mod.py.
#include "model_common.cuh"
#include "decode.cuh"
#include <math.h>
#include <stdlib.h>

using namespace ymt3;

namespace {

struct T5Layer {
  float* ln_sa = nullptr;  Linear qkv, o;           // layer.0 SelfAttention (q|k|v stacked)
  float* ln_ca = nullptr;  Linear xq, xkv, xo;      // layer.1 EncDecAttention (decoder only; k|v stacked)
  Linear xq_abs, xo_abs;                            // optional absorbed cross-attention (latent-space q / o + bias)
  // bf16 decoder only: copies of the norm-consuming weights with the RMSNorm scale folded into their columns
  // (fused RMSNorm, GemmParams::norm_ss_in): qkv <- ln_sa, xq / xq_abs <- ln_ca, wi <- ln_ff
  Linear qkv_n, xq_n, xq_abs_n, wi_n;
  float* ln_ff = nullptr;  Linear wi, wo;           // DenseReluDense (wi_0/wi_1 interleaved -> gated epilogue)
};

int load_layers(const ymt3_t5_cfg_t& c, const TensorTable& tt, bool decoder, DevicePool& pool,
                std::vector<T5Layer>& layers, float** final_ln, cudaStream_t s) {
  const int D = c.d_model, inner = c.num_heads * c.d_kv, F = c.d_ff;
  layers.resize(c.num_layers);
  for (int i = 0; i < c.num_layers; ++i) {
    T5Layer& L = layers[i];
    const std::string b = "block." + std::to_string(i) + ".layer.";
    const std::string sa = b + "0.SelfAttention.";
    int rc;
    if ((rc = pack_vec(pool, {tt.require(b + "0.layer_norm.weight", D)}, false, &L.ln_sa, s))) return rc;
    if ((rc = pack_rows(pool, {tt.require(sa + "q.weight", inner, D), tt.require(sa + "k.weight", inner, D),
                               tt.require(sa + "v.weight", inner, D)}, D, c.precision, false, &L.qkv, s))) return rc;
    if ((rc = pack_rows(pool, {tt.require(sa + "o.weight", D, inner)}, inner, c.precision, false, &L.o, s))) return rc;
    std::string ff = b + "1.";
    if (decoder) {
      const std::string ca = b + "1.EncDecAttention.";
      if ((rc = pack_vec(pool, {tt.require(b + "1.layer_norm.weight", D)}, false, &L.ln_ca, s))) return rc;
      if ((rc = pack_rows(pool, {tt.require(ca + "q.weight", inner, D)}, D, c.precision, false, &L.xq, s))) return rc;
      if ((rc = pack_rows(pool, {tt.require(ca + "k.weight", inner, D), tt.require(ca + "v.weight", inner, D)}, D,
                          c.precision, false, &L.xkv, s))) return rc;
      if ((rc = pack_rows(pool, {tt.require(ca + "o.weight", D, inner)}, inner, c.precision, false, &L.xo, s))) return rc;
      ff = b + "2.";
    }
    if ((rc = pack_vec(pool, {tt.require(ff + "layer_norm.weight", D)}, false, &L.ln_ff, s))) return rc;
    if ((rc = pack_rows(pool, {tt.require(ff + "DenseReluDense.wi_0.weight", F, D),
                               tt.require(ff + "DenseReluDense.wi_1.weight", F, D)}, D, c.precision, true, &L.wi, s)))
      return rc;
    if ((rc = pack_rows(pool, {tt.require(ff + "DenseReluDense.wo.weight", D, F)}, F, c.precision, false, &L.wo, s)))
      return rc;
    if (decoder && c.precision == YMT3_BF16) {
      const std::string ca = b + "1.EncDecAttention.";
      if ((rc = pack_rows(pool, {tt.require(sa + "q.weight", inner, D), tt.require(sa + "k.weight", inner, D),
                                 tt.require(sa + "v.weight", inner, D)}, D, c.precision, false, &L.qkv_n, s, L.ln_sa))) return rc;
      if ((rc = pack_rows(pool, {tt.require(ca + "q.weight", inner, D)}, D, c.precision, false, &L.xq_n, s, L.ln_ca))) return rc;
      if ((rc = pack_rows(pool, {tt.require(ff + "DenseReluDense.wi_0.weight", F, D),
                                 tt.require(ff + "DenseReluDense.wi_1.weight", F, D)}, D, c.precision, true, &L.wi_n, s,
                          L.ln_ff))) return rc;
    }
  }
  return pack_vec(pool, {tt.require("final_layer_norm.weight", D)}, false, final_ln, s);
}

int check_cfg(const ymt3_t5_cfg_t* c) {
  YMT3_REQUIRE(c, "t5: null config");
  YMT3_REQUIRE(c->precision == YMT3_F32 || c->precision == YMT3_BF16, "t5: bad precision %d", c->precision);
  YMT3_REQUIRE(c->d_model > 0 && c->d_model % 8 == 0 && c->num_heads > 0 && c->d_kv > 0 && c->d_kv % 4 == 0 &&
                   c->d_ff > 0 && c->d_ff % 8 == 0 && c->num_layers > 0,
               "t5: bad dimensions");
  return YMT3_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// encoder
// ------------------------------------------------------------------------------------------
struct ymt3_t5enc {
  ymt3_t5_cfg_t c;
  DevicePool weights, ws;
  std::vector<T5Layer> layers;
  float* final_ln = nullptr;
  void* pos = nullptr;
  int n_pos = 0;
  float* rel_bias = nullptr;   // (H, 2 * rel_P - 1) relative attention bias per distance, or null
  int rel_P = 0;
  int64_t cap_rows = 0;
  void *x = nullptr, *h = nullptr, *qkv = nullptr, *attn = nullptr, *g = nullptr;
};

extern "C" int ymt3_t5enc_create(const ymt3_t5_cfg_t* cfg, const ymt3_tensor_t* tensors, int n, ymt3_t5enc_t** out) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  YMT3_REQUIRE(tensors && out, "t5enc_create: null argument");
  ymt3_t5enc* e = new ymt3_t5enc();
  e->c = *cfg;
  TensorTable tt{tensors, n};
  rc = load_layers(*cfg, tt, false, e->weights, e->layers, &e->final_ln, 0);
  if (!rc) {
    if (const ymt3_tensor_t* p = tt.find("pos_table")) {
      if (p->ndim != 2 || p->shape[1] != cfg->d_model) {
        ymt3_set_error("t5enc_create: pos_table must be (n_pos, d_model)");
        rc = YMT3_ERR_INVALID;
      } else {
        e->n_pos = (int)p->shape[0];
        // the encoder's residual stream is fp32 in both precisions (see ymt3_t5enc_forward): fp32 position table
        rc = pack_table(e->weights, (const float*)p->data, false, p->shape[0] * p->shape[1], YMT3_F32, &e->pos, 0);
      }
    }
  }
  if (!rc) {
    // T5 relative attention bias (HF modeling_t5.py:189-268), folded per distance by the host module:
    // relative_bias_by_distance[h][P - 1 + (j - i)], bidirectional buckets
    if (const ymt3_tensor_t* rb = tt.find("relative_bias_by_distance")) {
      if (rb->ndim != 2 || rb->shape[0] != cfg->num_heads || rb->shape[1] % 2 != 1) {
        ymt3_set_error("t5enc_create: relative_bias_by_distance must be (num_heads, 2 * P - 1)");
        rc = YMT3_ERR_INVALID;
      } else {
        e->rel_P = (int)((rb->shape[1] + 1) / 2);
        void* t = nullptr;
        rc = pack_table(e->weights, (const float*)rb->data, false, rb->shape[0] * rb->shape[1], YMT3_F32, &t, 0);
        e->rel_bias = (float*)t;
      }
    }
  }
  if (!rc && cudaStreamSynchronize(0) != cudaSuccess) {
    ymt3_set_error("t5enc_create: weight packing failed: %s", cudaGetErrorString(cudaGetLastError()));
    rc = YMT3_ERR_CUDA;
  }
  if (rc) {
    e->weights.release();
    delete e;
    return rc;
  }
  *out = e;
  return YMT3_OK;
}

extern "C" int ymt3_t5enc_destroy(ymt3_t5enc_t* e) {
  if (!e) return YMT3_OK;
  e->weights.release();
  e->ws.release();
  delete e;
  return YMT3_OK;
}

extern "C" int ymt3_t5enc_forward(ymt3_t5enc_t* e, const float* x_in, int64_t B, int64_t S, void* out, void* stream) {
  YMT3_REQUIRE(e && out, "t5enc_forward: null argument");
  if (B <= 0 || S <= 0) return YMT3_OK;
  YMT3_REQUIRE(x_in, "t5enc_forward: null input");
  const ymt3_t5_cfg_t& c = e->c;
  const int D = c.d_model, H = c.num_heads, dk = c.d_kv, inner = H * dk, F = c.d_ff, dt = c.precision;
  const int64_t M = B * S;
  YMT3_REQUIRE(M < (1ll << 31), "t5enc_forward: too many tokens");
  YMT3_REQUIRE(!e->pos || S <= e->n_pos, "t5enc_forward: sequence %lld longer than pos_table %d", (long long)S, e->n_pos);
  YMT3_REQUIRE(!e->rel_bias || S <= e->rel_P, "t5enc_forward: sequence %lld longer than the relative bias table %d",
               (long long)S, e->rel_P);
  cudaStream_t s = (cudaStream_t)stream;
  // bf16 precision: GEMM operands / attention in bf16 (tcgen05), the RESIDUAL STREAM x in fp32 - what torch autocast
  // does with the reference modules.  A bf16 residual stream loses ~2^-9 of |x| at every add while the layer updates
  // are a fraction of |x|; measured on the unreduced 8-layer encoder (log-mel input, random weights) the relative L2
  // error of the hidden states vs the fp32 oracle was 9.3 % with a bf16 stream (profiles/r02_diag_t5enc_bf16.txt).
  const bool mixed = dt == YMT3_BF16;
  if (M > e->cap_rows) {
    YMT3_CUDA_CHECK(cudaStreamSynchronize(s));
    e->ws.release();
    const size_t es = dtype_size(dt);
    e->x = e->ws.alloc(M * D * 4);
    e->h = e->ws.alloc(M * D * es);
    e->qkv = e->ws.alloc(M * 3 * inner * es);
    e->attn = e->ws.alloc(M * inner * es);
    e->g = e->ws.alloc(M * F * es);
    e->cap_rows = 0;
    if (!e->x || !e->h || !e->qkv || !e->attn || !e->g) return YMT3_ERR_CUDA;
    e->cap_rows = M;
  }
  int rc;
  // inputs_embeds (+ absolute position table [RECALL upstream]); dropout is identity in eval
  if (e->pos) {
    if ((rc = add_rows(x_in, e->pos, e->x, M, (int)S, D, YMT3_F32, s))) return rc;
  } else {
    YMT3_CUDA_CHECK(cudaMemcpyAsync(e->x, x_in, (size_t)M * D * 4, cudaMemcpyDeviceToDevice, s));
  }
  const size_t es = dtype_size(dt);
  auto norm = [&](const float* w, void* y) -> int {
    return mixed ? rmsnorm_f32_bf16((const float*)e->x, w, y, M, D, c.layer_norm_eps, s)
                 : rmsnorm(e->x, w, y, M, D, c.layer_norm_eps, YMT3_F32, s);
  };
  for (const T5Layer& L : e->layers) {
    // T5LayerSelfAttention (modeling_t5.py:356-377): x += o(attn(rmsnorm(x)))
    if ((rc = norm(L.ln_sa, e->h))) return rc;
    if ((rc = linear_fwd(dt, e->h, D, L.qkv, e->qkv, 3 * inner, (int)M, 0, 0, nullptr, 0, 1.f, dt, s))) return rc;
    AttnParams a{};
    a.Q = e->qkv; a.K = (char*)e->qkv + inner * es; a.V = (char*)e->qkv + 2 * inner * es;
    a.q_sb = a.k_sb = a.v_sb = S * 3 * inner; a.q_sh = a.k_sh = a.v_sh = dk; a.q_ss = a.k_ss = a.v_ss = 3 * inner;
    a.O = e->attn; a.o_sb = S * inner; a.o_sh = dk; a.o_ss = inner;
    a.B = (int)B; a.H = H; a.Sq = (int)S; a.Sk = (int)S; a.dk = dk;
    a.scale = 1.0f;  // T5: no 1/sqrt(d) (modeling_t5.py:308)
    // position bias: computed in block 0 and shared by every layer (modeling_t5.py:755-760)
    a.rel_bias = e->rel_bias; a.rel_stride = 2 * e->rel_P - 1; a.rel_center = e->rel_P - 1;
    if ((rc = attention(a, dt, s))) return rc;
    if ((rc = linear_fwd(dt, e->attn, inner, L.o, e->x, D, (int)M, 0, 0, e->x, D, 1.f, YMT3_F32, s))) return rc;
    // T5LayerFF (modeling_t5.py:146-150) with gated-GELU (:115-131)
    if ((rc = norm(L.ln_ff, e->h))) return rc;
    if ((rc = linear_fwd(dt, e->h, D, L.wi, e->g, F, (int)M, YMT3_ACT_GELU_NEW, 1, nullptr, 0, 1.f, dt, s))) return rc;
    if ((rc = linear_fwd(dt, e->g, F, L.wo, e->x, D, (int)M, 0, 0, e->x, D, 1.f, YMT3_F32, s))) return rc;
  }
  return norm(e->final_ln, out);  // final_layer_norm (:767)
}

// ------------------------------------------------------------------------------------------
// decoder + greedy generation
// ------------------------------------------------------------------------------------------
struct ymt3_t5dec {
  ymt3_t5_cfg_t c;
  DevicePool weights, ws;
  std::vector<T5Layer> layers;
  float* final_ln = nullptr;
  void* pos = nullptr;
  int n_pos = 0;
  float* rel_bias = nullptr; // (H, rel_P) self-attention relative bias per distance (query pos - key pos), or null
  int rel_P = 0;
  void* embed = nullptr;     // (V, D) compute dtype
  Linear lm_head;            // (Vp, D), rows >= V are zero
  Linear lm_head_n;          // bf16: final_layer_norm folded in (fused RMSNorm)
  bool fuse_norm = false;    // bf16 && d_model % 128 == 0 && !YMT3_NO_FUSED_NORM
  float* ss[3] = {nullptr, nullptr, nullptr};   // (cap_N, d_model / 32) sum-of-squares partials, rotating
  int Vp = 0;
  // per-(N, T_enc, Lmax) state
  int64_t cap_N = 0, cap_T = 0, cap_L = 0;
  void *x = nullptr, *h = nullptr, *qkv = nullptr, *attn = nullptr, *qx = nullptr, *g = nullptr;
  float* logits = nullptr;
  std::vector<void*> selfK, selfV, crossKV;   // crossKV[i]: K (N, H, T, dk) followed by V (N, H, T, dk)
  void* kv_tmp = nullptr;                      // (N*T, 2*inner) projection output before the head split
  int *d_step = nullptr, *d_cur = nullptr, *d_fin = nullptr, *d_unfinished = nullptr;
  unsigned long long* d_amax = nullptr;   // (cap_N) packed arg-max keys of the fused vocab projection
  int* h_unfinished = nullptr;  // pinned
  // the decode loop runs on an internal stream (graph capture is illegal on the legacy default
  // stream); it is ordered after / before the caller's stream with events, no host sync
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  // cached CUDA graph of one decode step
  cudaGraphExec_t graph = nullptr;
  int64_t graph_N = -1, graph_T = -1, graph_L = -1;
  int graph_stop = -1, graph_prefix = -1, graph_score = -1;
  bool use_chain = false;       // bf16 + fused norm: the GEMMs between the attention kernels run as two chained launches per layer
  int* chain_done = nullptr;    // completion counters of the two chains (gemm_chain_counters each)
  int32_t* tok_buf = nullptr;   // (cap_N, cap_L) tokens of the running call: the graph writes here (stable pointer), one
                                // D2D copy hands them to the caller's buffer, so a fresh output tensor per call does
                                // not re-instantiate the graph
  int* d_forced = nullptr;   // (cap_N, cap_P) task-prefix tokens
  int64_t cap_P = 0;
  // absorbed cross-attention (bf16, cross_absorbed.cu): present when the tensor table carried *.q_absorbed.weight
  int zdim = 0;              // latent width (0 = not available)
  int cap_latent = -1;       // workspace currently laid out for: 0 = K/V mode, 1 = latent mode
  int64_t cap_Tp = 0;
  void *zbuf = nullptr, *qz = nullptr, *cz = nullptr;   // (N, Tp, zdim) latents, (N, H*zdim) queries / contexts
  int graph_latent = -1;
};

extern "C" int ymt3_t5dec_create(const ymt3_t5_cfg_t* cfg, const ymt3_tensor_t* tensors, int n, ymt3_t5dec_t** out) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  YMT3_REQUIRE(tensors && out, "t5dec_create: null argument");
  YMT3_REQUIRE(cfg->d_kv == 64, "t5dec_create: d_kv must be 64 (decode attention kernel)");
  YMT3_REQUIRE(cfg->vocab_size > 1 && cfg->max_length > 0, "t5dec_create: bad vocab/max_length");
  ymt3_t5dec* d = new ymt3_t5dec();
  d->c = *cfg;
  TensorTable tt{tensors, n};
  const int D = cfg->d_model, V = cfg->vocab_size;
  rc = load_layers(*cfg, tt, true, d->weights, d->layers, &d->final_ln, 0);
  if (!rc) {
    // optional absorbed cross-attention weights (host-folded, see include/ymt3_b200.h)
    const std::string k0 = "block.0.layer.1.EncDecAttention.q_absorbed.weight";
    if (const ymt3_tensor_t* q0 = tt.find(k0)) {
      if (cfg->precision != YMT3_BF16 || q0->ndim != 2 || q0->shape[1] != D || q0->shape[0] % cfg->num_heads) {
        ymt3_set_error("t5dec_create: absorbed cross-attention needs bf16 precision and (H*zdim, d_model) weights");
        rc = YMT3_ERR_INVALID;
      } else {
        d->zdim = (int)(q0->shape[0] / cfg->num_heads);
        const int HZ = cfg->num_heads * d->zdim;
        for (int i = 0; !rc && i < cfg->num_layers; ++i) {
          const std::string ca = "block." + std::to_string(i) + ".layer.1.EncDecAttention.";
          T5Layer& L = d->layers[i];
          rc = pack_rows(d->weights, {tt.require(ca + "q_absorbed.weight", HZ, D)}, D, YMT3_BF16, false, &L.xq_abs, 0);
          if (!rc) rc = pack_rows(d->weights, {tt.require(ca + "q_absorbed.weight", HZ, D)}, D, YMT3_BF16, false, &L.xq_abs_n, 0,
                                  L.ln_ca);
          if (!rc) rc = pack_rows(d->weights, {tt.require(ca + "o_absorbed.weight", D, HZ)}, HZ, YMT3_BF16, false, &L.xo_abs, 0);
          if (!rc) rc = pack_vec(d->weights, {tt.require(ca + "o_absorbed.bias", D)}, false, &L.xo_abs.bias, 0);
        }
      }
    }
  }
  if (!rc) {
    const ymt3_tensor_t* E = tt.require("embed_tokens.weight", V, D);
    if (!E) rc = YMT3_ERR_INVALID;
    if (!rc) rc = pack_table(d->weights, (const float*)E->data, false, (int64_t)V * D, cfg->precision, &d->embed, 0);
    if (!rc) {
      const ymt3_tensor_t* Lm = tt.find("lm_head.weight") ? tt.require("lm_head.weight", V, D) : E;
      if (!Lm) rc = YMT3_ERR_INVALID;
      if (!rc) {
        d->Vp = (V + 7) / 8 * 8;
        void* W = d->weights.alloc((size_t)d->Vp * D * dtype_size(cfg->precision));
        if (!W) rc = YMT3_ERR_CUDA;
        if (!rc && cudaMemsetAsync(W, 0, (size_t)d->Vp * D * dtype_size(cfg->precision), 0) != cudaSuccess) rc = YMT3_ERR_CUDA;
        if (!rc) rc = convert(Lm->data, YMT3_F32, W, cfg->precision, (int64_t)V * D, 0);
        d->lm_head.W = W; d->lm_head.N = d->Vp; d->lm_head.K = D;
        if (!rc && cfg->precision == YMT3_BF16) {
          void* Wn = d->weights.alloc((size_t)d->Vp * D * 2);
          if (!Wn) rc = YMT3_ERR_CUDA;
          if (!rc && cudaMemsetAsync(Wn, 0, (size_t)d->Vp * D * 2, 0) != cudaSuccess) rc = YMT3_ERR_CUDA;
          if (!rc) rc = pack_rows_at(Wn, 0, 1, Lm, D, YMT3_BF16, 0, d->final_ln);
          d->lm_head_n.W = Wn; d->lm_head_n.N = d->Vp; d->lm_head_n.K = D;
        }
      }
    }
  }
  if (!rc) {
    if (const ymt3_tensor_t* p = tt.find("pos_table")) {
      if (p->ndim != 2 || p->shape[1] != D || p->shape[0] < cfg->max_length) {
        ymt3_set_error("t5dec_create: pos_table must be (>= max_length, d_model)");
        rc = YMT3_ERR_INVALID;
      } else {
        d->n_pos = (int)p->shape[0];
        rc = pack_table(d->weights, (const float*)p->data, false, p->shape[0] * p->shape[1], cfg->precision, &d->pos, 0);
      }
    }
  }
  if (!rc) {
    // decoder self-attention relative bias (unidirectional buckets), folded per distance by the host module
    if (const ymt3_tensor_t* rb = tt.find("relative_bias_by_distance")) {
      if (rb->ndim != 2 || rb->shape[0] != cfg->num_heads || rb->shape[1] < cfg->max_length) {
        ymt3_set_error("t5dec_create: relative_bias_by_distance must be (num_heads, >= max_length)");
        rc = YMT3_ERR_INVALID;
      } else {
        d->rel_P = (int)rb->shape[1];
        void* t = nullptr;
        rc = pack_table(d->weights, (const float*)rb->data, false, rb->shape[0] * rb->shape[1], YMT3_F32, &t, 0);
        d->rel_bias = (float*)t;
      }
    }
  }
  d->fuse_norm = cfg->precision == YMT3_BF16 && D % 128 == 0 && getenv("YMT3_NO_FUSED_NORM") == nullptr;
  d->use_chain = d->fuse_norm && getenv("YMT3_GEMM_CHAIN") != nullptr;   // opt-in until it beats the separate launches
  if (!rc) {
    d->d_step = (int*)d->weights.alloc(64);
    d->d_unfinished = d->d_step + 4;
    if (!d->d_step) rc = YMT3_ERR_CUDA;
    if (!rc && cudaMallocHost((void**)&d->h_unfinished, 64) != cudaSuccess) rc = YMT3_ERR_CUDA;
    if (!rc && cudaStreamCreateWithFlags(&d->own_stream, cudaStreamNonBlocking) != cudaSuccess) rc = YMT3_ERR_CUDA;
    if (!rc && cudaEventCreateWithFlags(&d->ev_in, cudaEventDisableTiming) != cudaSuccess) rc = YMT3_ERR_CUDA;
    if (!rc && cudaEventCreateWithFlags(&d->ev_out, cudaEventDisableTiming) != cudaSuccess) rc = YMT3_ERR_CUDA;
  }
  if (!rc && cudaStreamSynchronize(0) != cudaSuccess) {
    ymt3_set_error("t5dec_create: weight packing failed: %s", cudaGetErrorString(cudaGetLastError()));
    rc = YMT3_ERR_CUDA;
  }
  if (rc) {
    d->weights.release();
    delete d;
    return rc;
  }
  *out = d;
  return YMT3_OK;
}

extern "C" int ymt3_t5dec_destroy(ymt3_t5dec_t* d) {
  if (!d) return YMT3_OK;
  if (d->graph) cudaGraphExecDestroy(d->graph);
  if (d->h_unfinished) cudaFreeHost(d->h_unfinished);
  if (d->own_stream) cudaStreamDestroy(d->own_stream);
  if (d->ev_in) cudaEventDestroy(d->ev_in);
  if (d->ev_out) cudaEventDestroy(d->ev_out);
  d->weights.release();
  d->ws.release();
  delete d;
  return YMT3_OK;
}

namespace {

int dec_ensure(ymt3_t5dec* d, int64_t N, int64_t T, int64_t Lmax, int latent, cudaStream_t s) {
  if (N <= d->cap_N && T <= d->cap_T && Lmax <= d->cap_L && latent == d->cap_latent) return YMT3_OK;
  YMT3_CUDA_CHECK(cudaStreamSynchronize(s));
  if (d->graph) {
    cudaGraphExecDestroy(d->graph);
    d->graph = nullptr;
    d->graph_N = -1;
  }
  // capacities only ever grow (a ragged last batch or alternating K/V <-> latent calls must not shrink them and
  // force a re-allocation + graph re-capture on the next full batch): take the maxima BEFORE resetting the caps
  const int64_t cN = N > d->cap_N ? N : d->cap_N, cT = T > d->cap_T ? T : d->cap_T, cL = Lmax > d->cap_L ? Lmax : d->cap_L;
  d->ws.release();
  d->d_forced = nullptr;
  d->cap_P = 0;
  d->cap_N = d->cap_T = d->cap_L = 0;
  d->cap_latent = -1;
  const ymt3_t5_cfg_t& c = d->c;
  const int D = c.d_model, inner = c.num_heads * c.d_kv, F = c.d_ff;
  const size_t es = dtype_size(c.precision);
  d->x = d->ws.alloc(cN * D * es);
  d->h = d->ws.alloc(cN * D * es);
  d->qkv = d->ws.alloc(cN * 3 * inner * es);
  d->attn = d->ws.alloc(cN * inner * es);
  d->qx = d->ws.alloc(cN * inner * es);
  d->g = d->ws.alloc(cN * F * es);
  d->logits = (float*)d->ws.alloc(cN * d->Vp * 4);
  const int64_t cTp = (cT + 15) / 16 * 16;
  if (latent) {
    // latent mode: no per-layer K/V at all; one padded latent tile per sequence (pad rows stay zero)
    const size_t zb = (size_t)cN * cTp * d->zdim * es;
    d->kv_tmp = d->zbuf = d->ws.alloc(zb);
    d->qz = d->ws.alloc((size_t)cN * c.num_heads * d->zdim * es);
    d->cz = d->ws.alloc((size_t)cN * c.num_heads * d->zdim * es);
    if (!d->zbuf || !d->qz || !d->cz) d->kv_tmp = nullptr;
  } else {
    d->kv_tmp = d->ws.alloc(cN * cT * 2 * inner * es);
  }
  d->d_cur = (int*)d->ws.alloc(cN * 4);
  d->d_fin = (int*)d->ws.alloc(cN * 4);
  d->d_amax = (unsigned long long*)d->ws.alloc(cN * 8);
  d->tok_buf = (int32_t*)d->ws.alloc((size_t)cN * cL * 4);
  d->chain_done = d->use_chain ? (int*)d->ws.alloc((size_t)2 * gemm_chain_counters((int)cN) * 4) : nullptr;
  for (int i = 0; i < 3; ++i) d->ss[i] = d->fuse_norm ? (float*)d->ws.alloc((size_t)cN * (D / 32) * 4) : nullptr;
  bool ok = d->x && d->h && d->qkv && d->attn && d->qx && d->g && d->logits && d->d_cur && d->d_fin && d->d_amax && d->tok_buf && d->kv_tmp && (!d->use_chain || d->chain_done) &&
            (!d->fuse_norm || (d->ss[0] && d->ss[1] && d->ss[2]));
  d->selfK.assign(c.num_layers, nullptr);
  d->selfV.assign(c.num_layers, nullptr);
  d->crossKV.assign(c.num_layers, nullptr);
  for (int i = 0; ok && i < c.num_layers; ++i) {
    d->selfK[i] = d->ws.alloc(cN * inner * cL * es);
    d->selfV[i] = d->ws.alloc(cN * inner * cL * es);
    if (!latent) d->crossKV[i] = d->ws.alloc(cN * cT * 2 * inner * es);
    ok = d->selfK[i] && d->selfV[i] && (latent || d->crossKV[i]);
  }
  if (!ok) {
    d->ws.release();
    return YMT3_ERR_CUDA;
  }
  d->cap_N = cN; d->cap_T = cT; d->cap_L = cL;
  d->cap_Tp = cTp;
  d->cap_latent = latent;
  return YMT3_OK;
}

// all kernels of ONE decode step; every step-dependent value is read from device memory
int dec_step(ymt3_t5dec* d, int64_t N, int64_t T, int Lmax, int stop_at_eos, int32_t* tokens_out, int n_prefix,
             int latent, cudaStream_t s, int32_t* score_out = nullptr) {
  const ymt3_t5_cfg_t& c = d->c;
  const int D = c.d_model, H = c.num_heads, dk = c.d_kv, inner = H * dk, F = c.d_ff, dt = c.precision;
  const size_t es = dtype_size(dt);
  int rc;
  // PROFILING AID ONLY (tools/time_phases.py): YMT3_DEBUG_SKIP is a bit mask of kernel families to leave out of the
  // step so their in-graph cost can be measured by difference; the tokens are meaningless when it is set.
  // 1 = self-attention, 2 = cross-attention kernel, 4 = all GEMMs of the layers, 8 = norms
  static const int skip = getenv("YMT3_DEBUG_SKIP") ? atoi(getenv("YMT3_DEBUG_SKIP")) : 0;
  // bf16: RMSNorm is fused into the GEMMs around it (GemmParams::norm_ss_in / ss_out): the kernel that writes the
  // residual stream x also emits per-row sum-of-squares partials, the GEMM that consumes norm(x) reads x itself with
  // the norm weight folded into its columns and scales its accumulator rows by rsqrt(mean(x^2) + eps).  24 of the
  // 107 launches of a step disappear.  fp32 keeps the reference order (separate norm kernels).
  const bool fuse = d->fuse_norm;
  const int ch = D / 32;
  auto consume = [&](float* ss) { NormFuse n; n.ss_in = ss; n.chunks = ch; n.eps = c.layer_norm_eps; return n; };
  auto produce = [&](float* ss) { NormFuse n; n.ss_out = ss; return n; };
  // y = W norm(x): separate kernel + GEMM (reference order) or one GEMM on x with the folded weight Wn
  auto normed_linear = [&](const float* ln, const Linear& W, const Linear& Wn, float* ss, void* y, int ny, int act,
                           int gated, float scale, int out_dt) -> int {
    if (fuse) {
      if (skip & 4) return YMT3_OK;
      return linear_fwd(dt, d->x, D, Wn, y, ny, (int)N, act, gated, nullptr, 0, scale, out_dt, s, consume(ss));
    }
    int r = (skip & 8) ? 0 : rmsnorm(d->x, ln, d->h, N, D, c.layer_norm_eps, dt, s);
    if (r || (skip & 4)) return r;
    return linear_fwd(dt, d->h, D, W, y, ny, (int)N, act, gated, nullptr, 0, scale, out_dt, s);
  };
  // x += W a (+ bias), emitting the sum-of-squares partials of the new x for the next norm
  auto residual_linear = [&](const void* a, int lda, const Linear& W, float* ss) -> int {
    if (skip & 4) return YMT3_OK;
    return linear_fwd(dt, a, lda, W, d->x, D, (int)N, 0, 0, d->x, D, 1.f, dt, s, fuse ? produce(ss) : NormFuse());
  };
  float *sA = d->ss[0], *sB = d->ss[1], *sC = d->ss[2];
  if ((rc = embed_pos(d->d_cur, d->embed, d->pos, d->d_step, d->x, (int)N, D, dt, s, fuse ? sA : nullptr))) return rc;
  if (d->use_chain && fuse && !skip) {
    // CHAINED decode step: the GEMMs between two attention kernels run as ONE persistent launch each
    //   chain A = [o-proj (+x) -> cross-q],   chain B = [cross-o (+x) -> wi (gated GELU) -> wo (+x) -> qkv of the next layer]
    // (4 launches per layer instead of 8; dependencies per 128-row tile inside the kernel, gemm_bf16_tc.cu).
    auto gp = [&](const void* A, int lda, const Linear& W, void* C, int ldc, int act, int gated, bool residual,
                  const float* ss_in, float* ss_out) {
      GemmParams p{};
      p.A = A; p.lda = lda; p.W = W.W; p.ldw = W.K; p.C = C; p.ldc = ldc; p.bias = W.bias;
      p.residual = residual ? C : nullptr; p.ldr = ldc;
      p.M = (int)N; p.N = W.N; p.K = W.K; p.act = act; p.gated = gated; p.out_scale = 1.f;
      p.norm_ss_in = ss_in; p.norm_ss_chunks = ss_in ? ch : 0; p.norm_eps = c.layer_norm_eps; p.ss_out = ss_out;
      return p;
    };
    const int nc = gemm_chain_counters((int)d->cap_N);
    if ((rc = linear_fwd(dt, d->x, D, d->layers[0].qkv_n, d->qkv, 3 * inner, (int)N, 0, 0, nullptr, 0, 1.f, dt, s, consume(sA))))
      return rc;
    for (int i = 0; i < c.num_layers; ++i) {
      const T5Layer& L = d->layers[i];
      if ((rc = decode_attention(d->qkv, 3 * inner, (char*)d->qkv + inner * es, (char*)d->qkv + 2 * inner * es, 3 * inner,
                                 d->selfK[i], d->selfV[i], (int64_t)H * d->cap_L * dk, (int64_t)d->cap_L * dk, dk, Lmax + n_prefix,
                                 d->d_step, 0, 1.0f, d->attn, inner, (int)N, H, dk, dt, s, d->rel_bias, d->rel_P)))
        return rc;
      GemmParams ca[2], cb[4];
      const int HZ = H * d->zdim;
      ca[0] = gp(d->attn, inner, L.o, d->x, D, 0, 0, true, nullptr, sB);
      ca[1] = latent ? gp(d->x, D, L.xq_abs_n, d->qz, HZ, 0, 0, false, sB, nullptr)
                     : gp(d->x, D, L.xq_n, d->qx, inner, 0, 0, false, sB, nullptr);
      if ((rc = gemm_chain_bf16(ca, 2, d->chain_done, d->d_step, c.num_layers, i, s))) return rc;
      if (latent) {
        if ((rc = cross_attn_absorbed(d->qz, HZ, d->zbuf, d->cz, HZ, N, H, (int)T, (int)((T + 15) / 16 * 16), d->zdim, s))) return rc;
        cb[0] = gp(d->cz, HZ, L.xo_abs, d->x, D, 0, 0, true, nullptr, sC);
      } else {
        if ((rc = decode_attention(d->qx, inner, nullptr, nullptr, 0, d->crossKV[i],
                                   (char*)d->crossKV[i] + (size_t)N * inner * T * es, (int64_t)H * T * dk, T * dk, dk, 0,
                                   d->d_step, (int)T, 1.0f, d->attn, inner, (int)N, H, dk, dt, s)))
          return rc;
        cb[0] = gp(d->attn, inner, L.xo, d->x, D, 0, 0, true, nullptr, sC);
      }
      cb[1] = gp(d->x, D, L.wi_n, d->g, F, YMT3_ACT_GELU_NEW, 1, false, sC, nullptr);
      cb[2] = gp(d->g, F, L.wo, d->x, D, 0, 0, true, nullptr, sA);
      int nb = 3;
      if (i + 1 < c.num_layers) cb[nb++] = gp(d->x, D, d->layers[i + 1].qkv_n, d->qkv, 3 * inner, 0, 0, false, sA, nullptr);
      if ((rc = gemm_chain_bf16(cb, nb, d->chain_done + nc, d->d_step, c.num_layers, i, s))) return rc;
    }
  } else
  for (int i = 0; i < c.num_layers; ++i) {
    const T5Layer& L = d->layers[i];
    // self-attention over the device-resident cache (modeling_t5.py:269-305, 356-377)
    if ((rc = normed_linear(L.ln_sa, L.qkv, L.qkv_n, sA, d->qkv, 3 * inner, 0, 0, 1.f, dt))) return rc;
    if (!(skip & 1) && (rc = decode_attention(d->qkv, 3 * inner, (char*)d->qkv + inner * es, (char*)d->qkv + 2 * inner * es, 3 * inner,
                               d->selfK[i], d->selfV[i], (int64_t)H * d->cap_L * dk, (int64_t)d->cap_L * dk, dk, Lmax + n_prefix,
                               d->d_step, 0, 1.0f, d->attn, inner, (int)N, H, dk, dt, s, d->rel_bias, d->rel_P)))
      return rc;
    if ((rc = residual_linear(d->attn, inner, L.o, sB))) return rc;
    // cross-attention (modeling_t5.py:387-408)
    if (latent) {
      // absorbed form: latent-space query, attention over the shared latent tile, folded (Wo Wv Wp) output
      const int HZ = H * d->zdim;
      if ((rc = normed_linear(L.ln_ca, L.xq_abs, L.xq_abs_n, sB, d->qz, HZ, 0, 0, 1.f, dt))) return rc;
      if (!(skip & 2) && (rc = cross_attn_absorbed(d->qz, HZ, d->zbuf, d->cz, HZ, N, H, (int)T, (int)((T + 15) / 16 * 16), d->zdim, s))) return rc;
      if ((rc = residual_linear(d->cz, HZ, L.xo_abs, sC))) return rc;
    } else {
      // over encoder K/V computed once
      if ((rc = normed_linear(L.ln_ca, L.xq, L.xq_n, sB, d->qx, inner, 0, 0, 1.f, dt))) return rc;
      if (!(skip & 2) && (rc = decode_attention(d->qx, inner, nullptr, nullptr, 0, d->crossKV[i],
                                 (char*)d->crossKV[i] + (size_t)N * inner * T * es, (int64_t)H * T * dk, T * dk, dk, 0,
                                 d->d_step, (int)T, 1.0f, d->attn, inner, (int)N, H, dk, dt, s)))
        return rc;
      if ((rc = residual_linear(d->attn, inner, L.xo, sC))) return rc;
    }
    // gated-GELU feed-forward
    if ((rc = normed_linear(L.ln_ff, L.wi, L.wi_n, sC, d->g, F, YMT3_ACT_GELU_NEW, 1, 1.f, dt))) return rc;
    if ((rc = residual_linear(d->g, F, L.wo, sA))) return rc;
  }
  // final norm + LM head; tied embeddings scale hidden by d_model^-0.5 (modeling_t5.py:1105-1110)
  const float sc = c.tie_word_embeddings ? 1.0f / sqrtf((float)D) : 1.0f;
  // greedy selection fused into the vocab-projection epilogue: every epilogue thread reduces its logits to one packed
  // (value, column) key and atomicMax-es it into d_amax[row]; select_advance turns keys into tokens and ends the step.
  // (A/B against the separate arg-max kernel + step-advance kernel: neutral, profiles/r01_ab_fused_select.txt.)
  NormFuse nf = fuse ? consume(sA) : NormFuse();
  nf.argmax_out = d->d_amax; nf.argmax_n = c.vocab_size;
  if (fuse) {
    if ((rc = linear_fwd(dt, d->x, D, d->lm_head_n, d->logits, d->Vp, (int)N, 0, 0, nullptr, 0, sc, YMT3_F32, s, nf))) return rc;
  } else {
    if ((rc = rmsnorm(d->x, d->final_ln, d->h, N, D, c.layer_norm_eps, dt, s))) return rc;
    if ((rc = linear_fwd(dt, d->h, D, d->lm_head, d->logits, d->Vp, (int)N, 0, 0, nullptr, 0, sc, YMT3_F32, s, nf))) return rc;
  }
  return select_advance(d->d_amax, (int)N, d->d_step, d->d_cur, d->d_fin, tokens_out, Lmax, c.eos_id, c.pad_id,
                        stop_at_eos, d->d_unfinished, n_prefix ? d->d_forced : nullptr, n_prefix, s, score_out);
}

}  // namespace

extern "C" int ymt3_t5dec_generate(ymt3_t5dec_t* d, const void* enc_hs, int64_t N, int64_t T, int32_t max_len,
                                   int32_t stop_at_eos, int32_t early_stop_interval, int32_t* tokens_out, void* stream) {
  return ymt3_t5dec_generate_prefixed(d, enc_hs, N, T, nullptr, 0, max_len, stop_at_eos, early_stop_interval, tokens_out,
                                      stream);
}

namespace {
struct ScoreOpts {
  int32_t* argmax_out = nullptr;        // (N, P): the model's arg-max at every teacher-forced step
  const int32_t* logit_steps = nullptr; // HOST array of step indices whose logits are copied out
  int n_logit_steps = 0;
  float* logits_out = nullptr;          // (n_logit_steps, N, V) fp32 device
};
int generate_impl(ymt3_t5dec_t* d, const void* enc_hs, int64_t N, int64_t T, int latent_channels,
                  const int32_t* prefix_ids, int32_t P, int32_t max_len, int32_t stop_at_eos,
                  int32_t early_stop_interval, int32_t* tokens_out, void* stream, const ScoreOpts* score = nullptr);
}

extern "C" int ymt3_t5dec_generate_prefixed(ymt3_t5dec_t* d, const void* enc_hs, int64_t N, int64_t T,
                                            const int32_t* prefix_ids, int32_t P, int32_t max_len, int32_t stop_at_eos,
                                            int32_t early_stop_interval, int32_t* tokens_out, void* stream) {
  return generate_impl(d, enc_hs, N, T, 0, prefix_ids, P, max_len, stop_at_eos, early_stop_interval, tokens_out, stream);
}

extern "C" int ymt3_t5dec_generate_latent(ymt3_t5dec_t* d, const void* latents, int64_t B, int64_t T, int32_t C,
                                          const int32_t* prefix_ids, int32_t P, int32_t max_len, int32_t stop_at_eos,
                                          int32_t early_stop_interval, int32_t* tokens_out, void* stream) {
  YMT3_REQUIRE(d && d->zdim > 0, "t5dec_generate_latent: handle was created without absorbed cross-attention weights");
  YMT3_REQUIRE(C >= 1, "t5dec_generate_latent: bad channel count %d", C);
  return generate_impl(d, latents, B * C, T, C, prefix_ids, P, max_len, stop_at_eos, early_stop_interval, tokens_out,
                       stream);
}

namespace {
// latent_channels == 0: enc_hs is (N, T, d_model), cross K/V per layer.  > 0: enc_hs is the encoder latent array
// (N / C, T, C, zdim) and the cross-attention runs in its absorbed form.
int generate_impl(ymt3_t5dec_t* d, const void* enc_hs, int64_t N, int64_t T, int latent_channels,
                  const int32_t* prefix_ids, int32_t P, int32_t max_len, int32_t stop_at_eos,
                  int32_t early_stop_interval, int32_t* tokens_out, void* stream, const ScoreOpts* score) {
  const int latent = latent_channels > 0;
  const bool scoring = score != nullptr;   // teacher-forced scoring: every step forced, nothing generated (max_len 0)
  YMT3_REQUIRE(d && (tokens_out || scoring), "t5dec_generate: null argument");
  if (N <= 0) return YMT3_OK;
  YMT3_REQUIRE(enc_hs && T > 0, "t5dec_generate: bad encoder states");
  YMT3_REQUIRE(P >= 0 && (P == 0 || prefix_ids), "t5dec_generate: bad task prefix");
  YMT3_REQUIRE((max_len > 0 || (scoring && P > 0)) && max_len + P <= d->c.max_length,
               "t5dec_generate: max_len %d + prefix %d outside (0, %d]", max_len, P, d->c.max_length);
  YMT3_REQUIRE(!d->pos || max_len + P <= d->n_pos, "t5dec_generate: position table too short");
  const ymt3_t5_cfg_t& c = d->c;
  const int D = c.d_model, inner = c.num_heads * c.d_kv, dt = c.precision;
  cudaStream_t caller = (cudaStream_t)stream;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  const bool caller_capturing = cudaStreamIsCapturing(caller, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
  // run on the internal stream unless the caller is itself capturing (then stay in its capture)
  cudaStream_t s = caller_capturing ? caller : d->own_stream;
  int rc;
  if (!caller_capturing) {
    YMT3_CUDA_CHECK(cudaEventRecord(d->ev_in, caller));
    YMT3_CUDA_CHECK(cudaStreamWaitEvent(s, d->ev_in, 0));
  }
  if ((rc = dec_ensure(d, N, T, max_len + P, latent, s))) return rc;
  if (P > 0) {
    if (N * P > d->cap_P) {
      YMT3_CUDA_CHECK(cudaStreamSynchronize(s));
      d->d_forced = (int*)d->ws.alloc((size_t)N * P * 4);
      if (!d->d_forced) return YMT3_ERR_CUDA;
      d->cap_P = N * P;
      if (d->graph) { cudaGraphExecDestroy(d->graph); d->graph = nullptr; d->graph_N = -1; }
    }
    YMT3_CUDA_CHECK(cudaMemcpyAsync(d->d_forced, prefix_ids, (size_t)N * P * 4, cudaMemcpyDeviceToDevice, s));
  }
  // state init (device side)
  if ((rc = fill_i32(d->d_step, 0, 8, s))) return rc;  // step + unfinished[2] (+pad)
  if ((rc = fill_i32(d->d_cur, c.start_id, N, s))) return rc;
  if ((rc = fill_i32(d->d_fin, 0, N, s))) return rc;
  YMT3_CUDA_CHECK(cudaMemsetAsync(d->d_amax, 0, (size_t)N * 8, s));
  if (d->chain_done) YMT3_CUDA_CHECK(cudaMemsetAsync(d->chain_done, 0, (size_t)2 * gemm_chain_counters((int)d->cap_N) * 4, s));
  if ((rc = fill_i32(d->tok_buf, c.pad_id, N * max_len, s))) return rc;
  int32_t* score_buf = nullptr;
  if (scoring && score->argmax_out) {
    // (N, P) arg-max record of the forced steps shares the token buffer's tail (cap_N * cap_L >= N * (P + max_len))
    score_buf = d->tok_buf + N * max_len;
  }
  if (latent) {
    // the only per-call encoder-side work: latents regrouped per sequence (channel-major), time padded to 16
    const int64_t Tp = (T + 15) / 16 * 16;
    if (Tp != T) YMT3_CUDA_CHECK(cudaMemsetAsync(d->zbuf, 0, (size_t)N * Tp * d->zdim * dtype_size(dt), s));
    if ((rc = gather_latents(enc_hs, d->zbuf, N / latent_channels, (int)T, latent_channels, (int)Tp, d->zdim, s))) return rc;
  }
  // cross-attention K/V of every layer, once (modeling_t5.py:287-299)
  for (int i = 0; !latent && i < c.num_layers; ++i) {
    if ((rc = linear_fwd(dt, enc_hs, D, d->layers[i].xkv, d->kv_tmp, 2 * inner, (int)(N * T), 0, 0, nullptr, 0, 1.f, dt,
                         s)))
      return rc;
    // head-major contiguous layout: each (sequence, head) becomes two contiguous T*dk streams for the 256+ re-reads
    if ((rc = split_kv_heads(d->kv_tmp, d->crossKV[i], (char*)d->crossKV[i] + (size_t)N * inner * T * dtype_size(dt), N,
                             (int)T, c.num_heads, c.d_kv, dt, s)))
      return rc;
  }

  // one decode step captured into a CUDA graph (all step-dependent scalars live on the device)
  const bool use_graph = getenv("YMT3_NO_GRAPH") == nullptr && !caller_capturing;
  if (use_graph && (!d->graph || d->graph_N != N || d->graph_T != T || d->graph_L != max_len ||
                    d->graph_score != (score_buf != nullptr) || d->graph_stop != stop_at_eos || d->graph_prefix != P ||
                    d->graph_latent != latent)) {
    if (d->graph) {
      cudaGraphExecDestroy(d->graph);
      d->graph = nullptr;
    }
    cudaGraph_t g = nullptr;
    YMT3_CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    rc = dec_step(d, N, T, max_len, stop_at_eos, d->tok_buf, P, latent, s, score_buf);
    cudaError_t ce = cudaStreamEndCapture(s, &g);
    if (rc) {
      if (g) cudaGraphDestroy(g);
      return rc;
    }
    YMT3_CUDA_CHECK(ce);
    ce = cudaGraphInstantiate(&d->graph, g, 0);
    cudaGraphDestroy(g);
    YMT3_CUDA_CHECK(ce);
    d->graph_N = N; d->graph_T = T; d->graph_L = max_len;
    d->graph_score = score_buf != nullptr;
    d->graph_stop = stop_at_eos;
    d->graph_prefix = P;
    d->graph_latent = latent;
  }
  for (int t = 0; t < max_len + P; ++t) {
    if (use_graph) {
      YMT3_CUDA_CHECK(cudaGraphLaunch(d->graph, s));
    } else if ((rc = dec_step(d, N, T, max_len, stop_at_eos, d->tok_buf, P, latent, s, score_buf))) {
      return rc;
    }
    for (int k = 0; scoring && k < score->n_logit_steps; ++k) {
      if (score->logit_steps[k] != t) continue;
      YMT3_CUDA_CHECK(cudaMemcpy2DAsync(score->logits_out + (size_t)k * N * c.vocab_size, (size_t)c.vocab_size * 4, d->logits,
                                        (size_t)d->Vp * 4, (size_t)c.vocab_size * 4, (size_t)N, cudaMemcpyDeviceToDevice, s));
    }
    if (stop_at_eos && early_stop_interval > 0 && (t + 1) % early_stop_interval == 0 && t + 1 < max_len + P) {
      // rows still unfinished after step t were counted into slot (t & 1)
      YMT3_CUDA_CHECK(cudaMemcpyAsync(d->h_unfinished, d->d_unfinished + (t & 1), 4, cudaMemcpyDeviceToHost, s));
      YMT3_CUDA_CHECK(cudaStreamSynchronize(s));
      if (*d->h_unfinished == 0) break;
    }
  }
  if (tokens_out && max_len > 0)
    YMT3_CUDA_CHECK(cudaMemcpyAsync(tokens_out, d->tok_buf, (size_t)N * max_len * 4, cudaMemcpyDeviceToDevice, s));
  if (score_buf)
    YMT3_CUDA_CHECK(cudaMemcpyAsync(score->argmax_out, score_buf, (size_t)N * P * 4, cudaMemcpyDeviceToDevice, s));
  if (!caller_capturing) {
    YMT3_CUDA_CHECK(cudaEventRecord(d->ev_out, s));
    YMT3_CUDA_CHECK(cudaStreamWaitEvent(caller, d->ev_out, 0));
  }
  return YMT3_OK;
}
}  // namespace

extern "C" int ymt3_t5dec_score_forced(ymt3_t5dec_t* d, const void* enc, int64_t B, int64_t T, int32_t C,
                                       const int32_t* forced_ids, int32_t L, int32_t* argmax_out,
                                       const int32_t* logit_steps_host, int32_t n_logit_steps, float* logits_out,
                                       void* stream) {
  YMT3_REQUIRE(d && forced_ids && L > 0, "t5dec_score_forced: null argument");
  YMT3_REQUIRE(C >= 0 && (C == 0 || d->zdim > 0), "t5dec_score_forced: latent mode needs absorbed cross-attention weights");
  YMT3_REQUIRE(n_logit_steps >= 0 && (n_logit_steps == 0 || (logit_steps_host && logits_out)),
               "t5dec_score_forced: bad logit capture arguments");
  for (int k = 0; k < n_logit_steps; ++k)
    YMT3_REQUIRE(logit_steps_host[k] >= 0 && logit_steps_host[k] < L, "t5dec_score_forced: logit step %d outside [0, %d)",
                 logit_steps_host[k], L);
  ScoreOpts so;
  so.argmax_out = argmax_out; so.logit_steps = logit_steps_host; so.n_logit_steps = n_logit_steps; so.logits_out = logits_out;
  return generate_impl(d, enc, C > 0 ? B * C : B, T, C, forced_ids, L, 0, 0, 0, nullptr, stream, &so);
}

extern "C" int ymt3_t5dec_last_logits(ymt3_t5dec_t* d, float* out, int64_t N, void* stream) {
  YMT3_REQUIRE(d && out && N <= d->cap_N, "t5dec_last_logits: bad argument");
  YMT3_CUDA_CHECK(cudaMemcpy2DAsync(out, (size_t)d->c.vocab_size * 4, d->logits, (size_t)d->Vp * 4,
                                    (size_t)d->c.vocab_size * 4, (size_t)N, cudaMemcpyDeviceToDevice,
                                    (cudaStream_t)stream));
  return YMT3_OK;
}
