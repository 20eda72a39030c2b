// Perceiver-TF encoder (YPTF) orchestrated natively over the shared kernels.  Semantics restated
// in oracle/perceiver_tf.py (layer arithmetic = HF PerceiverLayer / Mixtral MoE, cited there).
// No tensor is ever transposed: the temporal transformer addresses the (B, T, K, D) layout through
// two-level batch strides in the attention kernel, and all linears are row-wise.
#include "model_common.cuh"
#include "moe.cuh"
#include <math.h>
#include <stdlib.h>

using namespace ymt3;

namespace {

struct PLayer {
  bool cross = false;
  float *ln1_w = nullptr, *ln1_b = nullptr, *ln2_w = nullptr, *ln2_b = nullptr, *ln3_w = nullptr, *ln3_b = nullptr;
  Linear q, kv, qkv, o, d1, d2;
  MoEWeights moe;
};

}  // namespace

struct ymt3_ptf {
  ymt3_ptf_cfg_t c;
  DevicePool weights, ws;
  std::vector<PLayer> layers;          // per block: sca, local x N, temporal x M
  void* latents = nullptr;             // (K, D) compute dtype
  void* latent_pos = nullptr;          // (K, D) or null
  void* temporal_pos = nullptr;        // (max_time, D) or null
  float *fin_w = nullptr, *fin_b = nullptr;
  float *rope_cos_k = nullptr, *rope_sin_k = nullptr;   // (K, rot/2)
  float *rope_cos_t = nullptr, *rope_sin_t = nullptr;   // (max_time, rot/2)
  // workspace
  int64_t cap_rows = 0, cap_kv = 0;
  void *h = nullptr, *a = nullptr, *y = nullptr, *qkvb = nullptr, *ctx = nullptr, *mid = nullptr, *kvn = nullptr,
       *kvb = nullptr, *moe_ws = nullptr;
};

namespace {

int load_norm(const ymt3_ptf_cfg_t& c, const TensorTable& tt, DevicePool& pool, const std::string& name, int dim,
              float** w, float** b) {
  int rc = pack_vec(pool, {tt.require(name + ".weight", dim)}, false, w, 0);
  if (rc) return rc;
  *b = nullptr;
  if (c.norm_type == 0) rc = pack_vec(pool, {tt.require(name + ".bias", dim)}, false, b, 0);
  return rc;
}

int load_layer(const ymt3_ptf_cfg_t& c, const TensorTable& tt, DevicePool& pool, const std::string& p, bool cross,
               PLayer& L) {
  const int D = c.d_latent, C = c.kv_dim, dt = c.precision;
  const std::string a = p + "attention.self.";
  L.cross = cross;
  int rc;
  if ((rc = load_norm(c, tt, pool, a + "layernorm1", D, &L.ln1_w, &L.ln1_b))) return rc;
  if (cross) {
    if ((rc = load_norm(c, tt, pool, a + "layernorm2", C, &L.ln2_w, &L.ln2_b))) return rc;
    if ((rc = pack_rows(pool, {tt.require(a + "query.weight", D, D)}, D, dt, false, &L.q, 0))) return rc;
    if ((rc = pack_vec(pool, {tt.require(a + "query.bias", D)}, false, &L.q.bias, 0))) return rc;
    if ((rc = pack_rows(pool, {tt.require(a + "key.weight", D, C), tt.require(a + "value.weight", D, C)}, C, dt, false,
                        &L.kv, 0))) return rc;
    if ((rc = pack_vec(pool, {tt.require(a + "key.bias", D), tt.require(a + "value.bias", D)}, false, &L.kv.bias, 0)))
      return rc;
  } else {
    if ((rc = pack_rows(pool, {tt.require(a + "query.weight", D, D), tt.require(a + "key.weight", D, D),
                               tt.require(a + "value.weight", D, D)}, D, dt, false, &L.qkv, 0))) return rc;
    if ((rc = pack_vec(pool, {tt.require(a + "query.bias", D), tt.require(a + "key.bias", D),
                              tt.require(a + "value.bias", D)}, false, &L.qkv.bias, 0))) return rc;
  }
  if ((rc = pack_rows(pool, {tt.require(p + "attention.output.dense.weight", D, D)}, D, dt, false, &L.o, 0))) return rc;
  if ((rc = pack_vec(pool, {tt.require(p + "attention.output.dense.bias", D)}, false, &L.o.bias, 0))) return rc;
  if ((rc = load_norm(c, tt, pool, p + "layernorm", D, &L.ln3_w, &L.ln3_b))) return rc;
  const int I = c.ff_widening * D;
  if (c.ff_type == 0) {
    if ((rc = pack_rows(pool, {tt.require(p + "mlp.dense1.weight", I, D)}, D, dt, false, &L.d1, 0))) return rc;
    if ((rc = pack_vec(pool, {tt.require(p + "mlp.dense1.bias", I)}, false, &L.d1.bias, 0))) return rc;
    if ((rc = pack_rows(pool, {tt.require(p + "mlp.dense2.weight", D, I)}, I, dt, false, &L.d2, 0))) return rc;
    if ((rc = pack_vec(pool, {tt.require(p + "mlp.dense2.bias", D)}, false, &L.d2.bias, 0))) return rc;
  } else {
    const int E = c.moe_experts;
    float* gate = nullptr;
    const ymt3_tensor_t* g = tt.require(p + "moe.gate.weight", E, D);
    if (!g) return YMT3_ERR_INVALID;
    gate = (float*)pool.alloc((size_t)E * D * 4);
    if (!gate) return YMT3_ERR_CUDA;
    if ((rc = pack_rows_at(gate, 0, 1, g, D, YMT3_F32, 0))) return rc;
    void* w13 = pool.alloc((size_t)E * 2 * I * D * dtype_size(dt));
    void* w2 = pool.alloc((size_t)E * D * I * dtype_size(dt));
    if (!w13 || !w2) return YMT3_ERR_CUDA;
    for (int e = 0; e < E; ++e) {
      const std::string ep = p + "moe.experts." + std::to_string(e) + ".";
      if ((rc = pack_rows_at(w13, (int64_t)e * 2 * I, 2, tt.require(ep + "w1.weight", I, D), D, dt, 0))) return rc;
      if ((rc = pack_rows_at(w13, (int64_t)e * 2 * I + 1, 2, tt.require(ep + "w3.weight", I, D), D, dt, 0))) return rc;
      if ((rc = pack_rows_at(w2, (int64_t)e * D, 1, tt.require(ep + "w2.weight", D, I), I, dt, 0))) return rc;
    }
    L.moe.gate = gate; L.moe.w13 = w13; L.moe.w2 = w2;
    L.moe.D = D; L.moe.I = I; L.moe.E = E; L.moe.topk = c.moe_topk; L.moe.act = c.act;
  }
  return YMT3_OK;
}

int make_rope_tables(DevicePool& pool, int n_pos, int rot, float** cos_d, float** sin_d) {
  const int half = rot / 2;
  std::vector<float> hc((size_t)n_pos * half), hs((size_t)n_pos * half);
  for (int p = 0; p < n_pos; ++p)
    for (int i = 0; i < half; ++i) {
      const double inv = 1.0 / pow(10000.0, (double)(2 * i) / (double)rot);
      hc[(size_t)p * half + i] = (float)cos((double)p * inv);
      hs[(size_t)p * half + i] = (float)sin((double)p * inv);
    }
  *cos_d = (float*)pool.alloc(hc.size() * 4);
  *sin_d = (float*)pool.alloc(hs.size() * 4);
  if (!*cos_d || !*sin_d) return YMT3_ERR_CUDA;
  YMT3_CUDA_CHECK(cudaMemcpy(*cos_d, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice));
  YMT3_CUDA_CHECK(cudaMemcpy(*sin_d, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice));
  return YMT3_OK;
}

int norm_fwd(const ymt3_ptf_cfg_t& c, const void* x, const float* w, const float* b, void* y, int64_t rows, int dim,
             cudaStream_t s) {
  return c.norm_type == 1 ? rmsnorm(x, w, y, rows, dim, c.norm_eps, c.precision, s)
                          : layernorm(x, w, b, y, rows, dim, c.norm_eps, c.precision, s);
}

}  // namespace

extern "C" int ymt3_ptf_create(const ymt3_ptf_cfg_t* cfg, const ymt3_tensor_t* tensors, int n, ymt3_ptf_t** out) {
  YMT3_REQUIRE(cfg && tensors && out, "ptf_create: null argument");
  const ymt3_ptf_cfg_t& c = *cfg;
  YMT3_REQUIRE(c.precision == YMT3_F32 || c.precision == YMT3_BF16, "ptf_create: bad precision");
  YMT3_REQUIRE(c.num_latents > 0 && c.d_latent > 0 && c.d_latent % 8 == 0 && c.kv_dim > 0 && c.kv_dim % 8 == 0 &&
                   c.num_blocks > 0 && c.num_local >= 0 && c.num_temporal >= 0 && c.max_time > 0,
               "ptf_create: bad dimensions");
  YMT3_REQUIRE(c.cross_heads > 0 && c.self_heads > 0 && c.d_latent % c.cross_heads == 0 && c.d_latent % c.self_heads == 0,
               "ptf_create: heads must divide d_latent");
  const int dhx = c.d_latent / c.cross_heads, dhs = c.d_latent / c.self_heads;
  auto okdh = [](int d) { return d == 16 || d == 32 || d == 64 || d == 128; };
  YMT3_REQUIRE(okdh(dhx) && okdh(dhs), "ptf_create: head dims must be in {16,32,64,128} (got %d, %d)", dhx, dhs);
  YMT3_REQUIRE(c.ff_type == 0 || (c.moe_experts >= 1 && c.moe_experts <= 32 && c.moe_topk >= 1 && c.moe_topk <= 4),
               "ptf_create: bad MoE config");
  YMT3_REQUIRE(c.pos_type != 2 || (c.rope_dim > 0 && c.rope_dim % 2 == 0 && c.rope_dim <= dhs), "ptf_create: bad rope_dim");
  ymt3_ptf* h = new ymt3_ptf();
  h->c = c;
  TensorTable tt{tensors, n};
  const int K = c.num_latents, D = c.d_latent, dt = c.precision;
  int rc = YMT3_OK;
  const ymt3_tensor_t* lat = tt.require("latent_array.latents", K, D);
  if (!lat) rc = YMT3_ERR_INVALID;
  if (!rc) rc = pack_table(h->weights, (const float*)lat->data, false, (int64_t)K * D, dt, &h->latents, 0);
  if (!rc && c.pos_type == 1) {
    const ymt3_tensor_t *lp = tt.require("latent_pos_emb", K, D), *tp = tt.require("temporal_pos_emb", c.max_time, D);
    if (!lp || !tp) rc = YMT3_ERR_INVALID;
    if (!rc) rc = pack_table(h->weights, (const float*)lp->data, false, (int64_t)K * D, dt, &h->latent_pos, 0);
    if (!rc) rc = pack_table(h->weights, (const float*)tp->data, false, (int64_t)c.max_time * D, dt, &h->temporal_pos, 0);
  }
  if (!rc && c.pos_type == 2) {
    rc = make_rope_tables(h->weights, K, c.rope_dim, &h->rope_cos_k, &h->rope_sin_k);
    if (!rc) rc = make_rope_tables(h->weights, c.max_time, c.rope_dim, &h->rope_cos_t, &h->rope_sin_t);
  }
  for (int b = 0; b < c.num_blocks && !rc; ++b) {
    const std::string bp = "block." + std::to_string(b) + ".";
    h->layers.emplace_back();
    rc = load_layer(c, tt, h->weights, bp + "sca.", true, h->layers.back());
    for (int i = 0; i < c.num_local && !rc; ++i) {
      h->layers.emplace_back();
      rc = load_layer(c, tt, h->weights, bp + "local." + std::to_string(i) + ".", false, h->layers.back());
    }
    for (int i = 0; i < c.num_temporal && !rc; ++i) {
      h->layers.emplace_back();
      rc = load_layer(c, tt, h->weights, bp + "temporal." + std::to_string(i) + ".", false, h->layers.back());
    }
  }
  if (!rc) rc = load_norm(c, tt, h->weights, "layernorm", D, &h->fin_w, &h->fin_b);
  if (!rc && cudaDeviceSynchronize() != cudaSuccess) {
    ymt3_set_error("ptf_create: weight packing failed: %s", cudaGetErrorString(cudaGetLastError()));
    rc = YMT3_ERR_CUDA;
  }
  if (rc) {
    h->weights.release();
    delete h;
    return rc;
  }
  *out = h;
  return YMT3_OK;
}

extern "C" int ymt3_ptf_destroy(ymt3_ptf_t* h) {
  if (!h) return YMT3_OK;
  h->weights.release();
  h->ws.release();
  delete h;
  return YMT3_OK;
}

namespace {

// attention + output dense (+ query residual) + norm + feed-forward (+ residual) of one Perceiver layer.
// mode: 0 = spectral cross-attention, 1 = latent self-attention, 2 = temporal self-attention
int layer_fwd(ymt3_ptf* h, const PLayer& L, int mode, const void* x_kv, int64_t B, int64_t T, int64_t Fp, bool query_res,
              cudaStream_t s) {
  const ymt3_ptf_cfg_t& c = h->c;
  const int K = c.num_latents, D = c.d_latent, C = c.kv_dim, dt = c.precision;
  const size_t es = dtype_size(dt);
  const int64_t rows = B * T * K, kvrows = B * T * Fp;
  int rc;
  if ((rc = norm_fwd(c, h->h, L.ln1_w, L.ln1_b, h->y, rows, D, s))) return rc;
  AttnParams a{};
  a.O = h->ctx;
  if (mode == 0) {
    const int H = c.cross_heads, dh = D / H;
    if ((rc = norm_fwd(c, x_kv, L.ln2_w, L.ln2_b, h->kvn, kvrows, C, s))) return rc;
    if ((rc = linear_fwd(dt, h->y, D, L.q, h->qkvb, D, (int)rows, 0, 0, nullptr, 0, 1.f, dt, s))) return rc;
    if ((rc = linear_fwd(dt, h->kvn, C, L.kv, h->kvb, 2 * D, (int)kvrows, 0, 0, nullptr, 0, 1.f, dt, s))) return rc;
    a.Q = h->qkvb; a.q_sb = (int64_t)K * D; a.q_sh = dh; a.q_ss = D;
    a.K = h->kvb; a.V = (char*)h->kvb + (size_t)D * es;
    a.k_sb = a.v_sb = Fp * 2 * D; a.k_sh = a.v_sh = dh; a.k_ss = a.v_ss = 2 * D;
    a.o_sb = (int64_t)K * D; a.o_sh = dh; a.o_ss = D;
    a.B = (int)(B * T); a.H = H; a.Sq = K; a.Sk = (int)Fp; a.dk = dh;
    a.scale = 1.0f / sqrtf((float)dh);
  } else {
    const int H = c.self_heads, dh = D / H;
    if ((rc = linear_fwd(dt, h->y, D, L.qkv, h->qkvb, 3 * D, (int)rows, 0, 0, nullptr, 0, 1.f, dt, s))) return rc;
    if (c.pos_type == 2) {
      const float* ct = mode == 1 ? h->rope_cos_k : h->rope_cos_t;
      const float* st = mode == 1 ? h->rope_sin_k : h->rope_sin_t;
      const int S = mode == 1 ? K : (int)T;
      // fused into the tiny-sequence attention kernel (rotation applied while K is staged in shared memory and
      // q sits in registers) when that kernel is applicable; otherwise a separate in-place pass
      const int G = S <= 32 ? 32 : (S <= 64 ? 64 : 128);
      const bool small_ok = (dh == 16 || dh == 32) && S <= 128 && (size_t)(128 / G) * 2 * S * dh * 4 <= 48 * 1024;
      if (small_ok) {
        a.rope_cos = ct; a.rope_sin = st; a.rope_dim = c.rope_dim;
      } else {
        const int64_t div = mode == 1 ? 1 : K;
        // q heads then k heads are contiguous in the first 2*D columns: treat them as 2*H heads
        if ((rc = rope_inplace(h->qkvb, rows, 3 * D, 0, 2 * H, dh, c.rope_dim, div, S, ct, st, dt, s))) return rc;
      }
    }
    a.Q = h->qkvb; a.K = (char*)h->qkvb + (size_t)D * es; a.V = (char*)h->qkvb + (size_t)2 * D * es;
    a.q_sh = a.k_sh = a.v_sh = dh; a.o_sh = dh;
    if (mode == 1) {
      a.q_sb = a.k_sb = a.v_sb = (int64_t)K * 3 * D; a.q_ss = a.k_ss = a.v_ss = 3 * D;
      a.o_sb = (int64_t)K * D; a.o_ss = D;
      a.B = (int)(B * T); a.Sq = a.Sk = K;
    } else {
      a.inner = K;
      a.q_sb = a.k_sb = a.v_sb = T * K * 3 * D; a.q_sb2 = a.k_sb2 = a.v_sb2 = 3 * D;
      a.q_ss = a.k_ss = a.v_ss = (int64_t)K * 3 * D;
      a.o_sb = T * K * D; a.o_sb2 = D; a.o_ss = (int64_t)K * D;
      a.B = (int)(B * K); a.Sq = a.Sk = (int)T;
    }
    a.H = H; a.dk = dh;
    a.scale = 1.0f / sqrtf((float)dh);
  }
  if ((rc = attention(a, dt, s))) return rc;
  // output dense (+ query residual)  (modeling_perceiver.py:322-328)
  if ((rc = linear_fwd(dt, h->ctx, D, L.o, h->a, D, (int)rows, 0, 0, query_res ? h->h : nullptr, D, 1.f, dt, s))) return rc;
  // layer_out = ff(norm(attn_out)) + attn_out  (:401-411)
  if ((rc = norm_fwd(c, h->a, L.ln3_w, L.ln3_b, h->y, rows, D, s))) return rc;
  if (c.ff_type == 0) {
    const int I = c.ff_widening * D;
    if ((rc = linear_fwd(dt, h->y, D, L.d1, h->mid, I, (int)rows, c.act, 0, nullptr, 0, 1.f, dt, s))) return rc;
    return linear_fwd(dt, h->mid, I, L.d2, h->h, D, (int)rows, 0, 0, h->a, D, 1.f, dt, s);
  }
  return moe_forward(dt, h->y, h->a, h->h, rows, L.moe, h->moe_ws, s);
}

}  // namespace

extern "C" int ymt3_ptf_forward(ymt3_ptf_t* h, const void* x, int64_t B, int64_t T, int64_t Fp, void* out, void* stream) {
  YMT3_REQUIRE(h && out, "ptf_forward: null argument");
  if (B <= 0 || T <= 0) return YMT3_OK;
  YMT3_REQUIRE(x && Fp > 0, "ptf_forward: bad input");
  const ymt3_ptf_cfg_t& c = h->c;
  YMT3_REQUIRE(T <= c.max_time, "ptf_forward: T=%lld exceeds max_time=%d", (long long)T, c.max_time);
  const int K = c.num_latents, D = c.d_latent, C = c.kv_dim, dt = c.precision;
  const int64_t rows = B * T * K, kvrows = B * T * Fp;
  YMT3_REQUIRE(rows < (1ll << 31) && kvrows < (1ll << 31), "ptf_forward: batch too large");
  cudaStream_t s = (cudaStream_t)stream;
  if (rows > h->cap_rows || kvrows > h->cap_kv) {
    YMT3_CUDA_CHECK(cudaStreamSynchronize(s));
    h->ws.release();
    h->cap_rows = h->cap_kv = 0;
    const size_t es = dtype_size(dt);
    const int I = c.ff_widening * D;
    h->h = h->ws.alloc(rows * D * es);
    h->a = h->ws.alloc(rows * D * es);
    h->y = h->ws.alloc(rows * D * es);
    h->qkvb = h->ws.alloc(rows * 3 * D * es);
    h->ctx = h->ws.alloc(rows * D * es);
    h->kvn = h->ws.alloc(kvrows * C * es);
    h->kvb = h->ws.alloc(kvrows * 2 * D * es);
    bool ok = h->h && h->a && h->y && h->qkvb && h->ctx && h->kvn && h->kvb;
    if (ok && c.ff_type == 0) {
      h->mid = h->ws.alloc(rows * I * es);
      ok = h->mid != nullptr;
    } else if (ok) {
      h->moe_ws = h->ws.alloc(moe_workspace_bytes(rows, D, I, c.moe_experts, c.moe_topk, dt));
      ok = h->moe_ws != nullptr;
    }
    if (!ok) {
      h->ws.release();
      return YMT3_ERR_CUDA;
    }
    h->cap_rows = rows;
    h->cap_kv = kvrows;
  }
  int rc;
  // latent array broadcast over (b, t) (+ trainable latent position table)
  if ((rc = tile_rows(h->latents, h->latent_pos, h->h, rows, K, D, dt, s))) return rc;
  size_t li = 0;
  for (int b = 0; b < c.num_blocks; ++b) {
    if ((rc = layer_fwd(h, h->layers[li++], 0, x, B, T, Fp, c.sca_query_residual != 0, s))) return rc;
    for (int i = 0; i < c.num_local; ++i)
      if ((rc = layer_fwd(h, h->layers[li++], 1, nullptr, B, T, Fp, true, s))) return rc;
    if (c.pos_type == 1 && b == 0)
      if ((rc = add_rows(h->h, h->temporal_pos, h->h, rows, (int)T, D, dt, s, K))) return rc;
    for (int i = 0; i < c.num_temporal; ++i)
      if ((rc = layer_fwd(h, h->layers[li++], 2, nullptr, B, T, Fp, true, s))) return rc;
  }
  return norm_fwd(c, h->h, h->fin_w, h->fin_b, out, rows, D, s);
}

extern "C" int ymt3_op_permute_btcd_bctd(int32_t dtype, const void* x, void* y, int64_t B, int64_t T, int64_t C,
                                         int64_t D, void* stream) {
  return permute_btcd_bctd(x, y, B, T, C, D, dtype, (cudaStream_t)stream);
}

extern "C" int ymt3_op_convert(const void* src, int32_t sd, void* dst, int32_t dd, int64_t n, void* stream) {
  return convert(src, sd, dst, dd, n, (cudaStream_t)stream);
}
