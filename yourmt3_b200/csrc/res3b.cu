// ymt3_res3b_*: orchestration of the residual-conv pre-encoder over conv.cu / gemm kernels.
#include "model_common.cuh"
#include <stdlib.h>

namespace ymt3 {
int conv3x3_first(const float* x, int B, int Tn, int Fn, int C, float s1, float t1, const float* w1, const float* t2,
                  const float* wsc, const float* bsc, void* act_out, void* sc_out, int dtype, cudaStream_t stream);
int pool_bnrelu(const void* h, void* out, void* act, const float* s, const float* t, int64_t rows_in, int C, int dtype,
                cudaStream_t stream);
int im2col3x3(const float* x, float* col, int B, int Tn, int Fn, int C, cudaStream_t stream);
}  // namespace ymt3
using namespace ymt3;

namespace {

// s = w / sqrt(rv + eps), t = b - rm * s
__global__ void bn_fold_kernel(const float* w, const float* b, const float* rm, const float* rv, float eps, float* s,
                               float* t, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = w[c] / sqrtf(rv[c] + eps);
  s[c] = sc;
  t[c] = b[c] - rm[c] * sc;
}

// dst[co, (ky*3+kx)*Ci + ci] = src[co, ci, ky, kx] * (scale ? scale[co] : 1)
template <typename D>
__global__ void conv_w_pack_kernel(const float* __restrict__ src, const float* __restrict__ scale, D* __restrict__ dst,
                                   int Co, int Ci) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Co * Ci * 9) return;
  const int k = i % 9, ci = (i / 9) % Ci, co = i / (9 * Ci);
  const float v = src[i] * (scale ? scale[co] : 1.f);
  D* d = dst + (int64_t)co * 9 * Ci + k * Ci + ci;
  if constexpr (sizeof(D) == 4) *d = v; else *d = __float2bfloat16(v);
}

struct Block {
  int Ci = 0, Co = 0;
  float *s1 = nullptr, *t1 = nullptr;   // bn1 folded (applied to the block input, per Ci)
  float* t2 = nullptr;                  // bn2 shift (scale folded into conv1)
  Linear conv1, conv2;                  // (Co, 9*Ci) / (Co, 9*Co) in compute dtype (block 0 conv1: fp32 (Co, 9))
  Linear shortcut;                      // (Co, Ci) + bias, or empty (identity)
  float* sc_w32 = nullptr;              // block 0: shortcut weight as fp32 vector (Co)
};

}  // namespace

struct ymt3_res3b {
  ymt3_res3b_cfg_t c;
  DevicePool weights, ws;
  Block blk[3];
  float h_s1 = 1.f, h_t1 = 0.f;  // block-0 bn1 (single channel) as host scalars
  int64_t cap = 0;               // capacity in (B*T) rows
  void *X = nullptr, *Y = nullptr, *Z = nullptr, *P = nullptr, *Q = nullptr;
  float* col = nullptr;
};

extern "C" int ymt3_res3b_create(const ymt3_res3b_cfg_t* cfg, const ymt3_tensor_t* tensors, int n, ymt3_res3b_t** out) {
  YMT3_REQUIRE(cfg && tensors && out, "res3b_create: null argument");
  YMT3_REQUIRE(cfg->precision == YMT3_F32 || cfg->precision == YMT3_BF16, "res3b_create: bad precision");
  YMT3_REQUIRE(cfg->in_freq > 0 && cfg->in_freq % 8 == 0, "res3b_create: in_freq must be a multiple of 8");
  YMT3_REQUIRE(cfg->precision == YMT3_F32 || cfg->in_freq % 1024 == 0 ,
               "res3b_create: bf16 implicit-GEMM path needs in_freq %% 1024 == 0 (got %d)", cfg->in_freq);
  for (int i = 0; i < 3; ++i)
    YMT3_REQUIRE(cfg->channels[i] > 0 && cfg->channels[i] % 64 == 0, "res3b_create: channels must be multiples of 64");
  ymt3_res3b* h = new ymt3_res3b();
  h->c = *cfg;
  TensorTable tt{tensors, n};
  const int dt = cfg->precision;
  int rc = YMT3_OK;
  int Ci = 1;
  for (int i = 0; i < 3 && !rc; ++i) {
    Block& b = h->blk[i];
    b.Ci = Ci;
    b.Co = cfg->channels[i];
    const std::string p = "blocks." + std::to_string(i) + ".";
    auto bn = [&](const std::string& nm, int C, float** s, float** t) -> int {
      const ymt3_tensor_t *w = tt.require(p + nm + ".weight", C), *bb = tt.require(p + nm + ".bias", C),
                          *rm = tt.require(p + nm + ".running_mean", C), *rv = tt.require(p + nm + ".running_var", C);
      if (!w || !bb || !rm || !rv) return YMT3_ERR_INVALID;
      *s = (float*)h->weights.alloc(C * 4);
      *t = (float*)h->weights.alloc(C * 4);
      if (!*s || !*t) return YMT3_ERR_CUDA;
      bn_fold_kernel<<<ymt3_div_up(C, 128), 128>>>((const float*)w->data, (const float*)bb->data, (const float*)rm->data,
                                                   (const float*)rv->data, cfg->bn_eps, *s, *t, C);
      return YMT3_OK;
    };
    float* s2 = nullptr;
    if ((rc = bn("bn1", b.Ci, &b.s1, &b.t1))) break;
    if ((rc = bn("bn2", b.Co, &s2, &b.t2))) break;
    const ymt3_tensor_t* w1 = tt.require(p + "conv1.weight", b.Co, (int64_t)b.Ci * 9);
    const ymt3_tensor_t* w2 = tt.require(p + "conv2.weight", b.Co, (int64_t)b.Co * 9);
    if (!w1 || !w2) { rc = YMT3_ERR_INVALID; break; }
    // conv1 (bn2 scale folded); block 0 keeps fp32 (CUDA-core kernel)
    const int dt1 = i == 0 ? YMT3_F32 : dt;
    b.conv1.N = b.Co; b.conv1.K = 9 * b.Ci;
    b.conv1.W = h->weights.alloc((size_t)b.Co * 9 * b.Ci * dtype_size(dt1));
    b.conv1.bias = b.t2;
    b.conv2.N = b.Co; b.conv2.K = 9 * b.Co;
    b.conv2.W = h->weights.alloc((size_t)b.Co * 9 * b.Co * dtype_size(dt));
    if (!b.conv1.W || !b.conv2.W) { rc = YMT3_ERR_CUDA; break; }
    const int n1 = b.Co * b.Ci * 9, n2 = b.Co * b.Co * 9;
    if (dt1 == YMT3_F32)
      conv_w_pack_kernel<float><<<ymt3_div_up(n1, 256), 256>>>((const float*)w1->data, s2, (float*)b.conv1.W, b.Co, b.Ci);
    else
      conv_w_pack_kernel<__nv_bfloat16><<<ymt3_div_up(n1, 256), 256>>>((const float*)w1->data, s2, (__nv_bfloat16*)b.conv1.W, b.Co, b.Ci);
    if (dt == YMT3_F32)
      conv_w_pack_kernel<float><<<ymt3_div_up(n2, 256), 256>>>((const float*)w2->data, nullptr, (float*)b.conv2.W, b.Co, b.Co);
    else
      conv_w_pack_kernel<__nv_bfloat16><<<ymt3_div_up(n2, 256), 256>>>((const float*)w2->data, nullptr, (__nv_bfloat16*)b.conv2.W, b.Co, b.Co);
    if (b.Ci != b.Co) {
      const ymt3_tensor_t* ws = tt.require(p + "shortcut.weight", b.Co, b.Ci);
      const ymt3_tensor_t* bs = tt.require(p + "shortcut.bias", b.Co);
      if (!ws || !bs) { rc = YMT3_ERR_INVALID; break; }
      if ((rc = pack_vec(h->weights, {bs}, false, &b.shortcut.bias, 0))) break;
      if (i == 0) {
        if ((rc = pack_vec(h->weights, {ws}, false, &b.sc_w32, 0))) break;   // (Co,1,1,1) -> (Co)
      } else if ((rc = pack_rows(h->weights, {ws}, b.Ci, dt, false, &b.shortcut, 0))) {
        break;
      } else {
        // pack_rows reset bias; restore
        float* keep = nullptr;
        if ((rc = pack_vec(h->weights, {bs}, false, &keep, 0))) break;
        b.shortcut.bias = keep;
      }
    } else if (i == 0) {
      ymt3_set_error("res3b_create: block 0 must change the channel count (1 -> C)");
      rc = YMT3_ERR_INVALID;
      break;
    }
    Ci = b.Co;
  }
  if (!rc) {
    float st[2];
    if (cudaMemcpy(&st[0], h->blk[0].s1, 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(&st[1], h->blk[0].t1, 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
      ymt3_set_error("res3b_create: %s", cudaGetErrorString(cudaGetLastError()));
      rc = YMT3_ERR_CUDA;
    }
    h->h_s1 = st[0];
    h->h_t1 = st[1];
  }
  if (!rc && cudaDeviceSynchronize() != cudaSuccess) {
    ymt3_set_error("res3b_create: weight packing failed: %s", cudaGetErrorString(cudaGetLastError()));
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

extern "C" int ymt3_res3b_destroy(ymt3_res3b_t* h) {
  if (!h) return YMT3_OK;
  h->weights.release();
  h->ws.release();
  delete h;
  return YMT3_OK;
}

namespace {
// y(rows, Co) = epi(conv3x3(x NHWC (B,T,F,Ci)))
int conv_fwd(ymt3_res3b* h, const void* x, int B, int T, int F, int Ci, const Linear& w, void* y, int act,
             const void* residual, cudaStream_t s) {
  GemmParams p{};
  p.W = w.W; p.ldw = w.K; p.C = y; p.ldc = w.N; p.bias = w.bias; p.residual = residual; p.ldr = w.N;
  p.N = w.N; p.act = act; p.out_scale = 1.f;
  if (h->c.precision == YMT3_BF16) return conv3x3_bf16_tc(x, B, T, F, Ci, p, YMT3_BF16, s);
  int rc = im2col3x3((const float*)x, h->col, B, T, F, Ci, s);
  if (rc) return rc;
  p.A = h->col; p.lda = 9 * Ci; p.M = B * T * F; p.K = 9 * Ci;
  return gemm_f32(p, s);
}
}  // namespace

extern "C" int ymt3_res3b_forward(ymt3_res3b_t* h, const float* spec, int64_t B, int64_t T, void* out, void* stream) {
  YMT3_REQUIRE(h && out, "res3b_forward: null argument");
  if (B <= 0 || T <= 0) return YMT3_OK;
  YMT3_REQUIRE(spec, "res3b_forward: null input");
  const int F = h->c.in_freq, dt = h->c.precision;
  const int C0 = h->blk[0].Co, C1 = h->blk[1].Co, C2 = h->blk[2].Co;
  const int64_t BT = B * T;
  YMT3_REQUIRE(BT * F < (1ll << 31), "res3b_forward: too many pixels per call");
  cudaStream_t s = (cudaStream_t)stream;
  if (BT > h->cap) {
    YMT3_CUDA_CHECK(cudaStreamSynchronize(s));
    h->ws.release();
    h->cap = 0;
    const size_t es = dtype_size(dt);
    int64_t full = (int64_t)F * C0;
    if ((int64_t)F / 2 * C1 > full) full = (int64_t)F / 2 * C1;
    if ((int64_t)F / 4 * C2 > full) full = (int64_t)F / 4 * C2;
    h->X = h->ws.alloc(BT * full * es);
    h->Y = h->ws.alloc(BT * full * es);
    h->Z = h->ws.alloc(BT * full * es);
    h->P = h->ws.alloc(BT * full / 2 * es + 256);
    h->Q = h->ws.alloc(BT * full / 2 * es + 256);
    bool ok = h->X && h->Y && h->Z && h->P && h->Q;
    if (ok && dt == YMT3_F32) {
      int64_t kmax = (int64_t)F * 9 * C0;                    // block0 conv2
      if ((int64_t)F / 2 * 9 * C1 > kmax) kmax = (int64_t)F / 2 * 9 * C1;   // block1 conv2
      if ((int64_t)F / 4 * 9 * C2 > kmax) kmax = (int64_t)F / 4 * 9 * C2;
      h->col = (float*)h->ws.alloc(BT * kmax * 4);
      ok = h->col != nullptr;
    }
    if (!ok) {
      h->ws.release();
      return YMT3_ERR_CUDA;
    }
    h->cap = BT;
  }
  int rc;
  const Block &b0 = h->blk[0], &b1 = h->blk[1], &b2 = h->blk[2];
  const int iB = (int)B, iT = (int)T;
  // ---- block 0 (1 -> C0) at F ----
  if ((rc = conv3x3_first(spec, iB, iT, F, C0, h->h_s1, h->h_t1, (const float*)b0.conv1.W, b0.t2, b0.sc_w32,
                          b0.shortcut.bias, h->X, h->Y, dt, s))) return rc;
  if ((rc = conv_fwd(h, h->X, iB, iT, F, C0, b0.conv2, h->Z, 0, h->Y, s))) return rc;
  if ((rc = pool_bnrelu(h->Z, h->P, h->Q, b1.s1, b1.t1, BT * F, C0, dt, s))) return rc;
  // ---- block 1 (C0 -> C1) at F/2 ----
  if ((rc = conv_fwd(h, h->Q, iB, iT, F / 2, C0, b1.conv1, h->X, YMT3_ACT_RELU, nullptr, s))) return rc;
  if (b1.Ci != b1.Co) {
    if ((rc = linear_fwd(dt, h->P, C0, b1.shortcut, h->Y, C1, (int)(BT * F / 2), 0, 0, nullptr, 0, 1.f, dt, s))) return rc;
    if ((rc = conv_fwd(h, h->X, iB, iT, F / 2, C1, b1.conv2, h->Z, 0, h->Y, s))) return rc;
  } else if ((rc = conv_fwd(h, h->X, iB, iT, F / 2, C1, b1.conv2, h->Z, 0, h->P, s))) {
    return rc;
  }
  if ((rc = pool_bnrelu(h->Z, h->P, h->Q, b2.s1, b2.t1, BT * F / 2, C1, dt, s))) return rc;
  // ---- block 2 (C1 -> C2) at F/4 ----
  if ((rc = conv_fwd(h, h->Q, iB, iT, F / 4, C1, b2.conv1, h->X, YMT3_ACT_RELU, nullptr, s))) return rc;
  if (b2.Ci != b2.Co) {
    if ((rc = linear_fwd(dt, h->P, C1, b2.shortcut, h->Y, C2, (int)(BT * F / 4), 0, 0, nullptr, 0, 1.f, dt, s))) return rc;
    if ((rc = conv_fwd(h, h->X, iB, iT, F / 4, C2, b2.conv2, h->Z, 0, h->Y, s))) return rc;
  } else if ((rc = conv_fwd(h, h->X, iB, iT, F / 4, C2, b2.conv2, h->Z, 0, h->P, s))) {
    return rc;
  }
  return pool_bnrelu(h->Z, out, nullptr, nullptr, nullptr, BT * F / 4, C2, dt, s);
}
