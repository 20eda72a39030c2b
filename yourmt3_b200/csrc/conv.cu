// Pre-encoder "res3b" (Perceiver-TF front): 3 x [pre-activation residual 3x3 conv block +
// AvgPool(1,2) over frequency], NHWC activations (B, T, F, C).  Reference semantics restated in
// oracle/perceiver_tf.py::res_block (upstream model/conv_block.py [RECALL]); BatchNorm runs in
// eval mode and is folded: relu(bn2(conv1(a))) = relu(conv1'(a) + t2) with conv1' = conv1 * s2.
//
// Kernels here: the first (1 -> C) 3x3 convolution on CUDA cores (K = 9 is too thin for the tensor
// cores), the pool + BN + ReLU pass between blocks, and the im2col used by the fp32 parity path.
// The heavy convolutions (Cin in {64, 128}) run as implicit GEMMs on tcgen05
// (gemm_bf16_tc.cu::conv3x3_bf16_tc) or, in fp32 mode, as im2col + gemm_f32.
#include "model_common.cuh"

namespace ymt3 {

template <typename T> __device__ __forceinline__ float cv_to_f(T v);
template <> __device__ __forceinline__ float cv_to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float cv_to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T cv_from_f(float v);
template <> __device__ __forceinline__ float cv_from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 cv_from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// Block 0, conv1 (Cin = 1):  a = relu(bn1(x)) with zero padding AFTER the activation;
//   act_out[p, c]  = relu(sum_tap w1[c, tap] * a[p + tap] + t2[c])      (w1 already scaled by bn2)
//   sc_out[p, c]   = wsc[c] * x[p] + bsc[c]                             (1x1 shortcut on the raw input)
// One thread = 4 consecutive pixels (along f) x 8 output channels: the 72 weights of its channel group are
// loaded once and reused for 4 pixels; the 3 x 6 input patch is shared by the 4 pixels; 16-byte stores.
template <typename T>
__global__ void __launch_bounds__(256)
conv3x3_first_kernel(const float* __restrict__ x, int Tn, int Fn, int C, float s1, float t1,
                     const float* __restrict__ w1, const float* __restrict__ t2, const float* __restrict__ wsc,
                     const float* __restrict__ bsc, T* __restrict__ act_out, T* __restrict__ sc_out, int64_t total) {
  const int64_t gid = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (gid >= total) return;
  const int groups = C / 8;
  const int cg = (int)(gid % groups);
  const int64_t quad = gid / groups;           // index of the 4-pixel group
  const int fq = Fn / 4;
  const int f0 = (int)(quad % fq) * 4;
  const int64_t bt = quad / fq;
  const int t = (int)(bt % Tn);
  const int64_t pix0 = bt * Fn + f0;
  float a[3][6], xc[4];
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 4; ++dx) {
      const int tt = t + dy, ff = f0 + dx;
      float v = 0.f;
      if (tt >= 0 && tt < Tn && ff >= 0 && ff < Fn) {
        const float raw = x[pix0 + (int64_t)dy * Fn + dx];
        if (dy == 0 && dx >= 0 && dx < 4) xc[dx] = raw;
        v = fmaxf(fmaf(raw, s1, t1), 0.f);
      }
      a[dy + 1][dx + 1] = v;
    }
  float accv[4][8], scv[4][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cg * 8 + j;
    float w[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w[k] = w1[c * 9 + k];
    const float tb = t2[c], ws = wsc[c], bs = bsc[c];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float acc = tb;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) acc = fmaf(w[ky * 3 + kx], a[ky][p + kx], acc);
      accv[p][j] = fmaxf(acc, 0.f);
      scv[p][j] = fmaf(ws, xc[p], bs);
    }
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    T* ao = act_out + (pix0 + p) * C + cg * 8;
    T* so = sc_out + (pix0 + p) * C + cg * 8;
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(ao) = make_float4(accv[p][0], accv[p][1], accv[p][2], accv[p][3]);
      *reinterpret_cast<float4*>(ao + 4) = make_float4(accv[p][4], accv[p][5], accv[p][6], accv[p][7]);
      *reinterpret_cast<float4*>(so) = make_float4(scv[p][0], scv[p][1], scv[p][2], scv[p][3]);
      *reinterpret_cast<float4*>(so + 4) = make_float4(scv[p][4], scv[p][5], scv[p][6], scv[p][7]);
    } else {
      uint4 pa, ps;
      __nv_bfloat162* ha = reinterpret_cast<__nv_bfloat162*>(&pa);
      __nv_bfloat162* hs = reinterpret_cast<__nv_bfloat162*>(&ps);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        ha[q] = __floats2bfloat162_rn(accv[p][2 * q], accv[p][2 * q + 1]);
        hs[q] = __floats2bfloat162_rn(scv[p][2 * q], scv[p][2 * q + 1]);
      }
      *reinterpret_cast<uint4*>(ao) = pa;
      *reinterpret_cast<uint4*>(so) = ps;
    }
  }
}

int conv3x3_first(const float* x, int B, int Tn, int Fn, int C, float s1, float t1, const float* w1, const float* t2,
                  const float* wsc, const float* bsc, void* act_out, void* sc_out, int dtype, cudaStream_t stream) {
  YMT3_REQUIRE(C % 8 == 0 && Fn % 4 == 0, "conv3x3_first: C must be a multiple of 8 and F a multiple of 4");
  const int64_t total = (int64_t)B * Tn * (Fn / 4) * (C / 8);
  if (total <= 0) return YMT3_OK;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == YMT3_F32)
    conv3x3_first_kernel<float><<<grid, 256, 0, stream>>>(x, Tn, Fn, C, s1, t1, w1, t2, wsc, bsc, (float*)act_out,
                                                          (float*)sc_out, total);
  else
    conv3x3_first_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(x, Tn, Fn, C, s1, t1, w1, t2, wsc, bsc,
                                                                  (__nv_bfloat16*)act_out, (__nv_bfloat16*)sc_out, total);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

// out[b,t,f2,c] = 0.5 * (h[b,t,2*f2,c] + h[b,t,2*f2+1,c]);  act[...] = relu(out * s[c] + t[c]) (optional)
// 16 bytes (4 fp32 / 8 bf16 channels) per thread.
template <typename T>
__global__ void __launch_bounds__(256)
pool_bnrelu_kernel(const T* __restrict__ h, T* __restrict__ out, T* __restrict__ act, const float* __restrict__ s,
                   const float* __restrict__ t, int C, int64_t total) {
  constexpr int V = 16 / sizeof(T);
  const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * V;   // index into out: (rows/2, C)
  if (i >= total) return;
  const int c = (int)(i % C);
  const int64_t r2 = i / C;
  const uint4 u0 = *reinterpret_cast<const uint4*>(h + (2 * r2) * C + c);
  const uint4 u1 = *reinterpret_cast<const uint4*>(h + (2 * r2 + 1) * C + c);
  float v[V];
  if constexpr (sizeof(T) == 4) {
    const float *a = reinterpret_cast<const float*>(&u0), *b = reinterpret_cast<const float*>(&u1);
#pragma unroll
    for (int q = 0; q < V; ++q) v[q] = 0.5f * (a[q] + b[q]);
  } else {
    const __nv_bfloat162 *a = reinterpret_cast<const __nv_bfloat162*>(&u0), *b = reinterpret_cast<const __nv_bfloat162*>(&u1);
#pragma unroll
    for (int q = 0; q < V / 2; ++q) {
      v[2 * q] = 0.5f * (__bfloat162float(a[q].x) + __bfloat162float(b[q].x));
      v[2 * q + 1] = 0.5f * (__bfloat162float(a[q].y) + __bfloat162float(b[q].y));
    }
  }
  auto pack = [&](const float (&f)[V]) {
    uint4 o;
    if constexpr (sizeof(T) == 4) {
      float* p = reinterpret_cast<float*>(&o);
#pragma unroll
      for (int q = 0; q < V; ++q) p[q] = f[q];
    } else {
      __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int q = 0; q < V / 2; ++q) p[q] = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
    }
    return o;
  };
  // the activated copy is computed from the value as STORED (rounded to T), like the reference that applies
  // bn+relu to the materialised tensor
  const uint4 ov = pack(v);
  *reinterpret_cast<uint4*>(out + i) = ov;
  if (act) {
    float w[V];
    if constexpr (sizeof(T) == 4) {
#pragma unroll
      for (int q = 0; q < V; ++q) w[q] = fmaxf(fmaf(v[q], s[c + q], t[c + q]), 0.f);
    } else {
      const __nv_bfloat162* r = reinterpret_cast<const __nv_bfloat162*>(&ov);
#pragma unroll
      for (int q = 0; q < V / 2; ++q) {
        w[2 * q] = fmaxf(fmaf(__bfloat162float(r[q].x), s[c + 2 * q], t[c + 2 * q]), 0.f);
        w[2 * q + 1] = fmaxf(fmaf(__bfloat162float(r[q].y), s[c + 2 * q + 1], t[c + 2 * q + 1]), 0.f);
      }
    }
    *reinterpret_cast<uint4*>(act + i) = pack(w);
  }
}

int pool_bnrelu(const void* h, void* out, void* act, const float* s, const float* t, int64_t rows_in, int C, int dtype,
                cudaStream_t stream) {
  const int64_t total = rows_in / 2 * C;
  if (total <= 0) return YMT3_OK;
  YMT3_REQUIRE(C % 8 == 0, "pool_bnrelu: C must be a multiple of 8");
  const int64_t vec = dtype == YMT3_F32 ? 4 : 8;
  const unsigned grid = (unsigned)((total / vec + 255) / 256);
  if (dtype == YMT3_F32)
    pool_bnrelu_kernel<float><<<grid, 256, 0, stream>>>((const float*)h, (float*)out, (float*)act, s, t, C, total);
  else
    pool_bnrelu_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)h, (__nv_bfloat16*)out,
                                                                (__nv_bfloat16*)act, s, t, C, total);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

// fp32 parity path: col[p, tap*C + c] = x[p + tap, c] (zero outside), x NHWC (B,T,F,C)
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const float* __restrict__ x, float* __restrict__ col, int Tn, int Fn, int C, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;   // over pixels * 9 * C/4
  if (i >= total) return;
  const int c4 = C / 4;
  const int cq = (int)(i % c4);
  const int tap = (int)((i / c4) % 9);
  const int64_t pix = i / (9 * c4);
  const int f = (int)(pix % Fn);
  const int t = (int)((pix / Fn) % Tn);
  const int dy = tap / 3 - 1, dx = tap % 3 - 1;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t + dy >= 0 && t + dy < Tn && f + dx >= 0 && f + dx < Fn)
    v = *reinterpret_cast<const float4*>(x + (pix + (int64_t)dy * Fn + dx) * C + cq * 4);
  *reinterpret_cast<float4*>(col + pix * 9 * C + tap * C + cq * 4) = v;
}

int im2col3x3(const float* x, float* col, int B, int Tn, int Fn, int C, cudaStream_t stream) {
  const int64_t total = (int64_t)B * Tn * Fn * 9 * (C / 4);
  if (total <= 0) return YMT3_OK;
  im2col3x3_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x, col, Tn, Fn, C, total);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

}  // namespace ymt3
