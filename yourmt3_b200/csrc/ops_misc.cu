// Normalisation and elementwise kernels (fp32 statistics, f32 or bf16 IO).
#include "ops.cuh"

namespace ymt3 {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per row
template <typename T>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                      T* __restrict__ y, int64_t rows, int dim, float eps) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* xr = x + row * dim;
  float ss = 0.f;
  for (int i = lane; i < dim; i += 32) {
    float v = to_f(xr[i]);
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss / (float)dim + eps);
  T* yr = y + row * dim;
  for (int i = lane; i < dim; i += 32) yr[i] = from_f<T>(w[i] * (to_f(xr[i]) * inv));
}

template <typename T>
__global__ void __launch_bounds__(256) layernorm_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, T* __restrict__ y,
                                                        int64_t rows, int dim, float eps) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* xr = x + row * dim;
  float s = 0.f;
  for (int i = lane; i < dim; i += 32) s += to_f(xr[i]);
  const float mean = warp_sum(s) / (float)dim;
  float ss = 0.f;
  for (int i = lane; i < dim; i += 32) {
    float d = to_f(xr[i]) - mean;
    ss = fmaf(d, d, ss);
  }
  const float inv = rsqrtf(warp_sum(ss) / (float)dim + eps);
  T* yr = y + row * dim;
  for (int i = lane; i < dim; i += 32)
    yr[i] = from_f<T>((to_f(xr[i]) - mean) * inv * w[i] + (b ? b[i] : 0.f));
}

// Vectorised norms, 16-byte loads, the row stays in registers between the statistics pass and the write.
// LPR = lanes per row (power of two <= 32): narrow rows (dim = 128 bf16 -> 16 lanes) pack 32/LPR rows into a warp;
// wide rows use the whole warp with up to MAXV vectors per lane.  RMS = false -> LayerNorm (two-pass variance).
template <typename T, bool RMS, int LPR, int MAXV>
__global__ void __launch_bounds__(256) norm_vec_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ b, T* __restrict__ y, int64_t rows,
                                                       int dim, float eps) {
  constexpr int V = 16 / sizeof(T), RPW = 32 / LPR;
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int64_t row = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool live = row < rows;
  const T* xr = x + (live ? row : 0) * dim;
  float v[MAXV][V];
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int i = (k * LPR + sub) * V;
    if (live && i < dim) {
      const uint4 u = *reinterpret_cast<const uint4*>(xr + i);
      if constexpr (sizeof(T) == 4) {
        const float* f = reinterpret_cast<const float*>(&u);
#pragma unroll
        for (int q = 0; q < V; ++q) v[k][q] = f[q];
      } else {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int q = 0; q < V / 2; ++q) {
          v[k][2 * q] = __bfloat162float(h[q].x);
          v[k][2 * q + 1] = __bfloat162float(h[q].y);
        }
      }
#pragma unroll
      for (int q = 0; q < V; ++q) {
        s += v[k][q];
        ss = fmaf(v[k][q], v[k][q], ss);
      }
    }
  }
  auto group_sum = [](float t) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
  };
  float mean = 0.f, inv;
  if constexpr (RMS) {
    inv = rsqrtf(group_sum(ss) / (float)dim + eps);
  } else {
    mean = group_sum(s) / (float)dim;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      const int i = (k * LPR + sub) * V;
      if (live && i < dim) {
#pragma unroll
        for (int q = 0; q < V; ++q) {
          const float d = v[k][q] - mean;
          sq = fmaf(d, d, sq);
        }
      }
    }
    inv = rsqrtf(group_sum(sq) / (float)dim + eps);
  }
  if (!live) return;
  T* yr = y + row * dim;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int i = (k * LPR + sub) * V;
    if (i < dim) {
      float o[V];
#pragma unroll
      for (int q = 0; q < V; ++q) {
        if constexpr (RMS) o[q] = w[i + q] * (v[k][q] * inv);
        else o[q] = (v[k][q] - mean) * inv * w[i + q] + (b ? b[i + q] : 0.f);
      }
      uint4 u;
      if constexpr (sizeof(T) == 4) {
        float* f = reinterpret_cast<float*>(&u);
#pragma unroll
        for (int q = 0; q < V; ++q) f[q] = o[q];
      } else {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int q = 0; q < V / 2; ++q) h[q] = __floats2bfloat162_rn(o[2 * q], o[2 * q + 1]);
      }
      *reinterpret_cast<uint4*>(yr + i) = u;
    }
  }
}

template <typename T, bool RMS>
static void launch_norm_vec(const T* x, const float* w, const float* b, T* y, int64_t rows, int dim, float eps,
                            cudaStream_t stream) {
  constexpr int V = 16 / sizeof(T);
  const int vecs = dim / V;   // 16-byte vectors per row
#define NV(LPR, MAXV)                                                                                          \
  ymt3_launch_pdl(norm_vec_kernel<T, RMS, LPR, MAXV>, dim3((unsigned)((rows + 8 * (32 / LPR) - 1) / (8 * (32 / LPR)))), \
                  dim3(256), 0, stream, x, w, b, y, rows, dim, eps)
  if (vecs <= 8) NV(8, 1);
  else if (vecs <= 16) NV(16, 1);
  else if (vecs <= 32) NV(32, 1);
  else if (vecs <= 64) NV(32, 2);
  else NV(32, 4);
#undef NV
}

static bool norm_vec_ok(const void* x, const void* y, int dim, int dtype) {
  const int V = dtype == YMT3_F32 ? 4 : 8;
  return dim % V == 0 && dim <= 32 * 4 * V && (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
}

int rmsnorm(const void* x, const float* w, void* y, int64_t rows, int dim, float eps, int dtype,
            cudaStream_t stream) {
  if (rows <= 0) return YMT3_OK;
  YMT3_REQUIRE(x && w && y && dim > 0, "rmsnorm: bad argument");
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (norm_vec_ok(x, y, dim, dtype)) {
    if (dtype == YMT3_F32) launch_norm_vec<float, true>((const float*)x, w, nullptr, (float*)y, rows, dim, eps, stream);
    else launch_norm_vec<__nv_bfloat16, true>((const __nv_bfloat16*)x, w, nullptr, (__nv_bfloat16*)y, rows, dim, eps, stream);
    YMT3_CUDA_CHECK(cudaGetLastError());
    return YMT3_OK;
  }
  if (dtype == YMT3_F32)
    rmsnorm_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, w, (float*)y, rows, dim, eps);
  else
    rmsnorm_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, w, (__nv_bfloat16*)y, rows,
                                                            dim, eps);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

// fp32 residual stream in -> bf16 GEMM operand out (one warp per row, 16-byte loads, the row stays in registers)
template <int MAXV>
__global__ void __launch_bounds__(256) rmsnorm_f32_bf16_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               __nv_bfloat16* __restrict__ y, int64_t rows, int dim,
                                                               float eps) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * dim);
  const int vecs = dim >> 2;
  float4 v[MAXV];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int i = lane + 32 * j;
    if (i < vecs) {
      v[j] = xr[i];
      ss = fmaf(v[j].x, v[j].x, fmaf(v[j].y, v[j].y, fmaf(v[j].z, v[j].z, fmaf(v[j].w, v[j].w, ss))));
    }
  }
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss / (float)dim + eps);
  uint2* yr = reinterpret_cast<uint2*>(y + row * dim);
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int i = lane + 32 * j;
    if (i < vecs) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + i);
      uint2 o;
      *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(wv.x * (v[j].x * inv), wv.y * (v[j].y * inv));
      *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(wv.z * (v[j].z * inv), wv.w * (v[j].w * inv));
      yr[i] = o;
    }
  }
}

int rmsnorm_f32_bf16(const float* x, const float* w, void* y, int64_t rows, int dim, float eps, cudaStream_t stream) {
  if (rows <= 0) return YMT3_OK;
  YMT3_REQUIRE(x && w && y && dim > 0 && dim % 4 == 0 && dim <= 2048, "rmsnorm_f32_bf16: dim must be a multiple of 4, <= 2048");
  YMT3_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)w) & 15) == 0, "rmsnorm_f32_bf16: 16-byte alignment");
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (dim <= 512) rmsnorm_f32_bf16_kernel<4><<<grid, 256, 0, stream>>>(x, w, (__nv_bfloat16*)y, rows, dim, eps);
  else rmsnorm_f32_bf16_kernel<16><<<grid, 256, 0, stream>>>(x, w, (__nv_bfloat16*)y, rows, dim, eps);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

int layernorm(const void* x, const float* w, const float* b, void* y, int64_t rows, int dim, float eps,
              int dtype, cudaStream_t stream) {
  if (rows <= 0) return YMT3_OK;
  YMT3_REQUIRE(x && w && y && dim > 0, "layernorm: bad argument");
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (norm_vec_ok(x, y, dim, dtype)) {
    if (dtype == YMT3_F32) launch_norm_vec<float, false>((const float*)x, w, b, (float*)y, rows, dim, eps, stream);
    else launch_norm_vec<__nv_bfloat16, false>((const __nv_bfloat16*)x, w, b, (__nv_bfloat16*)y, rows, dim, eps, stream);
    YMT3_CUDA_CHECK(cudaGetLastError());
    return YMT3_OK;
  }
  if (dtype == YMT3_F32)
    layernorm_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, w, b, (float*)y, rows, dim, eps);
  else
    layernorm_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, w, b, (__nv_bfloat16*)y,
                                                              rows, dim, eps);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

template <typename T>
__global__ void __launch_bounds__(256) add_rows_kernel(const T* __restrict__ x, const T* __restrict__ table,
                                                       T* __restrict__ y, int64_t total, int period, int dim,
                                                       int64_t div) {
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  int64_t r = i / dim;
  int c = (int)(i - r * dim);
  y[i] = from_f<T>(to_f(x[i]) + to_f(table[((r / div) % period) * dim + c]));
}

template <typename T>
__global__ void __launch_bounds__(256) permute_btcd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t total,
                                                           int64_t Tn, int64_t Cn, int64_t Dn) {
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;   // index into y (b, c, t, d)
  if (i >= total) return;
  const int64_t d = i % Dn, t = (i / Dn) % Tn, c = (i / (Dn * Tn)) % Cn, b = i / (Dn * Tn * Cn);
  y[i] = x[((b * Tn + t) * Cn + c) * Dn + d];
}

// cross-attention K/V: (N*T, [K|V], H, dk) rows straight out of the projection GEMM  ->  K: (N, H, T, dk) then
// V: (N, H, T, dk), so that one (sequence, head) is two contiguous T*dk blocks for the decode attention kernel.
// 16 bytes per thread.
template <typename T>
__global__ void __launch_bounds__(256) split_kv_heads_kernel(const T* __restrict__ kv, T* __restrict__ Kout,
                                                             T* __restrict__ Vout, int64_t total, int Tn, int H, int dk) {
  constexpr int V = 16 / sizeof(T);
  const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * V;   // index into Kout (n, h, t, d)
  if (i >= total) return;
  const int d = (int)(i % dk);
  const int t = (int)((i / dk) % Tn);
  const int h = (int)((i / ((int64_t)dk * Tn)) % H);
  const int64_t n = i / ((int64_t)dk * Tn * H);
  const int64_t src = ((n * Tn + t) * 2) * (int64_t)H * dk + (int64_t)h * dk + d;
  *reinterpret_cast<uint4*>(Kout + i) = *reinterpret_cast<const uint4*>(kv + src);
  *reinterpret_cast<uint4*>(Vout + i) = *reinterpret_cast<const uint4*>(kv + src + (int64_t)H * dk);
}

int split_kv_heads(const void* kv, void* Kout, void* Vout, int64_t N, int Tn, int H, int dk, int dtype,
                   cudaStream_t stream) {
  const int64_t total = N * Tn * H * dk;
  if (total <= 0) return YMT3_OK;
  YMT3_REQUIRE(dk % 8 == 0, "split_kv_heads: dk must be a multiple of 8");
  const int64_t vec = dtype == YMT3_F32 ? 4 : 8;
  const unsigned grid = (unsigned)((total / vec + 255) / 256);
  if (dtype == YMT3_F32)
    split_kv_heads_kernel<float><<<grid, 256, 0, stream>>>((const float*)kv, (float*)Kout, (float*)Vout, total, Tn, H, dk);
  else
    split_kv_heads_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)kv, (__nv_bfloat16*)Kout,
                                                                   (__nv_bfloat16*)Vout, total, Tn, H, dk);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

int permute_btcd_bctd(const void* x, void* y, int64_t B, int64_t T, int64_t C, int64_t D, int dtype,
                      cudaStream_t stream) {
  const int64_t total = B * T * C * D;
  if (total <= 0) return YMT3_OK;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == YMT3_F32)
    permute_btcd_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, (float*)y, total, T, C, D);
  else
    permute_btcd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, total, T, C, D);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

int add_rows(const void* x, const void* table, void* y, int64_t rows, int period, int dim, int dtype,
             cudaStream_t stream, int64_t div) {
  if (rows <= 0) return YMT3_OK;
  const int64_t total = rows * dim;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == YMT3_F32)
    add_rows_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, (const float*)table, (float*)y, total,
                                                     period, dim, div);
  else
    add_rows_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)table,
                                                             (__nv_bfloat16*)y, total, period, dim, div);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

template <typename T>
__global__ void __launch_bounds__(256) tile_rows_kernel(const T* __restrict__ table, const T* __restrict__ table2,
                                                        T* __restrict__ y, int64_t total, int period, int dim) {
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  int64_t r = i / dim;
  int c = (int)(i - r * dim);
  const int64_t src = (r % period) * dim + c;
  float v = to_f(table[src]);
  if (table2) v += to_f(table2[src]);
  y[i] = from_f<T>(v);
}

int tile_rows(const void* table, const void* table2, void* y, int64_t rows, int period, int dim, int dtype,
              cudaStream_t stream) {
  if (rows <= 0) return YMT3_OK;
  const int64_t total = rows * dim;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == YMT3_F32)
    tile_rows_kernel<float><<<grid, 256, 0, stream>>>((const float*)table, (const float*)table2, (float*)y, total,
                                                      period, dim);
  else
    tile_rows_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)table, (const __nv_bfloat16*)table2,
                                                              (__nv_bfloat16*)y, total, period, dim);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

// one thread per (row, head, pair)
template <typename T>
__global__ void __launch_bounds__(256)
rope_kernel(T* __restrict__ x, int64_t total, int64_t ld, int col0, int heads, int dh, int half, int64_t pos_div,
            int pos_mod, const float* __restrict__ cos_t, const float* __restrict__ sin_t) {
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int pr = (int)(i % half);
  const int h = (int)((i / half) % heads);
  const int64_t r = i / ((int64_t)half * heads);
  const int pos = (int)((r / pos_div) % pos_mod);
  T* p = x + r * ld + col0 + h * dh + pr;
  const float c = cos_t[pos * half + pr], s = sin_t[pos * half + pr];
  const float x1 = to_f(p[0]), x2 = to_f(p[half]);
  p[0] = from_f<T>(x1 * c - x2 * s);
  p[half] = from_f<T>(x2 * c + x1 * s);
}

int rope_inplace(void* x, int64_t rows, int64_t ld, int col0, int heads, int dh, int rot, int64_t pos_div, int pos_mod,
                 const float* cos_t, const float* sin_t, int dtype, cudaStream_t stream) {
  if (rows <= 0 || rot <= 0) return YMT3_OK;
  YMT3_REQUIRE(rot % 2 == 0 && rot <= dh, "rope: bad rotary dim");
  const int half = rot / 2;
  const int64_t total = rows * heads * half;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == YMT3_F32)
    rope_kernel<float><<<grid, 256, 0, stream>>>((float*)x, total, ld, col0, heads, dh, half, pos_div, pos_mod, cos_t, sin_t);
  else
    rope_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((__nv_bfloat16*)x, total, ld, col0, heads, dh, half, pos_div,
                                                         pos_mod, cos_t, sin_t);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

template <typename S, typename D>
__global__ void __launch_bounds__(256) convert_kernel(const S* __restrict__ s, D* __restrict__ d, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) d[i] = from_f<D>(to_f(s[i]));
}

int convert(const void* src, int sd, void* dst, int dd, int64_t n, cudaStream_t stream) {
  if (n <= 0) return YMT3_OK;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (sd == YMT3_F32 && dd == YMT3_BF16)
    convert_kernel<float, __nv_bfloat16><<<grid, 256, 0, stream>>>((const float*)src, (__nv_bfloat16*)dst, n);
  else if (sd == YMT3_BF16 && dd == YMT3_F32)
    convert_kernel<__nv_bfloat16, float><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)src, (float*)dst, n);
  else if (sd == YMT3_F32 && dd == YMT3_F32)
    YMT3_CUDA_CHECK(cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToDevice, stream));
  else
    YMT3_CUDA_CHECK(cudaMemcpyAsync(dst, src, n * 2, cudaMemcpyDeviceToDevice, stream));
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

}  // namespace ymt3
