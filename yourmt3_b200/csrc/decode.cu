// Kernels of the autoregressive decode step (device-resident state, no host round trip):
// token embedding + position, single-query attention over a KV cache (self: append then
// attend; cross: attend over the encoder K/V computed once), greedy token selection with
// a device-side finished mask, and the device step counter.
#include "ops.cuh"
#include "decode.cuh"

namespace ymt3 {

template <typename T> __device__ __forceinline__ float dec_to_f(T v);
template <> __device__ __forceinline__ float dec_to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float dec_to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T dec_from_f(float v);
template <> __device__ __forceinline__ float dec_from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 dec_from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// x[n, :] = E[tok[n], :] + pos[*step, :]
template <typename T>
__global__ void __launch_bounds__(128) embed_pos_kernel(const int* __restrict__ tok, const T* __restrict__ E,
                                                        const T* __restrict__ pos, const int* __restrict__ step,
                                                        T* __restrict__ x, int dim) {
  const int n = blockIdx.x;
  const int t = tok[n];
  const int s = *step;
  for (int i = threadIdx.x; i < dim; i += 128) {
    float v = dec_to_f(E[(int64_t)t * dim + i]);
    if (pos) v += dec_to_f(pos[(int64_t)s * dim + i]);
    x[(int64_t)n * dim + i] = dec_from_f<T>(v);
  }
}

int embed_pos(const int* tok, const void* E, const void* pos, const int* step, void* x, int N, int dim,
              int dtype, cudaStream_t stream) {
  if (N <= 0) return YMT3_OK;
  if (dtype == YMT3_F32)
    embed_pos_kernel<float><<<N, 128, 0, stream>>>(tok, (const float*)E, (const float*)pos, step, (float*)x, dim);
  else
    embed_pos_kernel<__nv_bfloat16><<<N, 128, 0, stream>>>(tok, (const __nv_bfloat16*)E, (const __nv_bfloat16*)pos,
                                                          step, (__nv_bfloat16*)x, dim);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

// One CTA (128 threads) per (head, sequence). DK = 64.
//   self mode (knew != null): K/V row of this step is appended to the cache at index *step,
//   then the query attends over keys [0, *step]. cross mode: attends over [0, fixed_len).
template <typename T>
__global__ void __launch_bounds__(128)
decode_attn_kernel(const T* __restrict__ q, int64_t q_ld, const T* __restrict__ knew, const T* __restrict__ vnew,
                   int64_t new_ld, T* __restrict__ Kc, T* __restrict__ Vc, int64_t c_sn, int64_t c_sh, int64_t c_ss,
                   const int* __restrict__ step, int fixed_len, float scale, T* __restrict__ out, int64_t out_ld) {
  constexpr int DK = 64;
  extern __shared__ float smem[];
  float* P = smem;                 // [len]
  __shared__ float qs[DK];
  __shared__ float red[4];
  __shared__ float part[2][DK];
  const int h = blockIdx.x, n = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  T* Kb = Kc + (int64_t)n * c_sn + (int64_t)h * c_sh;
  T* Vb = Vc + (int64_t)n * c_sn + (int64_t)h * c_sh;
  int len = fixed_len;
  if (knew) {
    const int s = *step;
    len = s + 1;
    if (tid < DK) {
      Kb[(int64_t)s * c_ss + tid] = knew[(int64_t)n * new_ld + h * DK + tid];
      Vb[(int64_t)s * c_ss + tid] = vnew[(int64_t)n * new_ld + h * DK + tid];
    }
  }
  if (tid < DK) qs[tid] = dec_to_f(q[(int64_t)n * q_ld + h * DK + tid]) * scale;
  __syncthreads();   // also makes this CTA's freshly appended K/V row visible to its threads

  // scores
  float lmax = -INFINITY;
  for (int j = tid; j < len; j += 128) {
    const T* kr = Kb + (int64_t)j * c_ss;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < DK; d += 4) {
      float k0, k1, k2, k3;
      if constexpr (sizeof(T) == 4) {
        float4 kv = *reinterpret_cast<const float4*>(kr + d);
        k0 = kv.x; k1 = kv.y; k2 = kv.z; k3 = kv.w;
      } else {
        uint2 kv = *reinterpret_cast<const uint2*>(kr + d);
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&kv.x), b = *reinterpret_cast<__nv_bfloat162*>(&kv.y);
        k0 = __bfloat162float(a.x); k1 = __bfloat162float(a.y);
        k2 = __bfloat162float(b.x); k3 = __bfloat162float(b.y);
      }
      s = fmaf(qs[d], k0, s);
      s = fmaf(qs[d + 1], k1, s);
      s = fmaf(qs[d + 2], k2, s);
      s = fmaf(qs[d + 3], k3, s);
    }
    P[j] = s;
    lmax = fmaxf(lmax, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  const float gmax = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float lsum = 0.f;
  for (int j = tid; j < len; j += 128) {
    const float e = expf(P[j] - gmax);
    P[j] = e;
    lsum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if (lane == 0) red[warp] = lsum;
  __syncthreads();
  const float inv = 1.0f / (red[0] + red[1] + red[2] + red[3]);

  // PV: thread owns dim d = tid & 63 for keys of parity group g = tid >> 6
  const int d = tid & 63, g = tid >> 6;
  float acc = 0.f;
  for (int j = g; j < len; j += 2) acc = fmaf(P[j], dec_to_f(Vb[(int64_t)j * c_ss + d]), acc);
  part[g][d] = acc;
  __syncthreads();
  if (tid < DK) out[(int64_t)n * out_ld + h * DK + tid] = dec_from_f<T>((part[0][tid] + part[1][tid]) * inv);
}

int decode_attention(const void* q, int64_t q_ld, const void* knew, const void* vnew, int64_t new_ld, void* Kc,
                     void* Vc, int64_t c_sn, int64_t c_sh, int64_t c_ss, int Lmax, const int* step, int fixed_len,
                     float scale, void* out, int64_t out_ld, int N, int H, int dk, int dtype, cudaStream_t stream) {
  if (N <= 0) return YMT3_OK;
  YMT3_REQUIRE(dk == 64, "decode_attention: head dim must be 64 (got %d)", dk);
  YMT3_REQUIRE(N <= 65535, "decode_attention: too many sequences per call (%d > 65535)", N);
  const int max_len = knew ? Lmax : fixed_len;
  const size_t smem = (size_t)max_len * sizeof(float);
  YMT3_REQUIRE(smem <= 40 * 1024, "decode_attention: sequence too long (%d)", max_len);
  dim3 grid(H, N);
  if (dtype == YMT3_F32)
    decode_attn_kernel<float><<<grid, 128, smem, stream>>>((const float*)q, q_ld, (const float*)knew,
                                                           (const float*)vnew, new_ld, (float*)Kc, (float*)Vc, c_sn,
                                                           c_sh, c_ss, step, fixed_len, scale, (float*)out, out_ld);
  else
    decode_attn_kernel<__nv_bfloat16><<<grid, 128, smem, stream>>>(
        (const __nv_bfloat16*)q, q_ld, (const __nv_bfloat16*)knew, (const __nv_bfloat16*)vnew, new_ld,
        (__nv_bfloat16*)Kc, (__nv_bfloat16*)Vc, c_sn, c_sh, c_ss, step, fixed_len, scale, (__nv_bfloat16*)out,
        out_ld);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

// Greedy selection. One warp per sequence. torch.argmax tie rule: first maximal index.
//   tokens_out[n, *step] = finished[n] ? pad : argmax(logits[n, :V]);  finished |= (tok == eos)
//   cur_tok[n] = tokens_out[n, *step];  unfinished_count += !finished (for the host's optional early stop)
__global__ void __launch_bounds__(256)
greedy_select_kernel(const float* __restrict__ logits, int64_t ld, int V, int N, const int* __restrict__ step,
                     int* __restrict__ cur_tok, int* __restrict__ finished, int* __restrict__ tokens_out,
                     int max_len, int eos_id, int pad_id, int stop_at_eos, int* __restrict__ unfinished_count) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* row = logits + (int64_t)n * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = lane; i < V; i += 32) {
    const float v = row[i];
    if (v > best || (v == best && i < bi)) {
      best = v;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  if (lane == 0) {
    const int s = *step;
    int fin = finished[n];
    int tok = fin ? pad_id : bi;
    if (stop_at_eos && tok == eos_id) fin = 1;
    finished[n] = fin;
    cur_tok[n] = tok;
    if (s < max_len) tokens_out[(int64_t)n * max_len + s] = tok;
    if (!fin) atomicAdd(unfinished_count + (s & 1), 1);
  }
}

int greedy_select(const float* logits, int64_t ld, int V, int N, const int* step, int* cur_tok, int* finished,
                  int* tokens_out, int max_len, int eos_id, int pad_id, int stop_at_eos, int* unfinished_count,
                  cudaStream_t stream) {
  if (N <= 0) return YMT3_OK;
  greedy_select_kernel<<<ymt3_div_up(N, 8), 256, 0, stream>>>(logits, ld, V, N, step, cur_tok, finished,
                                                              tokens_out, max_len, eos_id, pad_id, stop_at_eos,
                                                              unfinished_count);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

// (*step)++ and reset the unfinished counter slot the NEXT step will accumulate into
__global__ void advance_step_kernel(int* step, int* unfinished_count) {
  const int s = *step + 1;
  *step = s;
  unfinished_count[s & 1] = 0;
}

int advance_step(int* step, int* unfinished_count, cudaStream_t stream) {
  advance_step_kernel<<<1, 1, 0, stream>>>(step, unfinished_count);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

__global__ void __launch_bounds__(256) fill_i32_kernel(int* p, int v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) p[i] = v;
}

int fill_i32(int* p, int v, int64_t n, cudaStream_t stream) {
  if (n <= 0) return YMT3_OK;
  fill_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, v, n);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

}  // namespace ymt3
