// Generic small-sequence multi-head attention (encoder self-attention, Perceiver-TF
// spectral cross-attention / latent / temporal self-attention, decoder prefill).
// fp32 math (scores, online softmax, PV) for both the f32 and the bf16 IO paths:
// T5 softmax is fp32 in the reference (HF modeling_t5.py:331).
//
// CTA = 4 warps = 16 query rows of one (batch, head); each warp owns 4 rows.
// K/V are streamed through shared memory in tiles of TK keys; per tile a lane
// computes the scores of TK/32 keys for its warp's 4 rows (float4 K reads,
// broadcast Q reads), then the probabilities are exchanged through a per-warp
// smem slab and every lane accumulates its DK/32 output dims.
#include "ops.cuh"

namespace ymt3 {

template <typename T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __bfloat162float(a.x); v[1] = __bfloat162float(a.y);
  v[2] = __bfloat162float(b.x); v[3] = __bfloat162float(b.y);
}
template <typename T> __device__ __forceinline__ void store1(T* p, float v);
template <> __device__ __forceinline__ void store1<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store1<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16(v);
}

template <typename T, int DK>
__global__ void __launch_bounds__(128) attn_kernel(AttnParams p, int n_qtiles) {
  constexpr int TK = (DK >= 128) ? 32 : 64;
  constexpr int KPL = TK / 32;              // keys per lane
  constexpr int LDK = DK + 4;
  constexpr int DPL = (DK >= 32) ? DK / 32 : 1;   // output dims per lane
  constexpr int ACTIVE = DK / DPL;                // lanes that own output dims
  __shared__ __align__(16) float Ks[TK][LDK];
  __shared__ __align__(16) float Vs[TK][DK];
  __shared__ __align__(16) float Qs[16][DK];
  __shared__ __align__(16) float Ps[4][TK][4];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int64_t bid = blockIdx.x;
  const int qt = (int)(bid % n_qtiles);
  bid /= n_qtiles;
  const int h = (int)(bid % p.H);
  const int b = (int)(bid / p.H);
  const int q0 = qt * 16;
  int64_t bo = b, bi = 0;
  if (p.inner > 1) {
    bo = b / p.inner;
    bi = b - bo * p.inner;
  }
  const T* Q = static_cast<const T*>(p.Q) + bo * p.q_sb + bi * p.q_sb2 + (int64_t)h * p.q_sh;
  const T* K = static_cast<const T*>(p.K) + bo * p.k_sb + bi * p.k_sb2 + (int64_t)h * p.k_sh;
  const T* V = static_cast<const T*>(p.V) + bo * p.v_sb + bi * p.v_sb2 + (int64_t)h * p.v_sh;
  T* O = static_cast<T*>(p.O) + bo * p.o_sb + bi * p.o_sb2 + (int64_t)h * p.o_sh;
  const int kv_len = p.kv_len ? min(p.kv_len[b], p.Sk) : p.Sk;

  // Q tile (pre-scaled)
  for (int idx = tid; idx < 16 * (DK / 4); idx += 128) {
    int r = idx / (DK / 4), d = (idx % (DK / 4)) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (q0 + r < p.Sq) load4<T>(Q + (int64_t)(q0 + r) * p.q_ss + d, v);
    *reinterpret_cast<float4*>(&Qs[r][d]) =
        make_float4(v[0] * p.scale, v[1] * p.scale, v[2] * p.scale, v[3] * p.scale);
  }

  float m[4], l[4], o[4][DPL];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) o[r][i] = 0.f;
  }
  const int causal_off = p.Sk - p.Sq;

  for (int k0 = 0; k0 < kv_len; k0 += TK) {
    __syncthreads();  // previous tile fully consumed (also orders the Q tile stores)
    for (int idx = tid; idx < TK * (DK / 4); idx += 128) {
      int j = idx / (DK / 4), d = (idx % (DK / 4)) * 4;
      float kv[4] = {0.f, 0.f, 0.f, 0.f}, vv[4] = {0.f, 0.f, 0.f, 0.f};
      if (k0 + j < kv_len) {
        load4<T>(K + (int64_t)(k0 + j) * p.k_ss + d, kv);
        load4<T>(V + (int64_t)(k0 + j) * p.v_ss + d, vv);
      }
      *reinterpret_cast<float4*>(&Ks[j][d]) = make_float4(kv[0], kv[1], kv[2], kv[3]);
      *reinterpret_cast<float4*>(&Vs[j][d]) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
    __syncthreads();

    float s[4][KPL];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < KPL; ++c) s[r][c] = 0.f;
#pragma unroll 4
    for (int d = 0; d < DK; d += 4) {
      float4 kk[KPL];
#pragma unroll
      for (int c = 0; c < KPL; ++c) kk[c] = *reinterpret_cast<const float4*>(&Ks[lane + 32 * c][d]);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 q = *reinterpret_cast<const float4*>(&Qs[warp * 4 + r][d]);
#pragma unroll
        for (int c = 0; c < KPL; ++c) {
          s[r][c] = fmaf(q.x, kk[c].x, s[r][c]);
          s[r][c] = fmaf(q.y, kk[c].y, s[r][c]);
          s[r][c] = fmaf(q.z, kk[c].z, s[r][c]);
          s[r][c] = fmaf(q.w, kk[c].w, s[r][c]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int qi = q0 + warp * 4 + r;
      float tmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < KPL; ++c) {
        const int jg = k0 + lane + 32 * c;
        const bool ok = (jg < kv_len) && (!p.causal || jg <= qi + causal_off);
        if (!ok) s[r][c] = -INFINITY;
        tmax = fmaxf(tmax, s[r][c]);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, off));
      const float m_new = fmaxf(m[r], tmax);
      float corr = 1.f, psum = 0.f;
      float pr[KPL];
      if (m_new == -INFINITY) {
#pragma unroll
        for (int c = 0; c < KPL; ++c) pr[c] = 0.f;
      } else {
        corr = expf(m[r] - m_new);  // m[r] = -inf -> 0
#pragma unroll
        for (int c = 0; c < KPL; ++c) {
          pr[c] = expf(s[r][c] - m_new);
          psum += pr[c];
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, off);
      l[r] = l[r] * corr + psum;
      m[r] = m_new;
#pragma unroll
      for (int i = 0; i < DPL; ++i) o[r][i] *= corr;
#pragma unroll
      for (int c = 0; c < KPL; ++c) Ps[warp][lane + 32 * c][r] = pr[c];
    }
    __syncwarp();
    if (lane < ACTIVE) {
      const int jn = min(TK, kv_len - k0);
      for (int j = 0; j < jn; ++j) {
        const float4 pj = *reinterpret_cast<const float4*>(&Ps[warp][j][0]);
#pragma unroll
        for (int i = 0; i < DPL; ++i) {
          const float v = Vs[j][lane * DPL + i];
          o[0][i] = fmaf(pj.x, v, o[0][i]);
          o[1][i] = fmaf(pj.y, v, o[1][i]);
          o[2][i] = fmaf(pj.z, v, o[2][i]);
          o[3][i] = fmaf(pj.w, v, o[3][i]);
        }
      }
    }
    __syncwarp();
  }

  if (lane < ACTIVE) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int qi = q0 + warp * 4 + r;
      if (qi >= p.Sq) continue;
      const float inv = l[r] > 0.f ? 1.0f / l[r] : 0.f;
#pragma unroll
      for (int i = 0; i < DPL; ++i)
        store1<T>(O + (int64_t)qi * p.o_ss + lane * DPL + i, o[r][i] * inv);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Tiny-sequence attention (Perceiver-TF latent / temporal self-attention: dk = 16, S = 26 / 110).
// The whole K and V of one (batch, head) live in shared memory and ONE THREAD owns one query row:
// single-pass online softmax entirely in registers -- no shuffles, no per-tile barriers, K/V read as
// broadcast float4.  A CTA (128 threads) holds 128/G (batch, head) problems, G = 32/64/128 >= Sq.
// ------------------------------------------------------------------------------------------------
template <typename T, int DK>
__global__ void __launch_bounds__(128) attn_small_kernel(AttnParams p, int G, int64_t total_bh) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int groups = 128 / G;
  const int g = tid / G, r = tid - g * G;            // group in CTA, row (thread) in group
  const int64_t bh = (int64_t)blockIdx.x * groups + g;
  float* Ks = sm + (size_t)g * 2 * p.Sk * DK;
  float* Vs = Ks + (size_t)p.Sk * DK;
  const bool live = bh < total_bh;
  const T* Q = nullptr;
  T* O = nullptr;
  int kv_len = 0;
  if (live) {
    const int h = (int)(bh % p.H);
    const int64_t b = bh / p.H;
    int64_t bo = b, bi = 0;
    if (p.inner > 1) {
      bo = b / p.inner;
      bi = b - bo * p.inner;
    }
    Q = static_cast<const T*>(p.Q) + bo * p.q_sb + bi * p.q_sb2 + (int64_t)h * p.q_sh;
    const T* K = static_cast<const T*>(p.K) + bo * p.k_sb + bi * p.k_sb2 + (int64_t)h * p.k_sh;
    const T* V = static_cast<const T*>(p.V) + bo * p.v_sb + bi * p.v_sb2 + (int64_t)h * p.v_sh;
    O = static_cast<T*>(p.O) + bo * p.o_sb + bi * p.o_sb2 + (int64_t)h * p.o_sh;
    kv_len = p.kv_len ? min(p.kv_len[b], p.Sk) : p.Sk;
    for (int idx = r; idx < kv_len * (DK / 4); idx += G) {
      const int j = idx / (DK / 4), d = (idx % (DK / 4)) * 4;
      float kv[4], vv[4];
      load4<T>(K + (int64_t)j * p.k_ss + d, kv);
      load4<T>(V + (int64_t)j * p.v_ss + d, vv);
      *reinterpret_cast<float4*>(Ks + j * DK + d) = make_float4(kv[0], kv[1], kv[2], kv[3]);
      *reinterpret_cast<float4*>(Vs + j * DK + d) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
  }
  const int half = p.rope_dim >> 1;
  if (half > 0) {   // uniform for the whole grid
    __syncthreads();
    if (live) {
      for (int idx = r; idx < kv_len * half; idx += G) {
        const int j = idx / half, i = idx - j * half;
        const float c = p.rope_cos[j * half + i], sn = p.rope_sin[j * half + i];
        const float x1 = Ks[j * DK + i], x2 = Ks[j * DK + i + half];
        Ks[j * DK + i] = x1 * c - x2 * sn;
        Ks[j * DK + i + half] = x2 * c + x1 * sn;
      }
    }
  }
  __syncthreads();
  if (!live || r >= p.Sq) return;
  float q[DK], o[DK];
#pragma unroll
  for (int d = 0; d < DK; d += 4) {
    float t[4];
    load4<T>(Q + (int64_t)r * p.q_ss + d, t);
    q[d] = t[0] * p.scale; q[d + 1] = t[1] * p.scale; q[d + 2] = t[2] * p.scale; q[d + 3] = t[3] * p.scale;
    o[d] = o[d + 1] = o[d + 2] = o[d + 3] = 0.f;
  }
  if (half > 0) {
#pragma unroll
    for (int i = 0; i < DK / 2; ++i) {
      if (i < half) {
        const float c = p.rope_cos[r * half + i], sn = p.rope_sin[r * half + i];
        const float x1 = q[i], x2 = q[i + half];
        q[i] = x1 * c - x2 * sn;
        q[i + half] = x2 * c + x1 * sn;
      }
    }
  }
  float m = -INFINITY, l = 0.f;
  const int jmax = p.causal ? min(kv_len, r + (p.Sk - p.Sq) + 1) : kv_len;
  for (int j = 0; j < jmax; ++j) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < DK; d += 4) {
      const float4 k = *reinterpret_cast<const float4*>(Ks + j * DK + d);
      s = fmaf(q[d], k.x, s); s = fmaf(q[d + 1], k.y, s); s = fmaf(q[d + 2], k.z, s); s = fmaf(q[d + 3], k.w, s);
    }
    if (s > m) {
      const float corr = expf(m - s);   // m = -inf -> 0
      l *= corr;
#pragma unroll
      for (int d = 0; d < DK; ++d) o[d] *= corr;
      m = s;
    }
    const float pj = expf(s - m);
    l += pj;
#pragma unroll
    for (int d = 0; d < DK; d += 4) {
      const float4 v = *reinterpret_cast<const float4*>(Vs + j * DK + d);
      o[d] = fmaf(pj, v.x, o[d]); o[d + 1] = fmaf(pj, v.y, o[d + 1]);
      o[d + 2] = fmaf(pj, v.z, o[d + 2]); o[d + 3] = fmaf(pj, v.w, o[d + 3]);
    }
  }
  const float inv = l > 0.f ? 1.0f / l : 0.f;
#pragma unroll
  for (int d = 0; d < DK; ++d) store1<T>(O + (int64_t)r * p.o_ss + d, o[d] * inv);
}

template <typename T, int DK>
static bool try_launch_small(const AttnParams& p, cudaStream_t stream) {
  if (p.Sq > 128 || p.Sk > 128) return false;
  const int G = p.Sq <= 32 ? 32 : (p.Sq <= 64 ? 64 : 128);
  const int groups = 128 / G;
  const size_t smem = (size_t)groups * 2 * p.Sk * DK * sizeof(float);
  if (smem > 48 * 1024) return false;
  const int64_t total_bh = (int64_t)p.B * p.H;
  const int64_t blocks = (total_bh + groups - 1) / groups;
  if (blocks >= (1ll << 31)) return false;
  attn_small_kernel<T, DK><<<(unsigned)blocks, 128, smem, stream>>>(p, G, total_bh);
  return true;
}

template <typename T>
static int launch_attn(const AttnParams& p, cudaStream_t stream) {
  if ((p.dk == 16 && try_launch_small<T, 16>(p, stream)) || (p.dk == 32 && try_launch_small<T, 32>(p, stream))) {
    YMT3_CUDA_CHECK(cudaGetLastError());
    return YMT3_OK;
  }
  YMT3_REQUIRE(p.rope_dim == 0, "attention: fused RoPE is only available in the tiny-sequence kernel");
  const int nq = ymt3_div_up(p.Sq, 16);
  const int64_t blocks = (int64_t)p.B * p.H * nq;
  YMT3_REQUIRE(blocks < (1ll << 31), "attention: grid too large");
  switch (p.dk) {
    case 16: attn_kernel<T, 16><<<(unsigned)blocks, 128, 0, stream>>>(p, nq); break;
    case 32: attn_kernel<T, 32><<<(unsigned)blocks, 128, 0, stream>>>(p, nq); break;
    case 64: attn_kernel<T, 64><<<(unsigned)blocks, 128, 0, stream>>>(p, nq); break;
    case 128: attn_kernel<T, 128><<<(unsigned)blocks, 128, 0, stream>>>(p, nq); break;
    default:
      ymt3_set_error("attention: unsupported head dim %d (16/32/64/128)", p.dk);
      return YMT3_ERR_UNSUPPORTED;
  }
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

int attention(const AttnParams& p, int dtype, cudaStream_t stream) {
  if (p.B <= 0 || p.H <= 0 || p.Sq <= 0) return YMT3_OK;
  YMT3_REQUIRE(p.Q && p.K && p.V && p.O && p.Sk > 0, "attention: bad argument");
  YMT3_REQUIRE((p.q_ss % 4 | p.k_ss % 4 | p.v_ss % 4 | p.q_sh % 4 | p.k_sh % 4 | p.v_sh % 4 | p.q_sb % 4 |
                p.k_sb % 4 | p.v_sb % 4 | p.q_sb2 % 4 | p.k_sb2 % 4 | p.v_sb2 % 4) == 0,
               "attention: strides must be multiples of 4 elements");
  return dtype == YMT3_F32 ? launch_attn<float>(p, stream) : launch_attn<__nv_bfloat16>(p, stream);
}

}  // namespace ymt3
