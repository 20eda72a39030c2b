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
#include <stdlib.h>

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
  const float* rb = p.rel_bias ? p.rel_bias + (int64_t)h * p.rel_stride + p.rel_center : nullptr;

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
        if (rb && ok && qi < p.Sq) s[r][c] += rb[jg - qi];   // T5 relative position bias (shared by all layers)
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

// ------------------------------------------------------------------------------------------------
// Tiny-sequence attention on the tensor cores (bf16 IO, dk = 16, Sq, Sk <= 128, non-causal): the Perceiver-TF
// latent (S = 26) and temporal (S = 110) self-attention of the throughput path.  ONE WARP per (batch, head):
// Q, K, V rows (32 bytes each) are staged in shared memory as bf16 (lane per row; scale and rotate-half RoPE
// applied in fp32 on the way in), then per 16-query tile  S = Q K^T  (mma.sync m16n8k16, one k-step, B fragments by
// ldmatrix), a register softmax over the whole key range (fp32, quad shuffles), P repacked in registers from the
// accumulator layout into A fragments, and  O = P V  (B fragments by ldmatrix.trans).  The fp32 SIMT kernel above
// ran at 16 TFLOP/s (2.5 ms per temporal layer at B = 256); this one is bound by the 32-byte row gathers.
// 16-byte chunk c of row r sits at r*32 + ((c ^ ((r >> 2) & 1)) << 4): conflict-free ldmatrix with 32-byte rows.
// ------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ uint32_t tc_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void tc_ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void tc_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tc_pack(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// stage `rows` rows of 16 bf16 (global row stride ss) into the swizzled tile; rows [rows, rows_pad) are zeroed.
// HALF = rope_dim / 2 (compile time so the value array stays in registers); position = row index
template <int HALF>
__device__ __forceinline__ void tc_stage(unsigned char* tile, const __nv_bfloat16* src, int64_t ss, int rows, int rows_pad,
                                         float scale, const float* rope_cos, const float* rope_sin, int lane) {
  for (int r = lane; r < rows_pad; r += 32) {
    uint4 c0 = make_uint4(0u, 0u, 0u, 0u), c1 = c0;
    if (r < rows) {
      const uint4* g = reinterpret_cast<const uint4*>(src + (int64_t)r * ss);
      c0 = g[0];
      c1 = g[1];
      if (HALF > 0 || scale != 1.0f) {
        float v[16];
        const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&c0);
        const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&c1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 a = __bfloat1622float2(h0[i]), b = __bfloat1622float2(h1[i]);
          v[2 * i] = a.x * scale; v[2 * i + 1] = a.y * scale;
          v[8 + 2 * i] = b.x * scale; v[8 + 2 * i + 1] = b.y * scale;
        }
#pragma unroll
        for (int i = 0; i < HALF; ++i) {
          const float c = rope_cos[r * HALF + i], sn = rope_sin[r * HALF + i];
          const float x1 = v[i], x2 = v[i + HALF];
          v[i] = x1 * c - x2 * sn;
          v[i + HALF] = x2 * c + x1 * sn;
        }
        c0 = make_uint4(tc_pack(v[0], v[1]), tc_pack(v[2], v[3]), tc_pack(v[4], v[5]), tc_pack(v[6], v[7]));
        c1 = make_uint4(tc_pack(v[8], v[9]), tc_pack(v[10], v[11]), tc_pack(v[12], v[13]), tc_pack(v[14], v[15]));
      }
    }
    *reinterpret_cast<uint4*>(tile + tc_row_off(r, 0)) = c0;
    *reinterpret_cast<uint4*>(tile + tc_row_off(r, 1)) = c1;
  }
}

// NT = Sk_pad / 8 key tiles (compile time so the score accumulators stay in registers)
template <int NT, int HALF>
__global__ void __launch_bounds__(128) attn_small_tc_kernel(AttnParams p, int64_t total_bh) {
  extern __shared__ __align__(16) unsigned char tc_sm[];
  constexpr int SKP = NT * 8;                 // padded key count (multiple of 16)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t bh = (int64_t)blockIdx.x * 4 + warp;
  if (bh >= total_bh) return;                 // whole warp; no block-level barrier is used below
  const int sqp = (p.Sq + 15) & ~15;
  unsigned char* Qs = tc_sm + (size_t)warp * (sqp + 2 * SKP) * 32;
  unsigned char* Ks = Qs + (size_t)sqp * 32;
  unsigned char* Vs = Ks + (size_t)SKP * 32;
  const int h = (int)(bh % p.H);
  const int64_t b = bh / p.H;
  int64_t bo = b, bi = 0;
  if (p.inner > 1) {
    bo = b / p.inner;
    bi = b - bo * p.inner;
  }
  typedef __nv_bfloat16 T;
  const T* Q = static_cast<const T*>(p.Q) + bo * p.q_sb + bi * p.q_sb2 + (int64_t)h * p.q_sh;
  const T* K = static_cast<const T*>(p.K) + bo * p.k_sb + bi * p.k_sb2 + (int64_t)h * p.k_sh;
  const T* V = static_cast<const T*>(p.V) + bo * p.v_sb + bi * p.v_sb2 + (int64_t)h * p.v_sh;
  T* O = static_cast<T*>(p.O) + bo * p.o_sb + bi * p.o_sb2 + (int64_t)h * p.o_sh;
  const int kv_len = p.kv_len ? min(p.kv_len[b], p.Sk) : p.Sk;
  tc_stage<HALF>(Qs, Q, p.q_ss, p.Sq, sqp, p.scale, p.rope_cos, p.rope_sin, lane);
  tc_stage<HALF>(Ks, K, p.k_ss, kv_len, SKP, 1.0f, p.rope_cos, p.rope_sin, lane);
  tc_stage<0>(Vs, V, p.v_ss, kv_len, SKP, 1.0f, nullptr, nullptr, lane);
  __syncwarp();

  const int g = lane >> 2, t = lane & 3;
  const int lr = lane & 7, lm = lane >> 3;
  const uint32_t qs = tc_smem_addr(Qs), ks = tc_smem_addr(Ks), vs = tc_smem_addr(Vs);
  for (int m0 = 0; m0 < p.Sq; m0 += 16) {
    // A fragment of the 16 x 16 query tile: matrices (rows 0-7, c0), (rows 8-15, c0), (rows 0-7, c1), (rows 8-15, c1)
    uint32_t a0, a1, a2, a3;
    tc_ldsm_x4(qs + tc_row_off(m0 + lr + ((lm & 1) << 3), lm >> 1), a0, a1, a2, a3);
    float sacc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; j += 2) {
      // B fragments of key tiles j, j+1: matrices (keys 8j.., c0), (keys 8j.., c1), (keys 8j+8.., c0), (keys 8j+8.., c1)
      uint32_t b0, b1, b2, b3;
      tc_ldsm_x4(ks + tc_row_off(8 * j + lr + ((lm >> 1) << 3), lm & 1), b0, b1, b2, b3);
      sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.f;
      sacc[j + 1][0] = sacc[j + 1][1] = sacc[j + 1][2] = sacc[j + 1][3] = 0.f;
      tc_mma(sacc[j], a0, a1, a2, a3, b0, b1);
      tc_mma(sacc[j + 1], a0, a1, a2, a3, b2, b3);
    }
    // softmax over keys [0, kv_len) for rows g (c0, c1) and g + 8 (c2, c3); this thread holds keys 8j + 2t, +1
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int k0 = 8 * j + 2 * t;
      if (k0 >= kv_len) sacc[j][0] = sacc[j][2] = -INFINITY;
      if (k0 + 1 >= kv_len) sacc[j][1] = sacc[j][3] = -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(sacc[j][0], sacc[j][1]));
      mx1 = fmaxf(mx1, fmaxf(sacc[j][2], sacc[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float LOG2E = 1.4426950408889634f;
    const float mb0 = mx0 == -INFINITY ? 0.f : mx0 * LOG2E, mb1 = mx1 == -INFINITY ? 0.f : mx1 * LOG2E;
    float l0 = 0.f, l1 = 0.f;
    float oacc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      // P of key tiles 2kk, 2kk+1 = one k-step of 16 keys, repacked from the accumulator layout into an A fragment
      float pv[8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        pv[4 * u + 0] = exp2f(fmaf(sacc[2 * kk + u][0], LOG2E, -mb0));
        pv[4 * u + 1] = exp2f(fmaf(sacc[2 * kk + u][1], LOG2E, -mb0));
        pv[4 * u + 2] = exp2f(fmaf(sacc[2 * kk + u][2], LOG2E, -mb1));
        pv[4 * u + 3] = exp2f(fmaf(sacc[2 * kk + u][3], LOG2E, -mb1));
      }
      l0 += (pv[0] + pv[1]) + (pv[4] + pv[5]);
      l1 += (pv[2] + pv[3]) + (pv[6] + pv[7]);
      const uint32_t pa0 = tc_pack(pv[0], pv[1]), pa1 = tc_pack(pv[2], pv[3]);
      const uint32_t pa2 = tc_pack(pv[4], pv[5]), pa3 = tc_pack(pv[6], pv[7]);
      // B fragments of V for keys 16kk .. 16kk+15: (keys lo, c0), (keys hi, c0), (keys lo, c1), (keys hi, c1), transposed
      uint32_t v0, v1, v2, v3;
      tc_ldsm_x4_t(vs + tc_row_off(16 * kk + lr + ((lm & 1) << 3), lm >> 1), v0, v1, v2, v3);
      tc_mma(oacc[0], pa0, pa1, pa2, pa3, v0, v1);
      tc_mma(oacc[1], pa0, pa1, pa2, pa3, v2, v3);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = l0 > 0.f ? 1.0f / l0 : 0.f, i1 = l1 > 0.f ? 1.0f / l1 : 0.f;
    // this thread: row g -> dims 2t, 2t+1 (tile 0) and 8 + 2t, +1 (tile 1); same for row g + 8
    const int r0 = m0 + g, r1 = m0 + g + 8;
    if (r0 < p.Sq) {
      T* o = O + (int64_t)r0 * p.o_ss + 2 * t;
      *reinterpret_cast<uint32_t*>(o) = tc_pack(oacc[0][0] * i0, oacc[0][1] * i0);
      *reinterpret_cast<uint32_t*>(o + 8) = tc_pack(oacc[1][0] * i0, oacc[1][1] * i0);
    }
    if (r1 < p.Sq) {
      T* o = O + (int64_t)r1 * p.o_ss + 2 * t;
      *reinterpret_cast<uint32_t*>(o) = tc_pack(oacc[0][2] * i1, oacc[0][3] * i1);
      *reinterpret_cast<uint32_t*>(o + 8) = tc_pack(oacc[1][2] * i1, oacc[1][3] * i1);
    }
    __syncwarp();
  }
}

template <int NT, int HALF>
bool launch_small_tc2(const AttnParams& p, cudaStream_t stream) {
  const int sqp = (p.Sq + 15) & ~15;
  const size_t smem = (size_t)4 * (sqp + 2 * NT * 8) * 32;
  static bool configured[64] = {false};   // per device
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    if (cudaFuncSetAttribute(attn_small_tc_kernel<NT, HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 3 * 128 * 32) !=
        cudaSuccess)
      return false;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const int64_t total_bh = (int64_t)p.B * p.H;
  const int64_t blocks = (total_bh + 3) / 4;
  if (blocks >= (1ll << 31)) return false;
  attn_small_tc_kernel<NT, HALF><<<(unsigned)blocks, 128, smem, stream>>>(p, total_bh);
  return true;
}

// ------------------------------------------------------------------------------------------------
// Tensor-core cross-attention for ONE WIDE head (bf16, dk = 128, Sq <= 32 queries, Sk <= 128 keys): the Perceiver-TF
// spectral cross-attention (26 latents attend 128 frequency tokens per time step; HF modeling_perceiver.py:135-242 with
// num_heads = 1).  The generic kernel staged fp32 K/V per 16-query tile (K/V read twice, fp32 FFMA dot products) and ran
// at ~10x its HBM floor.  Here one CTA (4 warps) owns one (batch, head): Q (32 x 128), K and V (128 x 128 each) are
// staged ONCE as bf16 by cp.async into swizzled shared memory (72 KB, 3 CTAs per SM); warp w takes query tile w & 1
// and output-dim half w >> 1: S = Q K^T for its 16 rows over all keys (mma.sync m16n8k16, B fragments by ldmatrix), a
// register softmax (fp32, quad shuffles), P repacked in registers into A fragments, O = P V for its 64 dims
// (ldmatrix.trans).  The two warps of a query tile recompute S (cheap) instead of exchanging it: no barrier after staging.
// 16-byte chunk c (0..15) of row r sits at r*256 + ((c ^ (r & 7)) << 4): conflict-free ldmatrix with 256-byte rows.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void xw_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__global__ void __launch_bounds__(128, 3) attn_wide_tc_kernel(AttnParams p) {
  extern __shared__ __align__(128) unsigned char xw_sm[];
  typedef __nv_bfloat16 T;
  constexpr int SQP = 32, SKP = 128, NT = SKP / 8;
  unsigned char* Qs = xw_sm;                   // 32 rows x 256 B
  unsigned char* Ks = Qs + SQP * 256;          // 128 rows x 256 B
  unsigned char* Vs = Ks + SKP * 256;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = (int)(blockIdx.x % p.H);
  const int64_t b = blockIdx.x / p.H;
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
  const uint32_t qs = tc_smem_addr(Qs), ks = tc_smem_addr(Ks), vs = tc_smem_addr(Vs);

  // stage: 16 chunks of 16 bytes per row; rows past the valid range are zero-filled with plain stores
  for (int idx = tid; idx < SKP * 16; idx += 128) {
    const int r = idx >> 4, c = idx & 15;
    if (r < kv_len) {
      xw_cp16(ks + xw_off(r, c), K + (int64_t)r * p.k_ss + 8 * c);
      xw_cp16(vs + xw_off(r, c), V + (int64_t)r * p.v_ss + 8 * c);
    } else {
      *reinterpret_cast<uint4*>(Ks + xw_off(r, c)) = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(Vs + xw_off(r, c)) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  for (int idx = tid; idx < SQP * 16; idx += 128) {
    const int r = idx >> 4, c = idx & 15;
    if (r < p.Sq) xw_cp16(qs + xw_off(r, c), Q + (int64_t)r * p.q_ss + 8 * c);
    else *reinterpret_cast<uint4*>(Qs + xw_off(r, c)) = make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  const int lr = lane & 7, lm = lane >> 3;
  const int m0 = (warp & 1) * 16;              // query tile of this warp
  const int dh = warp >> 1;                    // output-dim half of this warp
  if (m0 >= p.Sq) return;                      // (no barrier below)

  // S = Q K^T: 16 rows x 128 keys, 8 k-steps over the 128 dims
  float sacc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    uint32_t a0, a1, a2, a3;   // matrices (rows 0-7, c), (rows 8-15, c), (rows 0-7, c+1), (rows 8-15, c+1)
    tc_ldsm_x4(qs + xw_off(m0 + lr + ((lm & 1) << 3), 2 * kk + (lm >> 1)), a0, a1, a2, a3);
#pragma unroll
    for (int j = 0; j < NT; j += 2) {
      uint32_t b0, b1, b2, b3;  // (keys 8j.., c), (keys 8j.., c+1), (keys 8j+8.., c), (keys 8j+8.., c+1)
      tc_ldsm_x4(ks + xw_off(8 * j + lr + ((lm >> 1) << 3), 2 * kk + (lm & 1)), b0, b1, b2, b3);
      tc_mma(sacc[j], a0, a1, a2, a3, b0, b1);
      tc_mma(sacc[j + 1], a0, a1, a2, a3, b2, b3);
    }
  }
  // softmax over keys [0, kv_len) for rows g (c0, c1) and g + 8 (c2, c3); this thread holds keys 8j + 2t, +1
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int k0 = 8 * j + 2 * t;
    if (k0 >= kv_len) sacc[j][0] = sacc[j][2] = -INFINITY;
    if (k0 + 1 >= kv_len) sacc[j][1] = sacc[j][3] = -INFINITY;
    mx0 = fmaxf(mx0, fmaxf(sacc[j][0], sacc[j][1]));
    mx1 = fmaxf(mx1, fmaxf(sacc[j][2], sacc[j][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  const float sl = p.scale * 1.4426950408889634f;   // softmax(scale * s) via exp2
  const float mb0 = mx0 == -INFINITY ? 0.f : mx0 * sl, mb1 = mx1 == -INFINITY ? 0.f : mx1 * sl;
  float l0 = 0.f, l1 = 0.f;
  float oacc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    // P of key tiles 2kk, 2kk+1 = one k-step of 16 keys, repacked from the accumulator layout into an A fragment
    float pv[8];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      pv[4 * u + 0] = exp2f(fmaf(sacc[2 * kk + u][0], sl, -mb0));
      pv[4 * u + 1] = exp2f(fmaf(sacc[2 * kk + u][1], sl, -mb0));
      pv[4 * u + 2] = exp2f(fmaf(sacc[2 * kk + u][2], sl, -mb1));
      pv[4 * u + 3] = exp2f(fmaf(sacc[2 * kk + u][3], sl, -mb1));
    }
    l0 += (pv[0] + pv[1]) + (pv[4] + pv[5]);
    l1 += (pv[2] + pv[3]) + (pv[6] + pv[7]);
    const uint32_t pa0 = tc_pack(pv[0], pv[1]), pa1 = tc_pack(pv[2], pv[3]);
    const uint32_t pa2 = tc_pack(pv[4], pv[5]), pa3 = tc_pack(pv[6], pv[7]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      // V for keys 16kk..16kk+15, dims 64dh + 16nt .. +15: (keys lo, c), (keys hi, c), (keys lo, c+1), (keys hi, c+1), transposed
      uint32_t v0, v1, v2, v3;
      tc_ldsm_x4_t(vs + xw_off(16 * kk + lr + ((lm & 1) << 3), 8 * dh + 2 * nt + (lm >> 1)), v0, v1, v2, v3);
      tc_mma(oacc[2 * nt], pa0, pa1, pa2, pa3, v0, v1);
      tc_mma(oacc[2 * nt + 1], pa0, pa1, pa2, pa3, v2, v3);
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = l0 > 0.f ? 1.0f / l0 : 0.f, i1 = l1 > 0.f ? 1.0f / l1 : 0.f;
  // this thread: rows m0 + g and m0 + g + 8, dims 64dh + 8n + 2t, +1 for n = 0..7
  const int r0 = m0 + g, r1 = m0 + g + 8;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int d = 64 * dh + 8 * n + 2 * t;
    if (r0 < p.Sq) *reinterpret_cast<uint32_t*>(O + (int64_t)r0 * p.o_ss + d) = tc_pack(oacc[n][0] * i0, oacc[n][1] * i0);
    if (r1 < p.Sq) *reinterpret_cast<uint32_t*>(O + (int64_t)r1 * p.o_ss + d) = tc_pack(oacc[n][2] * i1, oacc[n][3] * i1);
  }
}

// bf16, one wide head (dk 128), few queries, <= 128 keys, 16-byte aligned rows: the kernel above
bool try_launch_wide_tc(const AttnParams& p, cudaStream_t stream) {
  if (p.dk != 128 || p.causal || p.Sq > 32 || p.Sk > 128 || p.rope_dim > 0) return false;
  if ((p.q_ss % 8 | p.k_ss % 8 | p.v_ss % 8 | p.q_sh % 8 | p.k_sh % 8 | p.v_sh % 8 | p.q_sb % 8 | p.k_sb % 8 | p.v_sb % 8 |
       p.q_sb2 % 8 | p.k_sb2 % 8 | p.v_sb2 % 8 | p.o_ss % 2 | p.o_sh % 2 | p.o_sb % 2 | p.o_sb2 % 2) != 0)
    return false;
  if ((((uintptr_t)p.Q | (uintptr_t)p.K | (uintptr_t)p.V) & 15) != 0 || ((uintptr_t)p.O & 3) != 0) return false;
  if (getenv("YMT3_NO_TC_ATTN")) return false;
  const int64_t blocks = (int64_t)p.B * p.H;
  if (blocks >= (1ll << 31)) return false;
  const size_t smem = (32 + 2 * 128) * 256;
  static bool configured[64] = {false};   // per device
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    if (cudaFuncSetAttribute(attn_wide_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return false;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  attn_wide_tc_kernel<<<(unsigned)blocks, 128, smem, stream>>>(p);
  return true;
}

template <int NT>
bool launch_small_tc(const AttnParams& p, cudaStream_t stream) {
  switch (p.rope_dim) {
    case 0: return launch_small_tc2<NT, 0>(p, stream);
    case 8: return launch_small_tc2<NT, 4>(p, stream);
    case 16: return launch_small_tc2<NT, 8>(p, stream);
    default: return false;   // other rotary widths: the fp32 kernel
  }
}

// bf16, dk 16, short non-causal sequences whose rows are 16-byte aligned: the tensor-core kernel
bool try_launch_small_tc(const AttnParams& p, cudaStream_t stream) {
  if (p.dk != 16 || p.causal || p.Sq > 128 || p.Sk > 128 || (p.rope_dim > 0 && p.Sq != p.Sk) || p.rope_dim > 16) return false;
  if ((p.q_ss % 8 | p.k_ss % 8 | p.v_ss % 8 | p.q_sh % 8 | p.k_sh % 8 | p.v_sh % 8 | p.q_sb % 8 | p.k_sb % 8 | p.v_sb % 8 |
       p.q_sb2 % 8 | p.k_sb2 % 8 | p.v_sb2 % 8 | p.o_ss % 2 | p.o_sh % 2 | p.o_sb % 2 | p.o_sb2 % 2) != 0)
    return false;
  if ((((uintptr_t)p.Q | (uintptr_t)p.K | (uintptr_t)p.V) & 15) != 0 || ((uintptr_t)p.O & 3) != 0) return false;
  if (getenv("YMT3_NO_TC_ATTN")) return false;
  const int nt = ((p.Sk + 15) / 16) * 2;
  switch (nt) {
    case 2: return launch_small_tc<2>(p, stream);
    case 4: return launch_small_tc<4>(p, stream);
    case 6: return launch_small_tc<6>(p, stream);
    case 8: return launch_small_tc<8>(p, stream);
    case 10: return launch_small_tc<10>(p, stream);
    case 12: return launch_small_tc<12>(p, stream);
    case 14: return launch_small_tc<14>(p, stream);
    default: return launch_small_tc<16>(p, stream);
  }
}

}  // namespace

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
  if (p.rel_bias) {   // only the generic kernel adds the relative position bias
    YMT3_REQUIRE(p.rope_dim == 0, "attention: relative bias and fused RoPE are exclusive");
    const int nq = ymt3_div_up(p.Sq, 16);
    const int64_t blocks = (int64_t)p.B * p.H * nq;
    YMT3_REQUIRE(blocks < (1ll << 31), "attention: grid too large");
    YMT3_REQUIRE(p.dk == 64, "attention: relative bias is implemented for head dim 64 (T5)");
    attn_kernel<T, 64><<<(unsigned)blocks, 128, 0, stream>>>(p, nq);
    YMT3_CUDA_CHECK(cudaGetLastError());
    return YMT3_OK;
  }
  if constexpr (sizeof(T) == 2) {
    if (try_launch_small_tc(p, stream) || try_launch_wide_tc(p, stream)) {
      YMT3_CUDA_CHECK(cudaGetLastError());
      return YMT3_OK;
    }
  }
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
