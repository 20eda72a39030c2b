// Kernels of the autoregressive decode step (device-resident state, no host round trip):
// token embedding + position, single-query attention over a KV cache (self: append then
// attend; cross: attend over the encoder K/V computed once), greedy token selection with
// a device-side finished mask, and the device step counter.
#include "ops.cuh"
#include "decode.cuh"
#include <type_traits>

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
                                                        T* __restrict__ x, int dim, float* __restrict__ ss) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.x;
  const int t = tok[n];
  const int s = *step;
  for (int i = threadIdx.x; i < dim; i += 128) {
    float v = dec_to_f(E[(int64_t)t * dim + i]);
    if (pos) v += dec_to_f(pos[(int64_t)s * dim + i]);
    const T o = dec_from_f<T>(v);
    x[(int64_t)n * dim + i] = o;
    if (ss) {
      // fused RMSNorm producer (dim % 32 == 0): in every round a warp covers exactly one 32-column chunk
      float sq = dec_to_f(o);
      sq *= sq;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
      if ((threadIdx.x & 31) == 0) ss[(int64_t)n * (dim >> 5) + (i >> 5)] = sq;
    }
  }
}

int embed_pos(const int* tok, const void* E, const void* pos, const int* step, void* x, int N, int dim,
              int dtype, cudaStream_t stream, float* ss_out) {
  if (N <= 0) return YMT3_OK;
  YMT3_REQUIRE(!ss_out || dim % 128 == 0, "embed_pos: sum-of-squares output needs dim %% 128 == 0");
  if (dtype == YMT3_F32)
    YMT3_CUDA_CHECK(ymt3_launch_pdl(embed_pos_kernel<float>, dim3(N), dim3(128), 0, stream, tok, (const float*)E,
                                    (const float*)pos, step, (float*)x, dim, ss_out));
  else
    YMT3_CUDA_CHECK(ymt3_launch_pdl(embed_pos_kernel<__nv_bfloat16>, dim3(N), dim3(128), 0, stream, tok,
                                    (const __nv_bfloat16*)E, (const __nv_bfloat16*)pos, step, (__nv_bfloat16*)x, dim,
                                    ss_out));
  return YMT3_OK;
}

// One CTA (128 threads) per (head, sequence), DK = 64, single pass over the cache (flash-decoding style):
// 8 lanes share one key row (16-byte slices -> a warp reads 4 consecutive rows = 512 contiguous bytes),
// the 16 lane-groups of the CTA walk the keys with stride 16 keeping a private online-softmax state
// (m, l, o[8 dims]) in registers; K and V of a key are fetched together; the 16 partial states are merged
// through shared memory at the end.  This kernel is HBM-bound (KV bytes), see DESIGN.md 3.3.
//   self mode (knew != null): K/V row of this step is appended to the cache at index *step,
//   then the query attends over keys [0, *step]. cross mode: attends over [0, fixed_len).
template <typename T> struct Slice8;
template <> struct Slice8<float> {
  struct Raw { float4 a, b; };
  static __device__ __forceinline__ Raw load_raw(const float* p) {
    Raw r;
    r.a = *reinterpret_cast<const float4*>(p);
    r.b = *reinterpret_cast<const float4*>(p + 4);
    return r;
  }
  static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[8]) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
  }
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void copy(float* d, const float* s) {
    *reinterpret_cast<float4*>(d) = *reinterpret_cast<const float4*>(s);
    *reinterpret_cast<float4*>(d + 4) = *reinterpret_cast<const float4*>(s + 4);
  }
};
template <> struct Slice8<__nv_bfloat16> {
  typedef uint4 Raw;
  static __device__ __forceinline__ Raw load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void unpack(const Raw& u, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __bfloat162float(h[i].x);
      v[2 * i + 1] = __bfloat162float(h[i].y);
    }
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __bfloat162float(h[i].x);
      v[2 * i + 1] = __bfloat162float(h[i].y);
    }
  }
  static __device__ __forceinline__ void copy(__nv_bfloat16* d, const __nv_bfloat16* s) {
    *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
  }
};

// W WARPS per (sequence, head); W = 1 (every large batch): no block barriers, no shared memory.  Lane = (grp 0..3,
// sub 0..7): the 8 lanes of a group share one key row (16-byte slices; the warp's 4 groups read 4 consecutive rows =
// 512 contiguous bytes in the self cache), each group walks keys grp, grp+4, ... four at a time (8 x 16-byte loads
// in flight per lane) with a private online-softmax state; the 4 group states are merged with xor-shuffles.
// W > 1 (few sequences: the single-channel decoders at small batch, where W = 1 leaves 2-10 warps per SM and the
// walk over up to 1024 keys is latency-bound): the W warps of a pair take the 16-key blocks round-robin and their
// states are merged through shared memory by warp 0 of the pair (split-length / flash-decoding).
template <typename T, int W>
__global__ void __launch_bounds__(W > 4 ? 32 * W : 128, sizeof(T) == 2 ? (W > 4 ? 3 : 6) : (W > 4 ? 2 : 4))
decode_attn_kernel(const T* __restrict__ q, int64_t q_ld, const T* __restrict__ knew, const T* __restrict__ vnew,
                   int64_t new_ld, T* __restrict__ Kc, T* __restrict__ Vc, int64_t c_sn, int64_t c_sh, int64_t c_ss,
                   const int* __restrict__ step, int fixed_len, int cap, int rows_cap, float scale, T* __restrict__ out,
                   int64_t out_ld, int H, int64_t total, const float* __restrict__ rel_bias, int rel_stride) {
  constexpr int DK = 64, NG = 4, U = 4;
  pdl_launch_dependents();
  pdl_wait();
  constexpr int WARPS = W > 4 ? W : 4, PPB = WARPS / W;   // warps / (sequence, head) pairs per block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wsub = warp % W;                                // this warp's share of the pair's key blocks
  int64_t pair = (int64_t)blockIdx.x * PPB + warp / W;      // (n, h) index, h fastest
  const bool valid = pair < total;
  if (W == 1 && !valid) return;
  if (!valid) pair = total - 1;                             // W > 1: idle warps still reach the block barrier
  const int h = (int)(pair % H);
  const int64_t n = pair / H;
  const int grp = lane >> 3, sub = lane & 7;
  T* Kb = Kc + n * c_sn + (int64_t)h * c_sh;
  T* Vb = Vc + n * c_sn + (int64_t)h * c_sh;
  // Every independent load is issued before the first one is consumed (the warp issues in order, so a consumed load
  // stalls everything behind it): step, q, - self mode - the new K/V row AND the warp's first 16-key block of the
  // cache, which is fetched SPECULATIVELY (rows < cap, the length bound known at launch) before *step has arrived:
  // the short-cache steps, whose cost is two dependent round trips per wave of warps, lose one of them.
  // Keys [0, lim) come from the cache (self mode: lim = *step, the rows written by earlier steps); the new row is
  // attended straight from registers at the end and appended to the cache on the side, so there is no store -> load
  // round trip and the key loop carries no "fresh row" test.  Cache rows are dense (row stride = DK, checked by the
  // launcher): the 8 loads of a block use immediate offsets from two running pointers.
  const typename Slice8<T>::Raw qraw = Slice8<T>::load_raw(q + n * q_ld + h * DK + 8 * sub);
  int s_raw = 0;
  typename Slice8<T>::Raw knraw, vnraw;
  if (knew) {
    s_raw = *step;
    // contract (include/ymt3_b200.h): *step < Lcap.  A step counter at or beyond the cache capacity would append out
    // of bounds; every warp of the launch sees the same *step, so the whole grid leaves uniformly (out is untouched)
    if (s_raw >= rows_cap) return;
    knraw = Slice8<T>::load_raw(knew + n * new_ld + h * DK + 8 * sub);
    vnraw = Slice8<T>::load_raw(vnew + n * new_ld + h * DK + 8 * sub);
  }
  constexpr int BLK = NG * U;
  const int base0 = wsub * BLK;
  const T* kp = Kb + (int64_t)(base0 + grp) * DK + 8 * sub;   // this lane's slice of its first row
  const T* vp = Vb + (int64_t)(base0 + grp) * DK + 8 * sub;
  typename Slice8<T>::Raw kk[U], vv[U];   // kept packed until use (register pressure -> occupancy)
  bool preloaded = cap > 0;               // first block already in registers (cap == 0: A/B switch, no speculation)
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (base0 + u * NG + grp < cap) {
      kk[u] = Slice8<T>::load_raw(kp + u * NG * DK);
      vv[u] = Slice8<T>::load_raw(vp + u * NG * DK);
    }
  }
  int lim = fixed_len;                    // keys [0, lim) are read from the cache
  if (knew) {
    lim = s_raw;
    if (wsub == 0 && valid) {
      if (lane < 8) *reinterpret_cast<typename Slice8<T>::Raw*>(Kb + (int64_t)lim * DK + 8 * sub) = knraw;
      else if (lane < 16) *reinterpret_cast<typename Slice8<T>::Raw*>(Vb + (int64_t)lim * DK + 8 * sub) = vnraw;
    }
  }
  if (!valid) lim = 0;
  float qv[8];
  Slice8<T>::unpack(qraw, qv);
#pragma unroll
  for (int i = 0; i < 8; ++i) qv[i] *= scale;

  float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
  // T5 relative position bias of the single query at position `lim` (self mode): rb[dist], dist = lim - key index
  const float* rb = (rel_bias && knew) ? rel_bias + (int64_t)h * rel_stride + lim : nullptr;

  // one 16-key block: scores of this group's 4 keys (8 lanes each, 3 xor-shuffles), online-softmax update.
  // MASKED = false: all 16 keys valid, straight-line code; true: the ragged last block
  auto block = [&](auto masked_tag, int base) {
    constexpr bool MASKED = decltype(masked_tag)::value;
    float sc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float kf[8], sacc = 0.f;
      Slice8<T>::unpack(kk[u], kf);
#pragma unroll
      for (int i = 0; i < 8; ++i) sacc = fmaf(qv[i], kf[i], sacc);
      sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
      sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
      sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
      if (rb && base + u * NG + grp < lim) sacc += rb[-(base + u * NG + grp)];
      sc[u] = (MASKED && base + u * NG + grp >= lim) ? -INFINITY : sacc;
    }
    const float mn = fmaxf(fmaxf(m, fmaxf(sc[0], sc[1])), fmaxf(sc[2], sc[3]));
    if (MASKED && mn == -INFINITY) return;   // this group has no valid key yet
    const float corr = expf(m - mn);         // m = -inf -> 0
    l *= corr;
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] *= corr;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float pu = expf(sc[u] - mn);     // masked key: exp(-inf) = 0, its (possibly garbage) V must not be used
      if (MASKED && sc[u] == -INFINITY) continue;
      l += pu;
      float vf[8];
      Slice8<T>::unpack(vv[u], vf);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(pu, vf[i], o[i]);
    }
    m = mn;
  };

  // uniform trip count for the whole warp (the shuffles name all 32 lanes)
  int base = base0;
  for (; base + BLK <= lim; base += W * BLK) {
    if (!preloaded) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        kk[u] = Slice8<T>::load_raw(kp + u * NG * DK);
        vv[u] = Slice8<T>::load_raw(vp + u * NG * DK);
      }
    }
    preloaded = false;
    kp += W * BLK * DK;
    vp += W * BLK * DK;
    block(std::false_type(), base);
  }
  if (base < lim) {   // ragged last block of this warp
    if (!preloaded) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (base + u * NG + grp < lim) {
          kk[u] = Slice8<T>::load_raw(kp + u * NG * DK);
          vv[u] = Slice8<T>::load_raw(vp + u * NG * DK);
        }
      }
    }
    block(std::true_type(), base);
  }
  if (knew && wsub == 0 && valid && grp == 0) {
    // the new key / value row, from registers: one more online-softmax step in group 0 (the merge below
    // treats the other groups' untouched states as empty)
    float kf[8], sacc = 0.f;
    Slice8<T>::unpack(knraw, kf);
#pragma unroll
    for (int i = 0; i < 8; ++i) sacc = fmaf(qv[i], kf[i], sacc);
    sacc += __shfl_xor_sync(0x000000ffu, sacc, 1);
    sacc += __shfl_xor_sync(0x000000ffu, sacc, 2);
    sacc += __shfl_xor_sync(0x000000ffu, sacc, 4);
    if (rb) sacc += rb[-lim];   // distance 0
    const float mn = fmaxf(m, sacc);
    const float corr = expf(m - mn), pu = expf(sacc - mn);
    l = l * corr + pu;
    float vf[8];
    Slice8<T>::unpack(vnraw, vf);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaf(pu, vf[i], o[i] * corr);
    m = mn;
  }
  __syncwarp();
  // merge the 4 group states (lanes with equal `sub`): xor 8, then xor 16
#pragma unroll
  for (int off = 8; off <= 16; off <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, off);
    const float l2 = __shfl_xor_sync(0xffffffffu, l, off);
    const float M = fmaxf(m, m2);
    const float a = m == -INFINITY ? 0.f : expf(m - M);
    const float b2 = m2 == -INFINITY ? 0.f : expf(m2 - M);
    l = l * a + l2 * b2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float o2 = __shfl_xor_sync(0xffffffffu, o[i], off);
      o[i] = o[i] * a + o2 * b2;
    }
    m = M;
  }
  if constexpr (W > 1) {
    // states of the pair's W warps -> shared memory -> merged in warp order by its first warp
    __shared__ float st[WARPS][8][10];
    if (grp == 0) {
      float* d = st[warp][sub];
      d[0] = m; d[1] = l;
#pragma unroll
      for (int i = 0; i < 8; ++i) d[2 + i] = o[i];
    }
    __syncthreads();
    if (wsub != 0 || !valid) return;
    if (grp == 0) {
#pragma unroll
      for (int w2 = 1; w2 < W; ++w2) {
        const float* d = st[warp + w2][sub];
        const float m2 = d[0], l2 = d[1];
        const float M = fmaxf(m, m2);
        const float a = m == -INFINITY ? 0.f : expf(m - M);
        const float b2 = m2 == -INFINITY ? 0.f : expf(m2 - M);
        l = l * a + l2 * b2;
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = o[i] * a + d[2 + i] * b2;
        m = M;
      }
    }
  }
  if (grp == 0) {
    const float inv = 1.0f / l;
    T* dst = out + n * out_ld + h * DK + 8 * sub;
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(dst) = make_float4(o[0] * inv, o[1] * inv, o[2] * inv, o[3] * inv);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4] * inv, o[5] * inv, o[6] * inv, o[7] * inv);
    } else {
      uint4 pk;
      __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int i = 0; i < 4; ++i) hh[i] = __floats2bfloat162_rn(o[2 * i] * inv, o[2 * i + 1] * inv);
      *reinterpret_cast<uint4*>(dst) = pk;
    }
  }
}

int decode_attention(const void* q, int64_t q_ld, const void* knew, const void* vnew, int64_t new_ld, void* Kc,
                     void* Vc, int64_t c_sn, int64_t c_sh, int64_t c_ss, int Lmax, const int* step, int fixed_len,
                     float scale, void* out, int64_t out_ld, int N, int H, int dk, int dtype, cudaStream_t stream,
                     const float* rel_bias, int rel_stride) {
  if (N <= 0) return YMT3_OK;
  YMT3_REQUIRE(dk == 64, "decode_attention: head dim must be 64 (got %d)", dk);
  YMT3_REQUIRE(c_ss == 64, "decode_attention: cache rows must be dense (row stride %lld != 64)", (long long)c_ss);
  YMT3_REQUIRE((q_ld % 8 | new_ld % 8 | c_sn % 8 | c_sh % 8 | c_ss % 8 | out_ld % 8) == 0,
               "decode_attention: strides must be multiples of 8 elements");
  const int64_t total = (int64_t)N * H;
  // warps per (sequence, head): 1 when the batch alone fills the SMs (>= 16 warps per SM), otherwise split the key
  // range over 2 / 4 / 8 warps, never finer than one 16-key block per warp at the longest length this call can see
  static int forced_w = -1;
  if (forced_w < 0) {
    const char* e = getenv("YMT3_DECODE_ATTN_WARPS");   // profiling aid: 1 / 2 / 4 / 8, 0 = automatic
    forced_w = e ? atoi(e) : 0;
  }
  const int64_t want = (int64_t)ymt3_num_sms() * 16;
  const int max_len = knew ? Lmax : fixed_len;
  static const bool no_spec = getenv("YMT3_DECODE_ATTN_NO_SPEC") != nullptr;   // A/B aid: no speculative first block
  const int cap = no_spec ? 0 : max_len;
  int w = 1;
  while (w < 8 && total * w < want && 16 * (2 * w) <= max_len) w *= 2;
  if (forced_w == 1 || forced_w == 2 || forced_w == 4 || forced_w == 8) w = forced_w;
#define YMT3_LAUNCH_DECODE_ATTN(TT, WW)                                                                                \
  YMT3_CUDA_CHECK(ymt3_launch_pdl(decode_attn_kernel<TT, WW>, dim3((unsigned)ymt3_div_up(total, (WW > 4 ? WW : 4) / WW)), \
                                  dim3(32 * (WW > 4 ? WW : 4)), 0, stream, (const TT*)q, q_ld, (const TT*)knew,         \
                                  (const TT*)vnew, new_ld, (TT*)Kc, (TT*)Vc, c_sn, c_sh, c_ss, step, fixed_len, cap, max_len, scale, \
                                  (TT*)out, out_ld, H, total, rel_bias, rel_stride))
#define YMT3_DISPATCH_DECODE_ATTN(TT)                                                                                  \
  switch (w) {                                                                                                         \
    case 1: YMT3_LAUNCH_DECODE_ATTN(TT, 1); break;                                                                     \
    case 2: YMT3_LAUNCH_DECODE_ATTN(TT, 2); break;                                                                     \
    case 4: YMT3_LAUNCH_DECODE_ATTN(TT, 4); break;                                                                     \
    default: YMT3_LAUNCH_DECODE_ATTN(TT, 8); break;                                                                    \
  }
  if (dtype == YMT3_F32) {
    YMT3_DISPATCH_DECODE_ATTN(float)
  } else {
    YMT3_DISPATCH_DECODE_ATTN(__nv_bfloat16)
  }
#undef YMT3_DISPATCH_DECODE_ATTN
#undef YMT3_LAUNCH_DECODE_ATTN
  return YMT3_OK;
}

// Greedy selection, second half.  The arg-max itself is fused into the vocab-projection epilogue
// (GemmParams::argmax_out: largest logit, first index among equals = torch.argmax); ONE block reads the packed key of
// every row and applies the loop rules of the reference's greedy generate:
//   tokens_out[n, *step] = finished[n] ? pad : argmax;  finished |= (tok == eos);  cur_tok[n] = that token
//   while *step < n_forced (task prefix) the next input is forced[n, *step], nothing is emitted, EOS is not tested
// then re-arms the keys (0) for the next step, publishes the number of unfinished rows for the host's optional early
// stop and - after a block barrier - advances the device step counter, so a step ends with this single launch.
__global__ void __launch_bounds__(1024)
select_advance_kernel(unsigned long long* __restrict__ keys, int N, int* __restrict__ step, int* __restrict__ cur_tok,
                      int* __restrict__ finished, int* __restrict__ tokens_out, int max_len, int eos_id, int pad_id,
                      int stop_at_eos, int* __restrict__ unfinished_count, const int* __restrict__ forced,
                      int n_forced, int* __restrict__ score_out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ int s_unfinished;
  if (threadIdx.x == 0) s_unfinished = 0;
  __syncthreads();
  const int s = *step;
  int mine = 0;
  for (int n = threadIdx.x; n < N; n += 1024) {
    const unsigned long long key = keys[n];
    keys[n] = 0ull;
    if (s < n_forced) {   // task prefix: teacher-forced input, nothing emitted, EOS not tested
      cur_tok[n] = forced[(int64_t)n * n_forced + s];
      // teacher-forced scoring (ymt3_t5dec_score_forced): record what the model WOULD have emitted at this step
      if (score_out) score_out[(int64_t)n * n_forced + s] = key ? (int)(0xFFFFFFFFu - (unsigned int)(key & 0xFFFFFFFFull)) : pad_id;
      ++mine;
      continue;
    }
    const int so = s - n_forced;
    const int bi = key ? (int)(0xFFFFFFFFu - (unsigned int)(key & 0xFFFFFFFFull)) : pad_id;
    int fin = finished[n];
    const int tok = fin ? pad_id : bi;
    if (stop_at_eos && tok == eos_id) fin = 1;
    finished[n] = fin;
    cur_tok[n] = tok;
    if (so < max_len) tokens_out[(int64_t)n * max_len + so] = tok;
    if (!fin) ++mine;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_unfinished, mine);
  __syncthreads();   // every thread has read *step
  if (threadIdx.x == 0) {
    unfinished_count[s & 1] = s_unfinished;   // rows still running after step s (host early-stop poll)
    unfinished_count[(s + 1) & 1] = 0;
    *step = s + 1;
  }
}

int select_advance(unsigned long long* keys, int N, int* step, int* cur_tok, int* finished, int* tokens_out,
                   int max_len, int eos_id, int pad_id, int stop_at_eos, int* unfinished_count, const int* forced,
                   int n_forced, cudaStream_t stream, int* score_out) {
  if (N <= 0) return YMT3_OK;
  YMT3_CUDA_CHECK(ymt3_launch_pdl(select_advance_kernel, dim3(1), dim3(1024), 0, stream, keys, N, step, cur_tok, finished,
                                  tokens_out, max_len, eos_id, pad_id, stop_at_eos, unfinished_count, forced, n_forced, score_out));
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
