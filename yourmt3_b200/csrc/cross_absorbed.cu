// Absorbed ("latent") cross-attention of the multi-channel decoder (bf16 throughput path).
//
// Reference semantics (HF modeling_t5.py:269-305 cross-attention, upstream projection_layer.py
// `mc_shared_linear` [RECALL]): enc_hs[n, t, :] = Wp z[n, t, :] + bp with z the 2 x 128 = 256 latent values of
// channel c at time t, then per layer K = Wk enc_hs, V = Wv enc_hs (6 heads x 64), softmax(q K^T) V, Wo.
// Because enc_hs is an affine image of a 256-wide latent, K and V of EVERY layer and head are linear in the
// same z.  Folding (host side, fp64, see t5mod_helper.fold_cross_projection):
//     q'_h = (Wk_h Wp)^T Wq_h x          (256-wide query in latent space; the bp term is constant over keys)
//     ctx_h = sum_t softmax_t(q'_h . z_t) z_t
//     out  = sum_h (Wo_h Wv_h Wp) ctx_h + Wo Wv bp
// turns the per-step cross-attention read from 8 layers x (K + V) x 6 heads x 64 = 6144 values per encoder
// token into ONE 256-value latent row shared by all heads (re-read per layer): 24x fewer HBM bytes for the
// kernel that was 35 % of the decode step.  Same function of the same weights; the association order differs,
// so it is used on the bf16 path only (the fp32 token-exact path keeps the reference order).
//
// Kernel: PERSISTENT CTAs (8 warps) walking sequences n = blockIdx.x, += gridDim.x.  One elected thread stages the
// (Tp x 256) bf16 latent tile and the sequence's H latent-space queries with TMA (four {64 col x rows} boxes each,
// 128-byte hardware swizzle -> conflict-free ldmatrix, one mbarrier per stage) through a ring of 1..3 stages.
// Default shape: THREE CTAs per SM with one stage each (see the launcher): their load / compute phases interleave.  Both products
// run on the tensor cores (mma.sync m16n8k16 - a skinny M = 6 heads batched product, HBM-bound by construction):
//     S (H x Tp)  = Q' (H x 256) . Z^T       B fragments by ldmatrix        (Z rows = keys)
//     O (H x 256) = P  (H x Tp)  . Z         B fragments by ldmatrix.trans
// The softmax never leaves registers except for per-warp (max, sum) partials and the bf16 P tile in shared memory;
// statistics are fp32, P is kept unnormalised and O is scaled by 1/l in fp32.
// History (profiles/r01_cross_absorbed_*): v1 one CTA per sequence, cp.async, no prefetch: 73 us for 3328 sequences
// (2.9 TB/s); v2 persistent + cp.async ring: same 73 us - ncu showed 6.8 k warp instructions per sequence, 28 % of
// them cp.async address arithmetic, issue-bound at 8 warps/SM; v3 (this) TMA + register softmax: 48 us (4.2 TB/s).
// v5 (this): same kernel, 1 stage x 3 co-resident CTAs per SM instead of 3 stages x 1 CTA: 121 -> 103.5 us for 9464
// sequences (5.8 TB/s), end to end +2.4 %.
// v4 experiment (NOT kept): a TMA producer warp + two 8-warp consumer groups on alternate sequences (named barriers,
// full/empty mbarriers) ran 43 us but hung about once per 3000 launches (0.4 s watchdog kill) even after adding
// warp reconvergence around every inline-asm wait/barrier; root cause not found in the time available.
#include "ops.cuh"
#include "decode.cuh"
#include <cuda.h>
#include <cstdlib>

namespace ymt3 {

namespace {

constexpr int ZD = 256;              // latent width handled by this kernel
constexpr int MAX_H = 8;             // heads live in rows 0..7 of the m16 tile
constexpr int NWARP = 8;
constexpr int NTHREAD = NWARP * 32;
constexpr int Q_STAGE_BYTES = 4 * MAX_H * 128;   // four {64 col x 8 row} boxes, 1024 B apart

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(addr));
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

// D += A(16x16, rows 8..15 zero) * B(16x8); only c0/c1 (rows 0..7) are meaningful
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  const uint32_t zero = 0u;
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(zero), "r"(a2), "r"(zero), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "XA_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra XA_DONE;\n\t"
      "bra XA_WAIT;\n\t"
      "XA_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_box(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// stage layout: 4 query boxes (8 rows x 128 B slots, H rows written) then 4 latent boxes (Tp rows x 128 B); within a
// box the 16-byte chunk c of row r sits at r*128 + ((c ^ (r & 7)) << 4) (TMA SWIZZLE_128B, boxes 1024-B aligned)
__device__ __forceinline__ void stage_issue(const CUtensorMap* mq, const CUtensorMap* mz, uint32_t stage, uint32_t bar,
                                            int64_t n, int H, int Tp) {
  mbar_expect_tx(bar, (uint32_t)(H + Tp) * 512u);
#pragma unroll
  for (int i = 0; i < 4; ++i) tma_box(mq, bar, stage + i * (MAX_H * 128), 64 * i, (int)(n * H));
#pragma unroll
  for (int i = 0; i < 4; ++i) tma_box(mz, bar, stage + Q_STAGE_BYTES + i * (Tp * 128), 64 * i, (int)(n * Tp));
}

// mq: q viewed as (N*H, 256) bf16;  mz: z viewed as (N*Tp, 256) bf16, rows >= T of a sequence are zero;
// out: (N, H*256) bf16 rows (leading dim out_ld)
__global__ void __launch_bounds__(NTHREAD, 1)
cross_attn_absorbed_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mz,
                           __nv_bfloat16* __restrict__ out, int64_t out_ld, int64_t N, int H, int T, int Tp,
                           int stages) {
  extern __shared__ unsigned char smem_raw[];
  // SWIZZLE_128B boxes must start on 1024-byte boundaries: align the dynamic segment by hand (1 KB slack requested)
  unsigned char* smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
  const int PLD = Tp + 8;   // probability row stride (bf16)
  const uint32_t stage_bytes = (uint32_t)Q_STAGE_BYTES + (uint32_t)Tp * 512u;
  __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(smem + (size_t)stages * stage_bytes);
  float* wmax = reinterpret_cast<float*>(P + MAX_H * PLD);   // [head][warp]
  float* wsum = wmax + MAX_H * NWARP;                        // [head][warp]
  uint64_t* bars = reinterpret_cast<uint64_t*>(wsum + MAX_H * NWARP);
  const uint32_t ring = smem_addr(smem);
  const uint32_t bar0 = smem_addr(bars);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int lr = lane & 7, lm = lane >> 3;   // ldmatrix: this lane addresses row lr of matrix lm
  const int64_t stride = gridDim.x;

  pdl_launch_dependents();   // prologue below touches only shared memory; global reads start after pdl_wait()
  if (tid == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < MAX_H * PLD; i += NTHREAD) P[i] = __float2bfloat16(0.f);   // rows >= H stay zero
  __syncthreads();
  pdl_wait();
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mq) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mz) : "memory");
    for (int p = 0; p < stages - 1; ++p) {
      const int64_t np = blockIdx.x + p * stride;
      if (np < N) stage_issue(&mq, &mz, ring + p * stage_bytes, bar0 + 8 * p, np, H, Tp);
    }
  }

  // per-thread invariant pieces of the swizzled addresses
  const uint32_t sub_bytes = (uint32_t)Tp * 128u;
  const uint32_t sx0 = (uint32_t)((lm ^ lr) << 4), sx1 = (uint32_t)(((4 + lm) ^ lr) << 4);   // scores: chunks 4kk+lm
  uint32_t qx[4];                                                                           // queries: chunks 2ks+(lm&1)
#pragma unroll
  for (int i = 0; i < 4; ++i) qx[i] = (uint32_t)(lr * 128 + (((2 * i + (lm & 1)) ^ lr) << 4));
  uint32_t px[2];                                                                           // P.Z: chunks 4w+2jj+(lm>>1)
#pragma unroll
  for (int jj = 0; jj < 2; ++jj)
    px[jj] = (uint32_t)(warp >> 1) * sub_bytes + (uint32_t)((lr + ((lm & 1) << 3)) * 128) +
             (uint32_t)(((4 * (warp & 1) + 2 * jj + (lm >> 1)) ^ lr) << 4);
  const int ntile = Tp >> 3;
  const float LOG2E = 1.4426950408889634f;

  int it = 0;
  for (int64_t n = blockIdx.x; n < N; n += stride, ++it) {
    __syncthreads();   // everyone is done with iteration it-1 (its stage, P, wsum)
    const int slot = it % stages;
    if (tid == 0) {
      if (stages >= 2) {
        const int64_t nn = n + (stages - 1) * stride;
        const int ps = (it + stages - 1) % stages;
        if (nn < N) stage_issue(&mq, &mz, ring + ps * stage_bytes, bar0 + 8 * ps, nn, H, Tp);
      } else {
        stage_issue(&mq, &mz, ring, bar0, n, H, Tp);   // single-stage fallback (very long encoders): no prefetch
      }
    }
    // mbarrier.try_wait is a per-thread poll (lanes may leave the spin loop on different iterations and the compiler
    // cannot see that inside inline asm): reconverge before the warp-synchronous ldmatrix / mma below
    mbar_wait(bar0 + 8 * slot, (uint32_t)((it / stages) & 1));
    __syncwarp();
    const uint32_t qs = ring + slot * stage_bytes;
    const uint32_t zs = qs + Q_STAGE_BYTES;

    // A fragments of Q' (matrix rows = heads; rows >= H hold stale bytes whose products land in ignored rows)
    uint32_t qa0[16], qa2[16];
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) ldsm_x2(qs + (ks >> 2) * (MAX_H * 128) + qx[ks & 3], qa0[ks], qa2[ks]);

    // scores: n-tile j = keys 8j .. 8j+7; warp w takes tiles w and w + NWARP (Tp <= 128) as two independent
    // accumulator chains; a warp without a tile recomputes tile 0 and discards it (keeps the barriers uniform)
    {
      const bool one = warp < ntile, two = warp + NWARP < ntile;
      const int j1 = one ? warp : 0, j2 = two ? warp + NWARP : j1;
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t r1 = zs + (uint32_t)(8 * j1 + lr) * 128u, r2 = zs + (uint32_t)(8 * j2 + lr) * 128u;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t o = (uint32_t)(kk >> 1) * sub_bytes + ((kk & 1) ? sx1 : sx0);
        uint32_t b0, b1, b2, b3, c0, c1, c2, c3;
        ldsm_x4(r1 + o, b0, b1, b2, b3);
        ldsm_x4(r2 + o, c0, c1, c2, c3);
        mma_16816(acc, qa0[2 * kk], qa2[2 * kk], b0, b1);
        mma_16816(acc2, qa0[2 * kk], qa2[2 * kk], c0, c1);
        mma_16816(acc, qa0[2 * kk + 1], qa2[2 * kk + 1], b2, b3);
        mma_16816(acc2, qa0[2 * kk + 1], qa2[2 * kk + 1], c2, c3);
      }
      // softmax in registers: this thread holds head g, keys k1, k1+1 (tile j1) and k2, k2+1 (tile j2)
      const int k1 = 8 * j1 + 2 * t, k2 = 8 * j2 + 2 * t;
      const float s0 = (one && k1 < T) ? acc[0] : -INFINITY, s1 = (one && k1 + 1 < T) ? acc[1] : -INFINITY;
      const float s2 = (two && k2 < T) ? acc2[0] : -INFINITY, s3 = (two && k2 + 1 < T) ? acc2[1] : -INFINITY;
      float wm = fmaxf(fmaxf(s0, s1), fmaxf(s2, s3));
      wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, 1));
      wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, 2));
      if (t == 0) wmax[g * NWARP + warp] = wm;
      __syncthreads();
      const float4 ma = *reinterpret_cast<const float4*>(wmax + g * NWARP);
      const float4 mc = *reinterpret_cast<const float4*>(wmax + g * NWARP + 4);
      const float m = fmaxf(fmaxf(fmaxf(ma.x, ma.y), fmaxf(ma.z, ma.w)), fmaxf(fmaxf(mc.x, mc.y), fmaxf(mc.z, mc.w)));
      const float mb = m * LOG2E;
      const float p0 = exp2f(fmaf(s0, LOG2E, -mb)), p1 = exp2f(fmaf(s1, LOG2E, -mb));
      const float p2 = exp2f(fmaf(s2, LOG2E, -mb)), p3 = exp2f(fmaf(s3, LOG2E, -mb));
      float ps = (p0 + p1) + (p2 + p3);
      ps += __shfl_xor_sync(0xffffffffu, ps, 1);
      ps += __shfl_xor_sync(0xffffffffu, ps, 2);
      if (t == 0) wsum[g * NWARP + warp] = ps;
      if (g < H) {
        if (one) *reinterpret_cast<__nv_bfloat162*>(P + g * PLD + k1) = __floats2bfloat162_rn(p0, p1);
        if (two) *reinterpret_cast<__nv_bfloat162*>(P + g * PLD + k2) = __floats2bfloat162_rn(p2, p3);
      }
    }
    __syncthreads();

    // O = P Z: warp w owns latent dims 32w .. 32w+31 (chunks 4w .. 4w+3), all keys
    {
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      const __nv_bfloat16* prow = P + g * PLD + 2 * t;
      const int nks = Tp >> 4;
#pragma unroll 2
      for (int ks = 0; ks < nks; ++ks) {
        const uint32_t a0 = *reinterpret_cast<const uint32_t*>(prow + 16 * ks);
        const uint32_t a2 = *reinterpret_cast<const uint32_t*>(prow + 16 * ks + 8);
        const uint32_t zr = zs + (uint32_t)ks * 2048u;   // 16 keys x 128 B
        uint32_t b0, b1, b2, b3, c0, c1, c2, c3;
        ldsm_x4_t(zr + px[0], b0, b1, b2, b3);
        ldsm_x4_t(zr + px[1], c0, c1, c2, c3);
        mma_16816(acc[0], a0, a2, b0, b1);
        mma_16816(acc[1], a0, a2, b2, b3);
        mma_16816(acc[2], a0, a2, c0, c1);
        mma_16816(acc[3], a0, a2, c2, c3);
      }
      if (g < H) {
        const float4 a = *reinterpret_cast<const float4*>(wsum + g * NWARP);
        const float4 b = *reinterpret_cast<const float4*>(wsum + g * NWARP + 4);
        const float inv = 1.0f / (((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w)));
        __nv_bfloat16* orow = out + n * out_ld + (int64_t)g * ZD + 32 * warp + 2 * t;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<__nv_bfloat162*>(orow + 8 * i) = __floats2bfloat162_rn(acc[i][0] * inv, acc[i][1] * inv);
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// (rows, 256) bf16 contiguous, box {64 cols, box_rows}, 128-byte swizzle
int make_row_map(CUtensorMap* map, const void* base, int64_t rows, int box_rows) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      enc = (EncodeTiledFn)p;
  }
  YMT3_REQUIRE(enc, "cross_attn_absorbed: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)ZD, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ZD * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YMT3_REQUIRE(r == CUDA_SUCCESS, "cross_attn_absorbed: cuTensorMapEncodeTiled failed (%d) rows=%lld box=%d", (int)r,
               (long long)rows, box_rows);
  return YMT3_OK;
}

// (B, T, C, 256) -> (B*C, Tp, 256): rows t >= T of the destination are never written (zeroed at allocation)
__global__ void __launch_bounds__(256)
gather_latents_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t rows, int T, int C, int Tp) {
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);   // source row (b, t, c)
  if (i >= rows) return;
  const int lane = threadIdx.x & 31;
  const int c = (int)(i % C);
  const int64_t bt = i / C;
  const int tt = (int)(bt % T);
  const int64_t b = bt / T;
  dst[((b * C + c) * Tp + tt) * 32 + lane] = src[i * 32 + lane];
}

}  // namespace

namespace {
size_t absorbed_smem(int Tp, int stages) {
  return (size_t)stages * (Q_STAGE_BYTES + (size_t)Tp * 512) + (size_t)MAX_H * (Tp + 8) * 2 + 2 * MAX_H * NWARP * 4 +
         8 * 4;
}
}  // namespace

int cross_attn_absorbed(const void* q, int64_t q_ld, const void* z, void* out, int64_t out_ld, int64_t N, int H, int T,
                        int Tp, int zdim, cudaStream_t stream) {
  if (N <= 0) return YMT3_OK;
  YMT3_REQUIRE(zdim == ZD, "cross_attn_absorbed: latent width must be %d (got %d)", ZD, zdim);
  YMT3_REQUIRE(H >= 1 && H <= MAX_H, "cross_attn_absorbed: 1..%d heads (got %d)", MAX_H, H);
  YMT3_REQUIRE(T >= 1 && Tp >= T && Tp % 16 == 0 && Tp <= 128,
               "cross_attn_absorbed: encoder length %d (padded %d) outside 1..128", T, Tp);
  YMT3_REQUIRE(q_ld == (int64_t)H * ZD && out_ld % 8 == 0, "cross_attn_absorbed: q must be contiguous (N, H*%d)", ZD);
  YMT3_REQUIRE(N * (int64_t)(Tp > H ? Tp : H) < (1ll << 31), "cross_attn_absorbed: too many rows");
  const size_t max_smem = 227 * 1024 - 1024;   // 1 KB slack for the 1024-byte alignment of the dynamic segment
  // pipeline shape: `stages` tiles per CTA x `ctas` persistent CTAs per SM.  Default 1 x 3: three co-resident CTAs
  // (65 KB of shared memory, 72 registers per thread) each hold ONE tile and the hardware interleaves their phases -
  // while one waits for its TMA load the others run their score / softmax / P.Z phases, which a single CTA with a
  // 3-deep ring cannot overlap (its 8 warps move through the phases in lock step).  Measured (9464 sequences):
  // 3 x 1 121 us = 4.96 TB/s, 1 x 3 103.5 us = 5.80 TB/s (89 % of measured HBM), 1 x 2 120 us, 2 x 1 122 us
  // (profiles/r01_ab_xattn_shape.txt).  YMT3_XATTN_STAGES / YMT3_XATTN_CTAS override (A/B aids).
  static const int env_stages = getenv("YMT3_XATTN_STAGES") ? atoi(getenv("YMT3_XATTN_STAGES")) : 0;
  static const int env_ctas = getenv("YMT3_XATTN_CTAS") ? atoi(getenv("YMT3_XATTN_CTAS")) : 0;
  int stages = env_stages >= 1 && env_stages <= 3 ? env_stages : 1;
  const int ctas = env_ctas >= 1 && env_ctas <= 3 ? env_ctas : 3;
  while (stages > 1 && (absorbed_smem(Tp, stages) + 1024) * ctas > max_smem + 1024 - 2048 * (ctas - 1)) --stages;
  const size_t smem = absorbed_smem(Tp, stages) + 1024;
  static size_t configured[64] = {0};   // per device: largest dynamic smem opted into so far
  int dev = 0;
  YMT3_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || smem > configured[dev]) {
    YMT3_CUDA_CHECK(cudaFuncSetAttribute(cross_attn_absorbed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) configured[dev] = smem;
  }
  CUtensorMap mq, mz;
  int rc;
  if ((rc = make_row_map(&mq, q, N * H, H))) return rc;
  if ((rc = make_row_map(&mz, z, N * Tp, Tp))) return rc;
  const int64_t max_grid = (int64_t)ymt3_num_sms() * ctas;
  const int64_t grid = N < max_grid ? N : max_grid;
  YMT3_CUDA_CHECK(ymt3_launch_pdl(cross_attn_absorbed_kernel, dim3((unsigned)grid), dim3(NTHREAD), smem, stream, mq, mz,
                                  (__nv_bfloat16*)out, out_ld, N, H, T, Tp, stages));
  return YMT3_OK;
}

int gather_latents(const void* src, void* dst, int64_t B, int T, int C, int Tp, int zdim, cudaStream_t stream) {
  YMT3_REQUIRE(zdim == ZD, "gather_latents: latent width must be %d (got %d)", ZD, zdim);
  const int64_t rows = B * T * C;
  if (rows <= 0) return YMT3_OK;
  gather_latents_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>((const uint4*)src, (uint4*)dst, rows, T, C, Tp);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

}  // namespace ymt3
