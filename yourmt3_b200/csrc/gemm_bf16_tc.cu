// bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged
// by TMA with 128-byte swizzle) with the fused epilogue of ops.cuh::GemmParams.
//
//   C[M, N] = residual + out_scale * row_scale * epi(A[M, K] @ W[N, K]^T + bias)
//
// Persistent CTA (one per SM) = 64 + 32 * EPI_WARPS threads walking the tile list:
//   warp 0     TMA producer  (one elected lane): A box {64 k, 128 m}, W box {64 k, BN n}, deep smem ring
//   warp 1     TMEM allocator + MMA issuer (one elected lane): 4 x tcgen05.mma (K=16) per k-block
//   warps 2..  epilogue (EPI_GROUPS warps per TMEM lane quadrant, interleaved 32-column chunks):
//              tcgen05.ld -> registers -> compile-time specialised epilogue math -> 16-byte stores
// The accumulator is double-buffered in TMEM (2 x BN columns): the epilogue of tile i overlaps the
// MMAs of tile i+1.
//
// Contents of this translation unit (everything that shares the PTX helpers and the operand layouts):
//   1. PTX helpers: mbarrier, TMA load / store, tcgen05 alloc / mma / commit / ld, descriptors
//   2. epilogue math: general path (epi_math / epi_pack_bf16 / epi_store) and the LEAN path (lean_half, lean_half_gated,
//      lean_tile: compile-time flags, pipelined 16-column TMEM halves - the instances every model GEMM runs)
//   3. gemm_bf16_tc_kernel<BN, CONV, EPI, CL>: the persistent GEMM / implicit-GEMM convolution (CL = 2: CTA pair)
//   4. moe_expert_fused_kernel<ACT>: both expert GEMMs of the MoE feed-forward, hidden tile kept in shared memory
//   5. gemm_chain_kernel: several dependent decode-step GEMMs in one persistent launch (opt-in)
//   6. host side: tensor maps, tile-width / cluster rules, instance dispatch (gemm_bf16_tc, conv3x3_bf16_tc,
//      gemm_chain_bf16, moe_expert_fused)
#include "ops.cuh"
#include <cuda.h>
#include <cstring>
#include <cstdlib>

namespace ymt3 {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;

// epilogue warps: EPI_GROUPS warps per TMEM lane quadrant, each owning an adjacent share of the tile's 32-column
// chunks.  (First reading of the slow epilogue - ncu on the K = 128 gated GEMM, profiles/r02_gemm_moe1_gated_silu_
// ncu_full.txt: ~1400 warp instructions per warp and 128 x 256 tile at ~0.22 IPC - was "latency-bound per warp, add
// warps"; the per-CTA timeline later located the cost in uniform-datapath parameter / branch chains, see lean_tile.)
// MEASURED with EPI_GROUPS = 4 (16 epilogue warps, 576 threads): the register cap drops to 96
// per thread, every epilogue spills (100-680 bytes) and the K = 128 gated GEMM goes 252 -> 380 us, the bench 1324 ->
// 1416 ms (profiles/r02_experiments_not_kept.md); two groups (8 warps, 168 registers, no spills) stay.
constexpr int EPI_GROUPS = 2;
constexpr int EPI_WARPS = 4 * EPI_GROUPS;
constexpr int EPI_THREADS = 32 * EPI_WARPS;
constexpr int THREADS = 64 + EPI_THREADS;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// multicast variant: the box lands at the same CTA-relative offset in every CTA of `mask`, each destination's
// mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA store of one staged box (shared -> global), bulk async-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {   // <= N groups may still be READING their smem source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
// K-major operand, 128-byte swizzle, rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);  // start address (>>4), 14 bits
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// arrive on the barrier at this CTA-relative offset in EVERY CTA of `mask` once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// ---- CTA pair (cta_group::2): one tcgen05.mma spans the two SMs of a cluster -----------------------------------------
// shared::cluster address of `p` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
// TMA loads of a CTA pair: the data lands in THIS CTA's shared memory, the bytes are signalled on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t leader_bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* map, uint32_t leader_bar, void* dst, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once the pair's MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_C;\n\t"
      "bra WAIT_LOOP_C;\n\t"
      "DONE_C:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 16 accumulator columns, NOT waited for: the caller overlaps other work and then calls tmem_wait_ld on the same array
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// wait for every outstanding TMEM load of this thread; the registers are read-write operands so that no use of them
// can be scheduled above the wait
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// debug timeline (tools/trace_chain.py, ymt3_debug_chain_trace): events < 8 in globaltimer ns, the rest in SM cycles
static unsigned long long* g_chain_trace = nullptr;
__device__ __forceinline__ void chain_stamp(unsigned long long* trace, int j, int e) {
  if (trace && j < 16) {
    unsigned long long t;
    if (e < 8) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    else t = (unsigned long long)clock64();   // events 8.. : SM cycles (one warp's epilogue, finer than the 256 ns timer)
    trace[((size_t)blockIdx.x * 16 + j) * 32 + e] = t;
  }
}
#ifdef YMT3_GEMM_TRACE   // stamps in the single-GEMM kernel too (special build for tools/trace_gemm.py only)
#define TC_STAMP(j, e) chain_stamp(p.trace, j, e)
#else
#define TC_STAMP(j, e) ((void)0)
#endif

struct TcParams {
  void* C; int64_t ldc;
  const float* bias;
  const void* residual; int64_t ldr;
  int M, N, K;
  int act, gated;
  float out_scale;
  const float* row_scale;
  const float* norm_ss_in; int norm_ss_chunks; float norm_eps;   // fused RMSNorm, consumer side (ops.cuh)
  float* ss_out; int ss_out_chunks;                               // fused RMSNorm, producer side
#ifdef YMT3_GEMM_TRACE
  unsigned long long* trace;
#endif
  unsigned long long* argmax_out; int argmax_n;                   // fused greedy selection (ops.cuh)
  int tma_store;   // bf16 output leaves through per-warp smem staging + TMA stores (mapC) instead of 16-byte st.global
  const int* group_offsets;
  int num_groups;
  int out_f32;
  // implicit-GEMM 3x3 convolution mode (CONV template flag): A is an NHWC activation (B, T, F, Cin) read
  // through a 4-D tensor map {Cin, F, T, B}; row m = ((b*T + t)*F + f); k-block kb = tap * (Cin/64) + cb.
  int conv_T, conv_F, conv_cblocks;
};


// ---------------- epilogue math, specialised at compile time (keeps the hot path a few hundred
// straight-line instructions: the runtime-switch version thrashed the instruction cache) -------------
__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int ACT>
__device__ __forceinline__ float act_fast(float x) {
  if constexpr (ACT == YMT3_ACT_GELU_NEW) return 0.5f * x * (1.0f + fast_tanh(0.7978845608028654f * (x + 0.044715f * x * x * x)));
  else if constexpr (ACT == YMT3_ACT_RELU) return fmaxf(x, 0.f);
  else if constexpr (ACT == YMT3_ACT_SILU) {
    // x * 1 / (1 + 2^(-x log2 e)) with the two MUFU approximations issued directly: 3 FMUL + FADD + 2 MUFU per element
    // (__fdividef(x, 1 + __expf(-x)) spends ~11 instructions on range fix-ups that cannot trigger here: the
    // denominator is >= 1, and +inf gives 0)
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
  }
  else if constexpr (ACT == YMT3_ACT_GELU) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
  else return x;
}
// f[0..31] = accumulators (+bias). Non-gated: f[j] = act(f[j]) * rs (32 outputs). Gated: f[j] = act(f[2j]) * f[2j+1] * rs (16).
template <int ACT, bool GATED>
__device__ __forceinline__ void epi_math(float (&f)[32], float rs) {
  if constexpr (GATED) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = act_fast<ACT>(f[2 * j]) * f[2 * j + 1] * rs;
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = act_fast<ACT>(f[j]) * rs;
  }
}
__device__ __forceinline__ void epi_dispatch(float (&f)[32], float rs, int act, bool gated) {
  switch (act * 2 + (gated ? 1 : 0)) {   // warp-uniform
    case YMT3_ACT_NONE * 2 + 0: epi_math<YMT3_ACT_NONE, false>(f, rs); break;
    case YMT3_ACT_NONE * 2 + 1: epi_math<YMT3_ACT_NONE, true>(f, rs); break;
    case YMT3_ACT_GELU_NEW * 2 + 0: epi_math<YMT3_ACT_GELU_NEW, false>(f, rs); break;
    case YMT3_ACT_GELU_NEW * 2 + 1: epi_math<YMT3_ACT_GELU_NEW, true>(f, rs); break;
    case YMT3_ACT_RELU * 2 + 0: epi_math<YMT3_ACT_RELU, false>(f, rs); break;
    case YMT3_ACT_RELU * 2 + 1: epi_math<YMT3_ACT_RELU, true>(f, rs); break;
    case YMT3_ACT_SILU * 2 + 0: epi_math<YMT3_ACT_SILU, false>(f, rs); break;
    case YMT3_ACT_SILU * 2 + 1: epi_math<YMT3_ACT_SILU, true>(f, rs); break;
    case YMT3_ACT_GELU * 2 + 0: epi_math<YMT3_ACT_GELU, false>(f, rs); break;
    default: epi_math<YMT3_ACT_GELU, true>(f, rs); break;
  }
}
// round `count` (multiple of 8, <= 32) outputs (+ residual R) of one row to bf16 and hand every 16-byte unit j to
// sink(j, unit) as soon as it is packed; returns the sum of squares of the ROUNDED values when want_ss (fused RMSNorm
// producer)
template <typename Sink>
__device__ __forceinline__ float epi_pack_bf16(const __nv_bfloat16* R, const float (&f)[32], int count, bool want_ss,
                                               Sink&& sink) {
  uint4 r[4];
  if (R) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (8 * j < count) r[j] = *reinterpret_cast<const uint4*>(R + 8 * j);   // all residual loads in flight first
  }
  float sqx = 0.f, sqy = 0.f;   // even / odd columns apart: the order of epi_chunk_lean's packed accumulator
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (8 * j >= count) break;
    uint4 pk;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
    if (R) {
      const __nv_bfloat162* th = reinterpret_cast<const __nv_bfloat162*>(&r[j]);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        h[q] = __floats2bfloat162_rn(f[8 * j + 2 * q] + __bfloat162float(th[q].x),
                                     f[8 * j + 2 * q + 1] + __bfloat162float(th[q].y));
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(f[8 * j + 2 * q], f[8 * j + 2 * q + 1]);
    }
    if (want_ss) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 v2 = __bfloat1622float2(h[q]);
        sqx = fmaf(v2.x, v2.x, sqx);
        sqy = fmaf(v2.y, v2.y, sqy);
      }
    }
    sink(j, pk);
  }
  return sqx + sqy;
}

// ---- LEAN epilogue: the common case of the bf16 epilogue (full 32-column chunks of a non-gated, activation-free tile that
// leaves through the TMA staging box) with the per-element work cut to the bone.  The general path spends ~290
// warp instructions per chunk (timeline + SASS counts: profiles/r02_gemm_epilogue_timeline.txt: two scalings, scalar
// adds / conversions, ~70 control-flow instructions), this one ~60 without and ~150 with residual + sum of squares:
// packed f32x2 multiply / add / fma, one scale factor, no per-element predicates.  Same values as the general path
// (IEEE mul / add / fma per element, identical summation order of the squares).
__device__ __forceinline__ float2 bf16x2_as_f2(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ float2 lean_mul2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  return __fmul2_rn(a, b);
#else
  return make_float2(a.x * b.x, a.y * b.y);
#endif
}
__device__ __forceinline__ float2 lean_add2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  return __fadd2_rn(a, b);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__device__ __forceinline__ float2 lean_fma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  return __ffma2_rn(a, b, c);
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
// v: 16 accumulator columns of this thread's row (one HALF of a 32-column chunk).  out = bf16(v * s + bias + residual)
// (s = the one scale factor of the general path that is not exactly 1).  rr: the 32 residual bytes of this row segment.
// ssacc: running (even, odd column) sums of squares of the rounded outputs of the chunk.  dst / u0 / sx: staging row,
// first 16-byte unit (compile-time after unrolling), swizzle.
// (BIAS / RES / SS: compile-time constants in the single-GEMM instances, warp-uniform register predicates in the chain)
__device__ __forceinline__ void lean_half(const bool BIAS, const bool RES, const bool SS, const uint32_t (&v)[16], float s,
                                          const float* bias_c, const uint4 (&rr)[2], float2& ssacc, uint32_t dst, int u0, int sx) {
  float2 f[8];
  const float2 a1 = make_float2(s, s);
#pragma unroll
  for (int q = 0; q < 8; ++q) f[q] = lean_mul2(make_float2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])), a1);
  if (BIAS) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias_c) + q);
      f[2 * q] = lean_add2(f[2 * q], make_float2(b.x, b.y));
      f[2 * q + 1] = lean_add2(f[2 * q + 1], make_float2(b.z, b.w));
    }
  }
  if (RES) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&rr[0]);
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] = lean_add2(f[q], bf16x2_as_f2(w[q]));
  }
  uint32_t h[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const __nv_bfloat162 b2 = __floats2bfloat162_rn(f[q].x, f[q].y);
    h[q] = *reinterpret_cast<const uint32_t*>(&b2);
  }
  if (SS) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float2 u = bf16x2_as_f2(h[q]);
      ssacc = lean_fma2(u, u, ssacc);
    }
  }
#pragma unroll
  for (int q = 0; q < 2; ++q)   // dst: shared-space address of the staging row (st.shared, not a generic store)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)(((u0 + q) << 4) ^ sx)), "r"(h[4 * q]),
                 "r"(h[4 * q + 1]), "r"(h[4 * q + 2]), "r"(h[4 * q + 3])
                 : "memory");
}
// gated form: v = 16 accumulator columns = 8 (activated, linear) pairs; h = 8 bf16 outputs act(a_even) * a_odd * rs with
// a = v * pre + bias - the per-element operations of epi_math<ACT, true>
// (SCALE false: pre and rs are known to be exactly 1 and are not applied)
template <int ACT, bool BIAS, bool SCALE = true>
__device__ __forceinline__ void lean_half_gated(const uint32_t (&v)[16], float pre, float rs, const float* bias_c, uint32_t (&h)[4]) {
  float2 a[8];
  const float2 p2 = make_float2(pre, pre);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    a[q] = make_float2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
    if constexpr (SCALE) a[q] = lean_mul2(a[q], p2);
  }
  if constexpr (BIAS) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias_c) + q);
      a[2 * q] = lean_add2(a[2 * q], make_float2(b.x, b.y));
      a[2 * q + 1] = lean_add2(a[2 * q + 1], make_float2(b.z, b.w));
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float o0 = act_fast<ACT>(a[2 * q].x) * a[2 * q].y;
    float o1 = act_fast<ACT>(a[2 * q + 1].x) * a[2 * q + 1].y;
    if constexpr (SCALE) {
      o0 *= rs;
      o1 *= rs;
    }
    const __nv_bfloat162 b2 = __floats2bfloat162_rn(o0, o1);
    h[q] = *reinterpret_cast<const uint32_t*>(&b2);
  }
}
// the residual bytes of one row segment of 16 bf16 - issued early, consumed by lean_half
template <bool L2_ONLY>
__device__ __forceinline__ void lean_load_res16(uint4 (&rr)[2], const __nv_bfloat16* R) {
#pragma unroll
  for (int q = 0; q < 2; ++q)
    rr[q] = L2_ONLY ? __ldcg(reinterpret_cast<const uint4*>(R) + q) : *(reinterpret_cast<const uint4*>(R) + q);
}
// store `count` (multiple of 8 for bf16 / 4 for f32, <= 32) consecutive outputs f[0..count) of one row (+ residual)
// ss != null (bf16 output only): *ss = sum of squares of the bf16-rounded values stored (fused RMSNorm producer)
template <bool OUT_F32>
__device__ __forceinline__ void epi_store(void* Cbase, const void* Rbase, int64_t off, const float (&f)[32], int count,
                                          float* ss = nullptr) {
  if constexpr (OUT_F32) {
    float* C = static_cast<float*>(Cbase) + off;
    const float* R = Rbase ? static_cast<const float*>(Rbase) + off : nullptr;
    float4 r[8];
    if (R) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (4 * j < count) r[j] = *reinterpret_cast<const float4*>(R + 4 * j);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (4 * j >= count) break;
      float4 q = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
      if (R) { q.x += r[j].x; q.y += r[j].y; q.z += r[j].z; q.w += r[j].w; }
      *reinterpret_cast<float4*>(C + 4 * j) = q;
    }
  } else {
    __nv_bfloat16* C = static_cast<__nv_bfloat16*>(Cbase) + off;
    const __nv_bfloat16* R = Rbase ? static_cast<const __nv_bfloat16*>(Rbase) + off : nullptr;
    const float sq = epi_pack_bf16(R, f, count, ss != nullptr,
                                   [&](int j, const uint4& pk) { *reinterpret_cast<uint4*>(C + 8 * j) = pk; });
    if (ss) *ss = sq;
  }
}

template <int BN, int CL = 1>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;   // 16 KB
  static constexpr int B_BYTES = (BN / CL) * BK * 2;   // pair: each CTA stages half of the W tile -> a deeper ring
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // epilogue staging for the TMA stores: EPI_WARPS warps x STG_BOXES boxes of (32 rows x <= 128 B), 1024-byte aligned
  static constexpr int CPW = BN / (32 * EPI_GROUPS) > 0 ? BN / (32 * EPI_GROUPS) : 1;   // 32-column chunks per warp
  // ONE 4 KB box per warp: a BN = 256 tile (two boxes per warp) hands them to the TMA engine one after the other
  // through the same buffer - the 32 KB this saves buy the single-CTA 128 x 256 kernel a 4th ring stage
  static constexpr int STG_BOXES = 1;
  static constexpr int STG_WARP_BYTES = STG_BOXES * 4096;
  static constexpr int STG_BYTES = EPI_WARPS * STG_WARP_BYTES;
  static constexpr int RING_BUDGET = 226 * 1024 - 2048 - STG_BYTES;
  static constexpr int STAGES = (RING_BUDGET / STAGE_BYTES) > 8 ? 8 : (RING_BUDGET / STAGE_BYTES);
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int STG_OFF = BAR_OFF + 1024;
  static constexpr int TOTAL = STG_OFF + STG_BYTES + 1024;  // barriers / tmem slot / group table, staging, alignment slack
  static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- context of one tile for lean_tile (everything it needs from the kernel around it)
struct LeanCtx {
  const CUtensorMap* mapC;
  int N; const void* residual; int64_t ldr; float* ss_out; int ss_out_chunks; void* C; int64_t ldc;
  const float* bias;
  int m0, row_end, n0, r; bool row_ok, warp_tma; float pre, rs;
  uint32_t tmem_base; int buf; uint64_t* full_bar; uint64_t* empty_bar; uint32_t full_parity;
  uint8_t* stg; int quad, half, lane, j; bool stamp; unsigned long long* trace;
  int flags;   // EPI == 72 only: bias | residual << 1 | sum of squares << 2
};
#define LEAN_STAMP(j, e) chain_stamp(c.trace, j, e)
// ---- LEAN tile epilogue (lean_half): the warp's CPW * 32 columns as 16-column halves, software-pipelined: the
// TMEM load of half g + 1 is in flight while half g is scaled / packed / staged, and the residual bytes are
// fetched three halves ahead (the first three before the accumulator is complete).  Bias / residual / sum of
// squares are COMPILE-TIME properties of the instance and everything that depends on kernel parameters is
// computed once per tile: the general path re-reads parameters through the uniform datapath and branches on
// them ~10 times per chunk (LDCU -> UISETP -> BRA chains, the `branch_resolving` / `wait` stalls of the ncu
// source view), which - not TMEM or shared-memory traffic - is what a 32-column chunk's ~1000 clk were made of
// (leave-out experiments + timeline: profiles/r02_gemm_epilogue_timeline.txt).
// EPI: 64 + flags (plain) / 80 + flags (gated) as in the kernel; L2: residual reads bypass L1 and start only after the
// accumulator barrier (chained launches: the rows were written earlier in the SAME launch, possibly by another SM).
template <int BN, int EPI, int CL, bool L2>
__device__ __forceinline__ void lean_tile(const LeanCtx& c) {
  constexpr bool LG = EPI >= 80;
  constexpr int CPW = SmemLayout<BN, CL>::CPW;
  constexpr int NCHUNK = BN / 32;
  const LeanCtx& p = c;
  const CUtensorMap& mapC = *c.mapC;
  const int m0 = c.m0, row_end = c.row_end, n0 = c.n0, r = c.r, quad = c.quad, half = c.half, lane = c.lane, buf = c.buf;
  [[maybe_unused]] const int j = c.j;
  const bool row_ok = c.row_ok, warp_tma = c.warp_tma;
  const float pre = c.pre, rs = c.rs;
  [[maybe_unused]] const float* bias = c.bias;
  const uint32_t tmem_base = c.tmem_base;
  uint8_t* stg = c.stg;
  constexpr int H = 2 * CPW;                                   // halves per warp
  constexpr int SRB = LG ? (CPW * 32 > 128 ? 128 : CPW * 32)   // bytes per staged box row (gated: 8 outputs per half)
                         : (CPW * 64 > 128 ? 128 : CPW * 64);
  const int sxor = tma_swizzle_xor(lane, SRB);   // index_maps.h
  // bias, residual, sum of squares: compile-time, or (EPI == 72: the chained kernel's plain phases) run-time flags that
  // live in predicate registers for the whole tile
  constexpr bool RT = EPI == 72;
  const bool LB = RT ? (c.flags & 1) != 0 : (EPI & 1) != 0;
  const bool LR = RT ? (c.flags & 2) != 0 : (!LG && (EPI & 2) != 0);
  const bool LS = RT ? (c.flags & 4) != 0 : (!LG && (EPI & 4) != 0);
  constexpr int G_ACT = (EPI & 2) ? YMT3_ACT_SILU : YMT3_ACT_GELU_NEW;   // (gated instances)
  const int c_first = n0 + half * CPW * 32;                    // first accumulator column of this warp
  const int n_cols = p.N;
  // a ragged last slab of a group (grouped gated GEMM) cannot leave through the box store (it would spill into the
  // next group's rows): direct 16-byte stores for its rows
  const bool slab_rows = m0 + quad * 32 < row_end;             // this warp's slab has rows in the tile at all
  [[maybe_unused]] const bool slab_direct = LG && slab_rows && !warp_tma;
  // number of this warp's halves that hold real columns (N % 32 == 0: always even); 0: nothing to do
  const int nv = (slab_rows && (warp_tma || slab_direct) && half * CPW < NCHUNK && c_first < n_cols)
      ? min(H, (n_cols - c_first) >> 4) : 0;
  // rows past the end of the tile: loads clamped to the last row, the box store clips them
  [[maybe_unused]] const __nv_bfloat16* Rrow = nullptr;
  [[maybe_unused]] uint4 rq[4][2];                             // residual ring: half g lives in rq[g & 3]
  if (LR) {
    Rrow = static_cast<const __nv_bfloat16*>(p.residual) + (int64_t)(row_ok ? r : row_end - 1) * p.ldr + c_first;
    if constexpr (!L2) {   // (chained launch: the rows may still be in flight from another SM until the accumulator barrier)
#pragma unroll
      for (int g = 0; g < 3 && g < H; ++g)
        if (g < nv) lean_load_res16<L2>(rq[g], Rrow + 16 * g);
    }
  }
  [[maybe_unused]] const float* bias_w = LB ? bias + c_first : nullptr;
  [[maybe_unused]] float* ss_w = LS ? p.ss_out + (int64_t)(row_ok ? r : row_end - 1) * p.ss_out_chunks + (c_first >> 5) : nullptr;
  const float s1 = LG ? pre : pre * rs;   // plain: the host selects this instance only when one factor is exactly 1
  [[maybe_unused]] __nv_bfloat16* Cdir = nullptr;   // gated, ragged slab: this thread's output row segment
  if constexpr (LG) {
    if (slab_direct) Cdir = static_cast<__nv_bfloat16*>(p.C) + (int64_t)(row_ok ? r : row_end - 1) * p.ldc + (c_first >> 1);
  }
  mbar_wait(c.full_bar, c.full_parity);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (c.stamp) LEAN_STAMP(j, 4);
  if (LR && L2) {
#pragma unroll
    for (int g = 0; g < 3 && g < H; ++g)
      if (g < nv) lean_load_res16<L2>(rq[g], Rrow + 16 * g);
  }
  const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN + half * CPW * 32);
  uint32_t va[16], vb[16];
  if (nv > 0) {
    tmem_ld16_nowait(t_addr, va);
    if (lane == 0) tma_store_wait_read<0>();   // the previous tile's box has left the staging buffer
    __syncwarp();
  } else {   // a warp without columns in this tile still releases the accumulator
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if constexpr (CL == 2) mbar_arrive_cluster(mapa_u32(c.empty_bar, 0));
    else mbar_arrive(c.empty_bar);
  }
  float2 ssacc = make_float2(0.f, 0.f);
  const uint32_t dst = smem_u32(stg) + (uint32_t)(lane * SRB);
#pragma unroll
  for (int g = 0; g < H; ++g) {
    if (g < nv) {                                             // warp-uniform
      const bool new_box = !LG && g > 0 && (g * 32) % SRB == 0;   // compile-time after unrolling
      if (new_box) {
        // this half opens the warp's next staging box: the finished one goes to the TMA engine; its read-out
        // overlaps the TMEM wait below
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&mapC, stg, c_first + ((g * 32) / SRB - 1) * (SRB >> 1), m0 + quad * 32);
          tma_store_commit();
        }
      }
      if (g & 1) tmem_wait_ld(vb); else tmem_wait_ld(va);      // half g is in registers
      if (g + 1 < nv) {
        if (g & 1) tmem_ld16_nowait(t_addr + 16 * (g + 1), va); else tmem_ld16_nowait(t_addr + 16 * (g + 1), vb);
      } else {
        // that was this warp's last TMEM load of the tile: hand the accumulator back to the MMA warp
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if constexpr (CL == 2) mbar_arrive_cluster(mapa_u32(c.empty_bar, 0));   // the LEADER's barrier
        else mbar_arrive(c.empty_bar);
      }
      if (LR) {
        if (g + 3 < nv) lean_load_res16<L2>(rq[(g + 3) & 3], Rrow + 16 * (g + 3));
      }
      if (new_box) {   // the buffer is rewritten only after the TMA engine has read the previous box
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
      }
      if constexpr (LG) {
        uint32_t h4[4];
        lean_half_gated<G_ACT, (EPI & 1) != 0>((g & 1) ? vb : va, s1, rs, LB ? bias_w + 16 * g : nullptr, h4);
        if (!slab_direct)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)((g << 4) ^ sxor)), "r"(h4[0]),
                       "r"(h4[1]), "r"(h4[2]), "r"(h4[3])
                       : "memory");
        else if (row_ok)
          *reinterpret_cast<uint4*>(Cdir + 8 * g) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
      } else {
        lean_half(LB, LR, LS, (g & 1) ? vb : va, s1, LB ? bias_w + 16 * g : nullptr, rq[g & 3], ssacc, dst,
                              ((g * 32) % SRB) >> 4, sxor);
      }
      if (LS) {
        if (g & 1) {
          if (row_ok) ss_w[g >> 1] = ssacc.x + ssacc.y;
          ssacc = make_float2(0.f, 0.f);
        }
      }
    }
  }
  if (nv > 0 && !slab_direct) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      // the box this warp wrote last (earlier ones left inside the loop); columns >= N / rows >= M are clipped
      if constexpr (LG) tma_store_2d(&mapC, stg, c_first >> 1, m0 + quad * 32);
      else tma_store_2d(&mapC, stg, c_first + (((nv - 1) * 32) / SRB) * (SRB >> 1), m0 + quad * 32);
      tma_store_commit();
    }
  }
  if (c.stamp) LEAN_STAMP(j, 5);
}
#undef LEAN_STAMP

// Persistent kernel: one CTA per SM walks the tile list (tile = blockIdx.x + i * gridDim.x).  The TMA
// producer runs ahead across tiles through a STAGES-deep smem ring; the accumulator is DOUBLE-BUFFERED in
// TMEM (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1; barriers, TMEM allocation
// and tensor-map prefetch are paid once per CTA instead of once per tile (decisive for K = 128 GEMMs).
// EPI >= 0: epilogue specialised at compile time for (activation, gated, fp32 output) = (EPI >> 2, (EPI >> 1) & 1,
// EPI & 1) with the store path fixed (bf16 non-gated: TMA store; gated / fp32: direct) - the executed SASS path shrinks
// from ~4000 to ~1500 instructions (ncu on the generic kernel: `no_inst` / `branch_resolving` stalls all over the
// epilogue, 460 warp instructions per 32-column chunk; profiles/r01_gemm_smallk_v4_tma_store_ncu_full.txt).
// EPI < 0: every combination decided at run time (rare shapes, convolutions, A/B switches).
// CL == 2: CTA PAIR (`cta_group::2`, thread-block cluster of 2 along M).  The pair computes a 256 x BN super-tile with ONE
// stream of tcgen05.mma issued by the leader (rank 0): each CTA stages its own 128 rows of A and HALF of the W tile
// (BN / 2 rows); the tensor cores of both SMs read A from their own and the two W halves from both shared memories;
// each CTA's 128 x BN accumulator lives in its own TMEM and is drained by its own epilogue warps.  Why: with a
// single-CTA 128 x 256 tile every k-block costs the SM 48 KB of TMA writes + 48 KB of operand reads = 188 B / clk for
// 512 clk of MMA against ~128 B / clk of shared-memory bandwidth (= the 71 % of cuBLAS measured on 8192^3); the pair
// stages 32 KB and reads 32 KB per SM.  (The first round-2 attempt - same cluster, cta_group::1 MMAs, W tile
// MULTICAST into both CTAs - halves the L2 -> SM bytes but not the shared-memory traffic, and measured neutral.)
// Protocol: both producers signal the LEADER's full barrier (.cta_group::2 TMA loads, leader expects both CTAs'
// bytes); the leader commits (multicast to both CTAs) onto the slot's empty barrier and the accumulator's full
// barrier; the epilogue threads of BOTH CTAs arrive on the leader's accumulator-empty barrier.
template <int BN, bool CONV, int EPI, int CL>
__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                    const __grid_constant__ CUtensorMap mapC, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  using L = SmemLayout<BN, CL>;
  constexpr int STAGES = L::STAGES;
  // SWIZZLE_128B operand tiles need 1024-byte aligned bases
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  int* gstart = reinterpret_cast<int*>(tmem_slot + 2);   // [num_groups + 1] tile-index prefix (grouped mode)

  constexpr bool GEN = EPI < 0;
  // EPI_LEAN_BF16: plain bf16 output, N % 32 == 0, no groups / arg-max, TMA store: ONLY the lean chunk path is compiled
  // (epi_chunk_lean; the host picks it whenever the shape allows - every plain GEMM of the decode step and the encoder)
  constexpr bool LEAN = EPI >= 64;   // 64 + (bias ? 1 : 0) + (residual ? 2 : 0) + (sum of squares ? 4 : 0)
  constexpr bool LG = EPI >= 80;     // gated lean: 80 + (bias ? 1 : 0) + (SiLU ? 2 : 0) (else gelu_new)
  const int e_act = GEN ? p.act : (LEAN ? (LG ? ((EPI & 2) ? YMT3_ACT_SILU : YMT3_ACT_GELU_NEW) : 0) : (EPI >> 2));
  const bool e_gated = GEN ? (p.gated != 0) : (LEAN ? LG : ((EPI >> 1) & 1) != 0);
  const bool e_f32 = GEN ? (p.out_f32 != 0) : (!LEAN && (EPI & 1) != 0);
  const bool e_tma = GEN ? (p.tma_store != 0) : (LEAN || (!e_gated && !e_f32));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_kb = (p.K + BK - 1) / BK;
  const int groups = p.group_offsets ? p.num_groups : 1;

  // PDL: dependents may be scheduled now; this kernel's own prologue (barriers, TMEM, tensor maps) overlaps the
  // tail of the previous kernel, global memory is only touched after pdl_wait()
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    if (e_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapC) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);    // (pair: the leader's commit is multicast to both CTAs)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], EPI_THREADS * CL);   // all epilogue threads arrive (pair: of both CTAs, on the leader's)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    int acc = 0;
    if (p.group_offsets) {
      pdl_wait();   // the offsets are produced by the preceding kernel
      for (int g = 0; g < groups; ++g) {
        gstart[g] = acc;
        acc += ((p.group_offsets[g + 1] - p.group_offsets[g] + BM - 1) / BM) * n_tiles;
      }
    } else {
      gstart[0] = 0;
      // clusters walk SUPER tiles: CL vertically adjacent M tiles x one N tile
      acc = ((p.M + BM * CL - 1) / (BM * CL)) * n_tiles;
    }
    gstart[groups] = acc;
  }
  if (warp == 1) {
    if constexpr (CL == 2) {   // issued by the same logical warp of both CTAs of the pair
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(2 * BN)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(2 * BN)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();   // every CTA's barriers are initialised before any remote arrive / multicast
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) TC_STAMP(0, 7);
  const int total_tiles = gstart[groups];
  // persistent walk: tile (CL == 1) or super-tile (CL > 1) index t = t_first, t_first + t_step, ...
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int t_first = CL > 1 ? (int)(blockIdx.x / CL) : (int)blockIdx.x;
  const int t_step = CL > 1 ? (int)(gridDim.x / CL) : (int)gridDim.x;
  static_assert(CL == 1 || CL == 2, "cluster size 1 (single CTA) or 2 (CTA pair)");

  // tile index -> (m0, row_end, n0, w_row0, bias offset group)
  auto decode_tile = [&](int t, int& m0, int& row_end, int& n0, int& w_row0, int& g_out) {
    int g = 0;
    if (p.group_offsets) {
      while (g + 1 < groups && t >= gstart[g + 1]) ++g;
    }
    const int local = t - gstart[g];
    const int mt = local / n_tiles, nt = local - mt * n_tiles;
    n0 = nt * BN;
    if (p.group_offsets) {
      m0 = p.group_offsets[g] + mt * BM;
      row_end = p.group_offsets[g + 1];
      w_row0 = g * p.N + n0;
    } else {
      m0 = (mt * CL + crank) * BM;   // may lie entirely beyond M in the last super-tile: TMA zero-fills, nothing is stored
      row_end = p.M;
      w_row0 = n0;
    }
    g_out = g;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_first; t < total_tiles; t += t_step) {
        int m0, row_end, n0, w_row0, g;
        decode_tile(t, m0, row_end, n0, w_row0, g);
        int cv_f0 = 0, cv_t = 0, cv_b = 0;
        if constexpr (CONV) {
          const int tpr = p.conv_F / BM;            // M tiles per (b, t) row; conv_F is a multiple of 128
          const int mt = m0 / BM;
          cv_f0 = (mt % tpr) * BM;
          const int bt = mt / tpr;
          cv_t = bt % p.conv_T;
          cv_b = bt / p.conv_T;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if constexpr (CL == 2) {
            // pair: own A tile + own half of the W tile (mapW's box is BN / 2 rows) into OWN shared memory, bytes
            // signalled on the leader's barrier, which expects both CTAs' shares
            const uint32_t lbar = mapa_u32(&full_bar[stage], 0);
            if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
            if constexpr (CONV) {
              const int tap = kb / p.conv_cblocks, cb = kb - tap * p.conv_cblocks;
              const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
              tma_load_4d_pair(&mapA, lbar, sa, cb * BK, cv_f0 + dx, cv_t + dy, cv_b);
            } else {
              tma_load_2d_pair(&mapA, lbar, sa, kb * BK, m0);
            }
            tma_load_2d_pair(&mapW, lbar, sb, kb * BK, w_row0 + crank * (BN / 2));
          } else {
            mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
            if constexpr (CONV) {
              // shifted window of the input; out-of-range rows/cols are zero-filled by TMA = conv zero padding
              const int tap = kb / p.conv_cblocks, cb = kb - tap * p.conv_cblocks;
              const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
              tma_load_4d(&mapA, &full_bar[stage], sa, cb * BK, cv_f0 + dx, cv_t + dy, cv_b);
            } else {
              tma_load_2d(&mapA, &full_bar[stage], sa, kb * BK, m0);
            }
            tma_load_2d(&mapW, &full_bar[stage], sb, kb * BK, w_row0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && (CL == 1 || crank == 0)) {
    // ===================== MMA issuer (pair: the leader CTA only) =====================
    // instruction descriptor: D=f32, A=B=bf16, both K-major, N = BN, M = 128 (pair: 256 = both CTAs' rows)
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CL) >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int j = 0;
    for (int t = t_first; t < total_tiles; t += t_step, ++j) {
      const int buf = j & 1;
      // epilogue drained this accumulator buffer (pair: the epilogues of BOTH CTAs, arriving from across the cluster)
      if constexpr (CL == 2) mbar_wait_cluster(&tmem_empty_bar[buf], ((j >> 1) & 1) ^ 1);
      else mbar_wait(&tmem_empty_bar[buf], ((j >> 1) & 1) ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint8_t* sa = smem + stage * L::STAGE_BYTES;
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + L::A_BYTES);
          if (kb == 0) TC_STAMP(j, 2);
          if constexpr (CL == 2) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_pair(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage]);                          // frees the slot in BOTH CTAs
            if (kb == num_kb - 1) umma_commit_pair(&tmem_full_bar[buf]);  // accumulators complete -> both epilogues
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)  // +32 B along K inside the 128 B swizzle row = +2 in the address field
              umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);                          // frees the smem slot when the MMAs retire
            if (kb == num_kb - 1) {
              umma_commit(&tmem_full_bar[buf]);  // accumulator complete -> epilogue
              TC_STAMP(j, 3);
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 2) {
    // ===================== epilogue (warps 2..9) =====================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;          // which share of the tile's 32-column chunks (0 .. EPI_GROUPS-1)
    // Each warp owns CPW ADJACENT 32-column chunks of the tile (BN / (32 * EPI_GROUPS), at least one).
    // TMA-store path: the warp stages its 32 rows x (CPW * 64 B, gated: CPW * 32 B) of bf16 output in shared memory
    // in the swizzle of mapC (Swizzle<log2(row bytes / 16), 4, 3>: conflict-free 16-byte writes) and one lane issues
    // one box store (two for BN = 256: a box row is at most the 128-byte swizzle span) per tile: full 32/64/128-byte
    // row segments instead of 32 scattered 16-byte st.global per warp instruction.  A single buffer per warp
    // suffices: its previous store is a whole tile old.
    constexpr int CPW = L::CPW;
    uint8_t* stg = smem + L::STG_OFF + (warp - 2) * L::STG_WARP_BYTES;
    const int stg_rb_all = CPW * (e_gated ? 32 : 64);              // staged bytes per row, all boxes
    const int stg_rb = stg_rb_all > 128 ? 128 : stg_rb_all;        // bytes per row of one box
    const int stg_row = lane * stg_rb;
    const int stg_xor = tma_swizzle_xor(lane, stg_rb);   // index_maps.h
    int j = 0;
    for (int t = t_first; t < total_tiles; t += t_step, ++j) {
      int m0, row_end, n0, w_row0, g;
      decode_tile(t, m0, row_end, n0, w_row0, g);
      const float* bias = p.bias ? p.bias + (p.group_offsets ? (int64_t)g * p.N : 0) : nullptr;
      const int buf = j & 1;
      const int r = m0 + quad * 32 + lane;       // output row of this thread
      const bool row_ok = r < row_end;
      // whole 32-row slab inside the tile's row range (always true without groups: TMA clips rows >= M itself;
      // a group's ragged last slab must not spill into the next group's rows -> direct stores there)
      const bool warp_tma = e_tma && m0 + quad * 32 < row_end && (!p.group_offsets || m0 + quad * 32 + 32 <= row_end);
      const float rs = p.out_scale * ((p.row_scale && row_ok) ? p.row_scale[r] : 1.0f);
      float pre = 1.0f;   // fused RMSNorm: r = rsqrt(mean(x^2) + eps) of this row of A, partials summed in a fixed order
      if (p.norm_ss_in && row_ok) {
        const float* sp = p.norm_ss_in + (int64_t)r * p.norm_ss_chunks;
        float tot = 0.f;
        for (int c4 = 0; c4 + 4 <= p.norm_ss_chunks; c4 += 4) {
          const float4 s4 = *reinterpret_cast<const float4*>(sp + c4);
          tot += (s4.x + s4.y) + (s4.z + s4.w);
        }
        for (int c1 = p.norm_ss_chunks & ~3; c1 < p.norm_ss_chunks; ++c1) tot += sp[c1];
        pre = rsqrtf(tot / (float)p.K + p.norm_eps);
      }
      constexpr int NCHUNK = BN / 32;
      if constexpr (LEAN) {
        LeanCtx lc;
        lc.mapC = &mapC;
        lc.N = p.N; lc.residual = p.residual; lc.ldr = p.ldr; lc.ss_out = p.ss_out; lc.ss_out_chunks = p.ss_out_chunks;
        lc.C = p.C; lc.ldc = p.ldc; lc.bias = bias;
        lc.m0 = m0; lc.row_end = row_end; lc.n0 = n0; lc.r = r; lc.row_ok = row_ok; lc.warp_tma = warp_tma; lc.pre = pre; lc.rs = rs;
        lc.tmem_base = tmem_base; lc.buf = buf; lc.full_bar = &tmem_full_bar[buf]; lc.empty_bar = &tmem_empty_bar[buf];
        lc.full_parity = (uint32_t)((j >> 1) & 1);
        lc.stg = stg; lc.quad = quad; lc.half = half; lc.lane = lane; lc.j = j; lc.stamp = false; lc.trace = nullptr; lc.flags = 0;
#ifdef YMT3_GEMM_TRACE
        lc.stamp = warp == 2 && lane == 0; lc.trace = p.trace;
#endif
        lean_tile<BN, EPI, CL, false>(lc);
        continue;
      }
      // (the per-row loads above do not depend on the accumulator: they are issued before the wait)
      mbar_wait(&tmem_full_bar[buf], (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      [[maybe_unused]] const bool stamp = warp == 2 && lane == 0;
      if (stamp) TC_STAMP(j, 4);
      unsigned long long best_key = 0;   // fused greedy selection: best (logit, column) this thread produced
      const int c_first = n0 + half * CPW * 32;          // first accumulator column of this warp
      const bool stage_any = warp_tma && half * CPW < NCHUNK && c_first < p.N;   // warp-uniform
      if (stage_any) {
        if (lane == 0) tma_store_wait_read<0>();   // the previous tile's box has left the staging buffer
        __syncwarp();
      }
#pragma unroll 1
      for (int k = 0; k < CPW; ++k) {
        // (a warp without a chunk - BN = 32, second half - still runs its round to reach the arrive below)
        const int ci = half * CPW + k;
        const bool has_chunk = ci < NCHUNK;
        const bool last = k == CPW - 1;         // this warp's last round for this tile
        uint32_t v[32];
        bool box_in_flight = false;
        if (stamp) TC_STAMP(j, 8 + 6 * k);
        __syncwarp();
        {
          // warp-uniform: this chunk opens the warp's NEXT staging box -> hand the finished one to the TMA engine
          const int b0k = k * (e_gated ? 32 : 64);
          if (stage_any && b0k > 0 && b0k % stg_rb == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              const int cb = (e_gated ? (c_first >> 1) : c_first) + (b0k / stg_rb - 1) * (stg_rb >> 1);
              if (cb < (e_gated ? (p.N >> 1) : p.N)) tma_store_2d(&mapC, stg, cb, m0 + quad * 32);
              tma_store_commit();
            }
            box_in_flight = true;   // the buffer may be rewritten only after wait_read (below, after this chunk's math)
          }
        }
        if (stamp) TC_STAMP(j, 9 + 6 * k);
        if (has_chunk) tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN + ci * 32), v);
        if (stamp) TC_STAMP(j, 10 + 6 * k);
        if (last) {
          // all TMEM reads of this warp for this tile are done: hand the buffer back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          if constexpr (CL == 2) mbar_arrive_cluster(mapa_u32(&tmem_empty_bar[buf], 0));   // the LEADER's barrier
          else mbar_arrive(&tmem_empty_bar[buf]);
        }
        const int c = n0 + ci * 32;
        const bool active = has_chunk && row_ok && c < p.N;   // per thread
        float f[32];
        if (active) {
          {
#pragma unroll
            for (int q = 0; q < 32; ++q) f[q] = __uint_as_float(v[q]) * pre;
            if (bias) {
              if (c + 32 <= p.N) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float4 bq = __ldg(reinterpret_cast<const float4*>(bias + c) + q);
                  f[4 * q] += bq.x; f[4 * q + 1] += bq.y; f[4 * q + 2] += bq.z; f[4 * q + 3] += bq.w;
                }
              } else {
#pragma unroll
                for (int q = 0; q < 32; ++q)
                  if (c + q < p.N) f[q] += bias[c + q];
              }
            }
            if constexpr (GEN) epi_dispatch(f, rs, e_act, e_gated);
            else epi_math<(EPI >> 2), ((EPI >> 1) & 1) != 0>(f, rs);
            if (p.argmax_out) {
#pragma unroll
              for (int q = 0; q < 32; ++q)
                if (c + q < p.argmax_n) {
                  const unsigned long long key = argmax_key(f[q], c + q);
                  best_key = key > best_key ? key : best_key;
                }
            }
          }
        }
        if (stamp) TC_STAMP(j, 11 + 6 * k);
        if (box_in_flight) {   // warp-uniform: the math above ran while the TMA engine read the previous box
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
        if (stamp) TC_STAMP(j, 12 + 6 * k);
        if (active) {
          {
            const int n_out = e_gated ? min(16, (p.N - c) >> 1) : min(32, p.N - c);
            const int64_t off = (int64_t)r * p.ldc + (e_gated ? (c >> 1) : c);
            const int64_t roff = (int64_t)r * p.ldr + (e_gated ? (c >> 1) : c);
            // residual may alias C (in-place x += ...): each element is read then written by this thread only
            const void* Rb = p.residual ? (const void*)(static_cast<const char*>(p.residual) + (roff - off) * (e_f32 ? 4 : 2)) : nullptr;
            if (e_f32) epi_store<true>(p.C, Rb, off, f, n_out);
            else if (!warp_tma)
              epi_store<false>(p.C, Rb, off, f, n_out, p.ss_out ? p.ss_out + (int64_t)r * p.ss_out_chunks + (c >> 5) : nullptr);
            else {
              const __nv_bfloat16* R = Rb ? static_cast<const __nv_bfloat16*>(Rb) + off : nullptr;
              const int b0 = k * (e_gated ? 32 : 64);  // byte offset of this chunk in the staged row (all boxes)
              uint8_t* dst = stg + stg_row;                          // this lane's row in the box
              const int u0 = (b0 % stg_rb) >> 4;                     // first 16-byte unit inside the box row
              const int sx = stg_xor;
              const float sq = epi_pack_bf16(R, f, n_out, p.ss_out != nullptr, [&](int q, const uint4& pk) {
                *reinterpret_cast<uint4*>(dst + ((((u0 + q) << 4)) ^ sx)) = pk;
              });
              if (p.ss_out) p.ss_out[(int64_t)r * p.ss_out_chunks + (c >> 5)] = sq;
            }
          }
        }
        if (stamp) TC_STAMP(j, 13 + 6 * k);
      }
      if (stage_any) {
        // generic-proxy smem writes -> visible to the async proxy, then one lane hands the box to the TMA engine
        // (columns >= N and rows >= M of the box are clipped by the tensor map)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          const int co = e_gated ? (c_first >> 1) : c_first;       // first output column of this warp
          const int n_out_cols = e_gated ? (p.N >> 1) : p.N;
          const int b = (stg_rb_all - 1) / stg_rb;   // the warp's last box (earlier ones left inside the chunk loop)
          if (co + b * (stg_rb >> 1) < n_out_cols) tma_store_2d(&mapC, stg, co + b * (stg_rb >> 1), m0 + quad * 32);
          tma_store_commit();
        }
      }
      if (best_key) atomicMax(p.argmax_out + r, best_key);
      if (stamp) TC_STAMP(j, 5);
    }
    // (waiting only for the shared-memory sources to be read - `.read` - measured neutral: the grid is not complete
    //  before the writes are, profiles/r02_experiments_not_kept.md)
    if (e_tma && lane == 0) tma_store_wait_all();
    if (warp == 2 && lane == 0) TC_STAMP(0, 6);   // smem sources read and writes complete before exit
  }

  // ---- teardown: everyone done with TMEM, then the allocating warp frees it ----
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into it / arrive on its barriers
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if constexpr (CL == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
}


// =====================================================================================================================
// FUSED MoE EXPERT MLP (Perceiver-TF feed-forward, model/ff_layer.py MoE experts = HF modeling_mixtral.py:62-85):
//   ys[slot, :] = slot_w[slot] * ( act(xs W1_e^T) * (xs W3_e^T) ) W2_e^T        for the rows of every expert e
// in ONE kernel: the hidden tile never leaves the SM.  The two grouped GEMMs it replaces write and re-read the
// (2 N, I) hidden matrix (4.3 GB + 4.3 GB per layer at 728 segments: GEMM1 is bound by that write).
// Shapes are compile-time: d_model 128, hidden 512 (the model's), rows of W13 interleaved (2j = activated, 2j+1 = linear).
// One CTA per SM walks the 128-row tiles of the expert-sorted rows (ragged last tile per expert: rows of the next
// expert are computed and not stored).  Per tile, with X = the tile's rows (2 k-blocks, resident), H = k-blocks of the
// hidden tile (128 x 64 bf16 in the tensor core's K-major 128-byte-swizzle layout, HALF a tile per g job):
//   MMA job list   c0 c1 c2 c3 c4 g0 c5 c6 c7 g1
//     c_i : acc[i & 1] (128 TMEM columns) = X  W13[e, 128 i .. 128 i + 127, :]^T      (2 k-blocks, N = 128)
//     g_h : Y (128 TMEM columns)        += H_h W2[e, :, 256 h .. 256 h + 255]^T      (4 k-blocks, N = 128)
//   epilogue warps, chunk i: acc[i & 1] -> act(even) * odd -> 64 bf16 outputs per row -> H slot (i & 3); after chunk 3 / 7
//   the half is published to the MMA warp; after g1: Y * slot_w -> bf16 -> global.
// Every weight block (W13 128 x 64, W2 128 x 64: 16 KB) streams through one 8-stage TMA ring in job order.
// Same arithmetic as the two grouped GEMMs (same k order of the accumulations, H rounded to bf16 exactly as the hidden
// matrix was): bit-identical outputs (tests/test_moe_gpu.py).
// =====================================================================================================================
namespace moefused {
constexpr int D = 128, I = 512;
constexpr int STAGE = 16384;                 // one 128 x 64 bf16 block
constexpr int HSLOTS = 5;                    // hidden k-blocks in shared memory: 4 being read by a g job + 1 being written
constexpr int X_OFF = 0, H_OFF = 2 * STAGE, RING_OFF = (2 + HSLOTS) * STAGE, NST = 7;
constexpr int BAR_OFF = RING_OFF + NST * STAGE;
constexpr int TOTAL = BAR_OFF + 1024 + 1024;   // barriers / tables, alignment slack
static_assert(TOTAL <= 227 * 1024, "shared memory budget");
constexpr int MAX_E = 32;
}  // namespace moefused

template <int ACT>
__global__ void __launch_bounds__(THREADS, 1)
moe_expert_fused_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW13,
                        const __grid_constant__ CUtensorMap mapW2, const int* __restrict__ offsets, int E,
                        const float* __restrict__ slot_w, __nv_bfloat16* __restrict__ ys) {
  using namespace moefused;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* ring_full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* ring_empty = ring_full + NST;
  uint64_t* x_full = ring_empty + NST;
  uint64_t* x_empty = x_full + 1;
  uint64_t* acc_full = x_empty + 1;     // [2]
  uint64_t* acc_empty = acc_full + 2;   // [2]
  uint64_t* h_full = acc_empty + 2;     // half of the hidden tile written (all epilogue threads arrive)
  uint64_t* h_empty = h_full + 1;       // the MMAs that read it have retired
  uint64_t* y_full = h_empty + 1;
  uint64_t* y_empty = y_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_empty + 1);
  int* goff = reinterpret_cast<int*>(tmem_slot + 2);   // [E + 1] row offsets of the experts
  int* gstart = goff + MAX_E + 1;                      // [E + 1] tile-index prefix
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW13) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW2) : "memory");
    for (int i = 0; i < NST; ++i) {
      mbar_init(&ring_full[i], 1);
      mbar_init(&ring_empty[i], 1);
    }
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], EPI_THREADS);
    }
    mbar_init(h_full, EPI_THREADS);
    mbar_init(h_empty, 1);
    mbar_init(y_full, 1);
    mbar_init(y_empty, EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    int acc = 0;
    for (int e = 0; e < E; ++e) {
      const int o = offsets[e];
      goff[e] = o;
      gstart[e] = acc;
      acc += (offsets[e + 1] - o + BM - 1) / BM;
    }
    goff[E] = offsets[E];
    gstart[E] = acc;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int total_tiles = gstart[E];

  auto decode_tile = [&](int t, int& e, int& m0, int& row_end) {
    e = 0;
    while (e + 1 < E && t >= gstart[e + 1]) ++e;
    m0 = goff[e] + (t - gstart[e]) * BM;
    row_end = goff[e + 1];
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      auto push = [&](const CUtensorMap* map, int c0, int c1) {
        mbar_wait(&ring_empty[stage], phase ^ 1);
        mbar_expect_tx(&ring_full[stage], STAGE);
        tma_load_2d(map, &ring_full[stage], smem + RING_OFF + stage * STAGE, c0, c1);
        if (++stage == NST) {
          stage = 0;
          phase ^= 1;
        }
      };
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        int e, m0, row_end;
        decode_tile(t, e, m0, row_end);
        mbar_wait(x_empty, (uint32_t)(it & 1) ^ 1);
        mbar_expect_tx(x_full, 2 * STAGE);
        tma_load_2d(&mapX, x_full, smem + X_OFF, 0, m0);
        tma_load_2d(&mapX, x_full, smem + X_OFF + STAGE, BK, m0);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          // job order of the MMA warp: c0 c1 c2 c3 c4 g0 | c5 c6 c7 g1
          const int c_lo = h == 0 ? 0 : 5, c_hi = h == 0 ? 4 : 7;
          for (int c = c_lo; c <= c_hi; ++c)
            for (int kb = 0; kb < 2; ++kb) push(&mapW13, kb * BK, e * (2 * I) + c * 128);
          for (int kb = 0; kb < 4; ++kb) push(&mapW2, (h * 4 + kb) * BK, e * D);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t n_acc[2] = {0, 0};   // uses of each accumulator buffer so far
    uint32_t n_h = 0;             // halves of the hidden tile consumed so far
    int it = 0;
    const uint64_t xdesc0 = umma_desc_sw128(smem + X_OFF), xdesc1 = umma_desc_sw128(smem + X_OFF + STAGE);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      mbar_wait(x_full, (uint32_t)(it & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      auto chunk = [&](int c) {
        const int buf = c & 1;
        mbar_wait(&acc_empty[buf], (n_acc[buf] & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * 128);
        for (int kb = 0; kb < 2; ++kb) {
          mbar_wait(&ring_full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            const uint64_t adesc = kb ? xdesc1 : xdesc0;
            const uint64_t bdesc = umma_desc_sw128(smem + RING_OFF + stage * STAGE);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
            umma_commit(&ring_empty[stage]);
            if (kb == 1) {
              umma_commit(&acc_full[buf]);
              if (c == 7) umma_commit(x_empty);   // the tile's rows are no longer read: the next tile's may land
            }
          }
          __syncwarp();
          if (++stage == NST) {
            stage = 0;
            phase ^= 1;
          }
        }
        ++n_acc[buf];
      };
      auto second = [&](int h) {
        mbar_wait(h_full, n_h & 1);                    // this half of the hidden tile is in shared memory
        if (h == 0) mbar_wait(y_empty, (uint32_t)(it & 1) ^ 1);   // the previous tile's output has been read out of TMEM
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_y = tmem_base + 256u;
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(&ring_full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            const uint64_t adesc = umma_desc_sw128(smem + H_OFF + (int)((4 * n_h + kb) % HSLOTS) * STAGE);
            const uint64_t bdesc = umma_desc_sw128(smem + RING_OFF + stage * STAGE);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_y, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (h | kb | k) ? 1u : 0u);
            umma_commit(&ring_empty[stage]);
            if (kb == 3) {
              umma_commit(h_empty);
              if (h == 1) umma_commit(y_full);
            }
          }
          __syncwarp();
          if (++stage == NST) {
            stage = 0;
            phase ^= 1;
          }
        }
        ++n_h;
      };
      chunk(0); chunk(1); chunk(2); chunk(3); chunk(4);
      second(0);
      chunk(5); chunk(6); chunk(7);
      second(1);
    }
  } else {
    // ===================== epilogue (8 warps) =====================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;           // which 64 of a chunk's 128 accumulator columns
    const int row_in_tile = quad * 32 + lane;
    const uint32_t h_base = smem_u32(smem + H_OFF);
    uint32_t n_acc[2] = {0, 0};
    uint32_t n_chunk = 0;                        // chunks written so far (all tiles): chunk n lives in slot n % HSLOTS
    uint32_t g_seen = 0;                         // g jobs (halves consumed by the MMA warp) known to have retired
    int it = 0;
    // the output of tile i is read out of TMEM after the first two chunks of tile i + 1 (its MMAs retire meanwhile)
    bool y_pending = false;
    int y_r = 0, y_it = 0;
    bool y_ok = false;
    float y_sw = 0.f;
    auto y_epilogue = [&]() {
      // ---- output: Y * slot weight -> bf16 -> global (each thread owns 64 contiguous columns of its row)
      mbar_wait(y_full, (uint32_t)(y_it & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t y_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(256 + half * 64);
      __nv_bfloat16* yrow = ys + (int64_t)(y_ok ? y_r : 0) * D + half * 64;
      uint32_t va[16], vb[16];
      tmem_ld16_nowait(y_addr, va);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (g & 1) tmem_wait_ld(vb); else tmem_wait_ld(va);
        if (g < 3) {
          if (g & 1) tmem_ld16_nowait(y_addr + 16 * (g + 1), va); else tmem_ld16_nowait(y_addr + 16 * (g + 1), vb);
        } else {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(y_empty);
        }
        const uint32_t(&v)[16] = (g & 1) ? vb : va;
        uint32_t o[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(v[2 * q]) * y_sw, __uint_as_float(v[2 * q + 1]) * y_sw);
          o[q] = *reinterpret_cast<const uint32_t*>(&b2);
        }
        if (y_ok) {
          *reinterpret_cast<uint4*>(yrow + 16 * g) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(yrow + 16 * g + 8) = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
      y_pending = false;
    };
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      int e, m0, row_end;
      decode_tile(t, e, m0, row_end);
      const int r = m0 + row_in_tile;
      const bool row_ok = r < row_end;
      const float sw = row_ok ? slot_w[r] : 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c, ++n_chunk) {
        if (c == 2 && y_pending) y_epilogue();
        const int buf = c & 1;
        if (n_chunk >= (uint32_t)HSLOTS) {
          // the slot's previous contents (chunk n - HSLOTS) were read by g job (n - HSLOTS) / 4: it must have retired
          const uint32_t need = (n_chunk - HSLOTS) / 4;
          while (g_seen <= need) {
            mbar_wait(h_empty, g_seen & 1);
            ++g_seen;
          }
        }
        mbar_wait(&acc_full[buf], n_acc[buf] & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 128 + half * 64);
        const uint32_t h_blk = h_base + (n_chunk % HSLOTS) * (uint32_t)STAGE;
        uint32_t va[16], vb[16];
        tmem_ld16_nowait(t_addr, va);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (g & 1) tmem_wait_ld(vb); else tmem_wait_ld(va);
          if (g < 3) {
            if (g & 1) tmem_ld16_nowait(t_addr + 16 * (g + 1), va); else tmem_ld16_nowait(t_addr + 16 * (g + 1), vb);
          } else {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acc_empty[buf]);
          }
          uint32_t h4[4];
          lean_half_gated<ACT, false, false>((g & 1) ? vb : va, 1.0f, 1.0f, nullptr, h4);
          // 8 outputs = 16-byte unit (half * 4 + g) of this row of the 128 x 64 K-major operand block, placed where the
          // 128-byte swizzle of the tensor core's shared-memory descriptor expects it (tma_box_offset, index_maps.h:
          // the layout a TMA load of the same block would produce; pinned on the CPU in tests/test_index_maps_emu.py)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(h_blk + (uint32_t)tma_box_offset(row_in_tile, half * 4 + g, 128)),
                       "r"(h4[0]), "r"(h4[1]), "r"(h4[2]), "r"(h4[3])
                       : "memory");
        }
        ++n_acc[buf];
        if ((c & 3) == 3) {
          // the half is complete: generic-proxy writes -> visible to the tensor core's reads, then publish
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(h_full);
        }
      }
      y_pending = true;
      y_r = r; y_ok = row_ok; y_sw = sw; y_it = it;
    }
    if (y_pending) y_epilogue();
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// =====================================================================================================================
// GEMM CHAIN: several dependent GEMMs of the decode step in ONE persistent launch.
//   phase 0:  C0 = epi0(A0 W0^T)      phase 1:  C1 = epi1(A1 W1^T)  with A1 = C0 (or an earlier output) ...
// Tile order is ROW-TILE MAJOR: the grid is cut into groups of `split` CTAs; group g owns the 128-row tiles g, g + G, ...
// and walks each of them through ALL phases, CTA l of the group taking the N tiles l, l + split, ... of every phase.
// A row tile's data therefore stays with the same `split` SMs from the first GEMM to the last, no CTA ever waits for a
// whole-grid wave to drain between two GEMMs, and the groups drift apart freely (the separate launches pay a launch,
// a pipeline fill and a wave tail per GEMM for 1-3 tiles of work per SM).  The TMA ring / double-buffered TMEM
// accumulator / epilogue warps, the barriers, the TMEM allocation and the tensor maps are set up ONCE for the chain.
// Dependencies are per row tile: a tile (phase p, row tile m) may load its A operand only after ALL N tiles of
// (phase p-1, m) have been stored.  A CTA publishes its share of (p, m) once, after its last N tile of that phase:
// stores complete (TMA bulk groups waited for WRITE completion) and fenced, then one red.release adds the number of
// tiles it produced to done[p][m]; the TMA producer of every CTA of the group polls done[p-1][m] with an acquire load
// before the first A load of (p, m) - the W halves of the first ring stages are already in flight by then (they do not
// depend on the previous phase).  Counters are monotonic: launch number `ordinal` (read from device memory = decode step
// x layers + layer, so the captured CUDA graph needs no patching) waits for (ordinal + 1) x n_tiles.  No deadlock: a
// CTA only waits for tiles of an earlier phase of the same row tile, all of which belong to CTAs of its own group,
// which are resident (grid <= number of SMs, one CTA per SM) and walk the same list in the same order.
// Memory: C leaves through TMA stores (async proxy) or direct stores; the producer's acquire + fence.proxy.async orders
// them before the next phase's TMA loads; residual / sum-of-squares reads bypass L1 (a line cached earlier in
// this launch may be stale).  Results are bit-identical to the separate launches (same tiles, same MMA order, same
// epilogue arithmetic) - tests/test_t5_gpu.py::test_gemm_chain_is_bit_identical.
// =====================================================================================================================
constexpr int CHAIN_MAX_PHASES = 4;
struct alignas(64) ChainPhase {
  CUtensorMap mapA, mapW, mapC;
  TcParams p;
  int tile0, n_tiles, num_kb, dep;   // first tile index, N tiles, k-blocks, 1 = depends on the previous phase
  int kind;                          // lean epilogue instance (EPI_LEAN_BF16 + flags / EPI_LEAN_GATED)
};
struct alignas(64) ChainArgs {
  ChainPhase ph[CHAIN_MAX_PHASES];
  int n_phases, m_tiles, total_tiles, split;   // split = CTAs per row tile
  int* done;               // [CHAIN_MAX_PHASES][m_tiles] completion counters, monotonic across launches
  const int* epoch_ptr;    // device int (decode step counter) or null
  int epoch_mul, epoch_add;
  unsigned long long* trace;   // debug (ymt3_debug_chain_trace): [cta][16 tiles][32 events] globaltimer ns, or null
};


__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

__global__ void __launch_bounds__(THREADS, 1) gemm_chain_kernel(const __grid_constant__ ChainArgs a) {
  constexpr int BN = 256;
  extern __shared__ uint8_t smem_raw[];
  using L = SmemLayout<BN, 1>;
  constexpr int STAGES = L::STAGES;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < a.n_phases; ++i) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.ph[i].mapA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.ph[i].mapW) : "memory");
      if (a.ph[i].p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&a.ph[i].mapC) : "memory");
    }
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], EPI_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) chain_stamp(a.trace, 0, 7);

  // this CTA's tile list: row tiles grp, grp + groups, ...; all phases of a row tile; N tiles ln, ln + split, ...
  // body(j, phase, row tile, N tile, first, last): j = running tile number of this CTA, first / last = first / last N
  // tile this CTA computes of (phase, row tile)
  const int split = a.split, groups = (int)gridDim.x / split;
  const int grp = (int)blockIdx.x / split, ln = (int)blockIdx.x % split;
  auto walk = [&](auto&& body) {
    int j = 0;
    for (int mt = grp; mt < a.m_tiles; mt += groups)
      for (int ph = 0; ph < a.n_phases; ++ph) {
        const int nn = a.ph[ph].n_tiles;
        for (int nt = ln; nt < nn; nt += split, ++j) body(j, ph, mt, nt, nt == ln, nt + split >= nn);
      }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const int ordinal = (a.epoch_ptr ? *a.epoch_ptr : 0) * a.epoch_mul + a.epoch_add;
      int stage = 0;
      uint32_t phase = 0;
      walk([&](int j, int ph, int mt, int nt, bool first, bool) {
        const ChainPhase& P = a.ph[ph];
        int kb0 = 0;
        chain_stamp(a.trace, j, 0);
        if (P.dep && first) {
          // every N tile of the previous phase for this row tile has been stored (and fenced) by its CTA
          const int* flag = a.done + (ph - 1) * a.m_tiles + mt;
          const int want = (ordinal + 1) * a.ph[ph - 1].n_tiles;
          if (ld_acquire_gpu(flag) < want) {
            // not yet: put the W halves of the first ring stages in flight, wait, then add the A halves
            const int npre = P.num_kb < STAGES ? P.num_kb : STAGES;
            int st = stage;
            uint32_t phs = phase;
            for (int kb = 0; kb < npre; ++kb) {
              mbar_wait(&empty_bar[st], phs ^ 1);
              mbar_expect_tx(&full_bar[st], L::STAGE_BYTES);
              tma_load_2d(&P.mapW, &full_bar[st], smem + st * L::STAGE_BYTES + L::A_BYTES, kb * BK, nt * BN);
              if (++st == STAGES) {
                st = 0;
                phs ^= 1;
              }
            }
            while (ld_acquire_gpu(flag) < want) __nanosleep(20);
            asm volatile("fence.proxy.async;" ::: "memory");   // those writes -> visible to the TMA loads below
            for (int kb = 0; kb < npre; ++kb) {
              tma_load_2d(&P.mapA, &full_bar[stage], smem + stage * L::STAGE_BYTES, kb * BK, mt * BM);
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
            kb0 = npre;
          } else {
            asm volatile("fence.proxy.async;" ::: "memory");
          }
        }
        chain_stamp(a.trace, j, 1);
        for (int kb = kb0; kb < P.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          tma_load_2d(&P.mapA, &full_bar[stage], sa, kb * BK, mt * BM);
          tma_load_2d(&P.mapW, &full_bar[stage], sb, kb * BK, nt * BN);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      });
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    walk([&](int j, int ph, int, int, bool, bool) {
      const int num_kb = a.ph[ph].num_kb;
      const int buf = j & 1;
      mbar_wait(&tmem_empty_bar[buf], ((j >> 1) & 1) ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          if (kb == 0) chain_stamp(a.trace, j, 2);
          const uint8_t* sa = smem + stage * L::STAGE_BYTES;
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + L::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) {
            umma_commit(&tmem_full_bar[buf]);
            chain_stamp(a.trace, j, 3);
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    });
  } else {
    // ===================== epilogue (8 warps): lean_tile, the kind of the phase selected once per tile ================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    uint8_t* stg = smem + L::STG_OFF + (warp - 2) * L::STG_WARP_BYTES;
    walk([&](int j, int ph, int mt, int nt, bool, bool last_nt) {
      const TcParams& p = a.ph[ph].p;
      const int buf = j & 1;
      const int m0 = mt * BM;
      LeanCtx lc;
      lc.mapC = &a.ph[ph].mapC;
      lc.N = p.N; lc.residual = p.residual; lc.ldr = p.ldr; lc.ss_out = p.ss_out; lc.ss_out_chunks = p.ss_out_chunks;
      lc.C = p.C; lc.ldc = p.ldc; lc.bias = p.bias;
      lc.m0 = m0; lc.row_end = p.M; lc.n0 = nt * BN; lc.r = m0 + quad * 32 + lane; lc.row_ok = lc.r < p.M;
      lc.warp_tma = m0 + quad * 32 < p.M;
      lc.rs = p.out_scale;
      lc.tmem_base = tmem_base; lc.buf = buf; lc.full_bar = &tmem_full_bar[buf]; lc.empty_bar = &tmem_empty_bar[buf];
      lc.full_parity = (uint32_t)((j >> 1) & 1);
      lc.stg = stg; lc.quad = quad; lc.half = half; lc.lane = lane; lc.j = j;
      lc.stamp = warp == 2 && lane == 0; lc.trace = a.trace;
      // the accumulator barrier also orders this thread after the producer's acquire of the previous phase: only now may
      // the sum-of-squares partials (written earlier in this launch, L1 bypassed) be read
      mbar_wait(lc.full_bar, lc.full_parity);
      float pre = 1.0f;
      if (p.norm_ss_in && lc.row_ok) {
        const float* sp = p.norm_ss_in + (int64_t)lc.r * p.norm_ss_chunks;
        float tot = 0.f;
        for (int c4 = 0; c4 + 4 <= p.norm_ss_chunks; c4 += 4) {
          const float4 s4 = __ldcg(reinterpret_cast<const float4*>(sp + c4));
          tot += (s4.x + s4.y) + (s4.z + s4.w);
        }
        for (int c1 = p.norm_ss_chunks & ~3; c1 < p.norm_ss_chunks; ++c1) tot += __ldcg(sp + c1);
        pre = rsqrtf(tot / (float)p.K + p.norm_eps);
      }
      lc.pre = pre;
      const int kind = a.ph[ph].kind;   // warp-uniform, once per tile
      lc.flags = kind & 7;
      if (kind < 80) lean_tile<BN, 72, 1, true>(lc);   // plain: bias / residual / sum-of-squares flags in registers
      else lean_tile<BN, 80, 1, true>(lc);             // gated gelu_new
      if (last_nt) {
        // publish this CTA's share of (phase, row tile): the bytes of all its tiles are WRITTEN (not only read out of
        // shared memory), every epilogue thread's stores (C, sum-of-squares) are fenced, then one thread adds the count
        if (lane == 0) tma_store_wait_all();
        __threadfence();
        epi_bar_sync();
        if (warp == 2 && lane == 0) {
          asm volatile("fence.proxy.async;" ::: "memory");
          red_release_gpu_add(a.done + ph * a.m_tiles + mt, (a.ph[ph].n_tiles - ln + split - 1) / split);
          chain_stamp(a.trace, j, 6);
        }
      }
    });
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 row-major (rows, cols) tensor with leading dimension ld (elements); box {64 cols, box_rows}
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  YMT3_REQUIRE(enc, "gemm_bf16_tc: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YMT3_REQUIRE(r == CUDA_SUCCESS, "gemm_bf16_tc: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
               (long long)rows, (long long)cols, (long long)ld);
  return YMT3_OK;
}

// bf16 output (rows, cols) with leading dimension ld: store box {row_bytes / 2 columns, 32 rows}, swizzle span = the
// box row (32 / 64 / 128 bytes)
int make_out_map(CUtensorMap* map, void* base, int64_t rows, int64_t cols, int64_t ld, int row_bytes) {
  EncodeTiledFn enc = get_encode_fn();
  YMT3_REQUIRE(enc, "gemm_bf16_tc: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)(row_bytes / 2), 32};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YMT3_REQUIRE(r == CUDA_SUCCESS, "gemm_bf16_tc: output cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
               (long long)rows, (long long)cols, (long long)ld);
  return YMT3_OK;
}

// NHWC activation (B, T, F, C) bf16 contiguous -> 4-D map {C, F, T, B}, box {64, 128, 1, 1}
int make_conv_map(CUtensorMap* map, const void* base, int64_t B, int64_t T, int64_t F, int64_t C) {
  EncodeTiledFn enc = get_encode_fn();
  YMT3_REQUIRE(enc, "conv3x3_bf16_tc: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)F, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)F * C * 2, (cuuint64_t)T * F * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)BM, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YMT3_REQUIRE(r == CUDA_SUCCESS, "conv3x3_bf16_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return YMT3_OK;
}

struct ConvGeom {
  int B = 0, T = 0, F = 0, Cin = 0;
};

// epilogue combinations with a specialised kernel (non-convolution): plain bf16 / plain fp32 (LM head) / gated GELU
// (decoder FFN) / gated SiLU (MoE experts); code = act * 4 + gated * 2 + out_f32
constexpr int EPI_PLAIN_BF16 = 0, EPI_PLAIN_F32 = 1, EPI_GATED_GELU_NEW = YMT3_ACT_GELU_NEW * 4 + 2,
              EPI_GATED_SILU = YMT3_ACT_SILU * 4 + 2;
// plain bf16 with every chunk on the lean path (N % 32 == 0, no groups, no arg-max, TMA store, one scale factor) - see the
// kernel; + 1: bias, + 2: residual, + 4: sum-of-squares output
constexpr int EPI_LEAN_BF16 = 64;
// gated (gelu_new / SiLU) bf16 output on the lean path (N % 32 == 0, tile width >= 64, groups allowed): + 1: bias, + 2: SiLU
constexpr int EPI_LEAN_GATED = 80;

template <int BN, bool CONV, int EPI, int CL>
int launch(const GemmParams& p, int out_dtype, cudaStream_t stream, const ConvGeom& cg = ConvGeom()) {
  CUtensorMap mapA, mapW, mapC;
  memset(&mapC, 0, sizeof(mapC));
  int rc;
  const int groups = p.group_offsets ? p.num_groups : 1;
  YMT3_REQUIRE(CL == 1 || !p.group_offsets, "gemm_bf16_tc: grouped GEMMs do not run in clusters");
  if constexpr (CONV) {
    if ((rc = make_conv_map(&mapA, p.A, cg.B, cg.T, cg.F, cg.Cin))) return rc;
  } else {
    if ((rc = make_map(&mapA, p.A, p.M, p.K, p.lda, BM))) return rc;
  }
  // grouped: weights of all groups are stacked along rows ((groups*N, K), strideW == N*ldw)
  // cluster: every CTA loads (and multicasts) a BN / CL-row slice of the W tile
  if ((rc = make_map(&mapW, p.W, (int64_t)p.N * groups, p.K, p.ldw, BN / CL))) return rc;
  // bf16 outputs leave through TMA stores (full-line writes issued by one lane per 32 x 32 chunk instead of 32
  // scattered 16-byte st.global per warp instruction); YMT3_GEMM_DIRECT_STORE=1 keeps the direct stores (A/B aid)
  static const bool direct_store = getenv("YMT3_GEMM_DIRECT_STORE") != nullptr;   // (forces the generic kernel)
  // (gated epilogues keep the direct stores: measured 383 -> 401 us on the MoE expert GEMM with 64-byte box rows,
  //  its limiter is the epilogue math, not the stores - profiles/r01_ab_gemm_tma_store.txt)
  static const bool tma_gated = getenv("YMT3_GEMM_TMA_GATED") != nullptr;
  const int tma_store = EPI >= EPI_LEAN_BF16 || (out_dtype == YMT3_BF16 && !direct_store && (!p.gated || tma_gated));
  const int out_row_bytes = SmemLayout<BN, CL>::CPW * (p.gated ? 32 : 64);
  if (tma_store && (rc = make_out_map(&mapC, p.C, p.M, p.gated ? p.N / 2 : p.N, p.ldc,
                                      out_row_bytes > 128 ? 128 : out_row_bytes)))
    return rc;
  TcParams t;
  t.tma_store = tma_store;
  t.C = p.C; t.ldc = p.ldc; t.bias = p.bias; t.residual = p.residual; t.ldr = p.ldr;
  t.M = p.M; t.N = p.N; t.K = p.K; t.act = p.act; t.gated = p.gated; t.out_scale = p.out_scale;
  t.norm_ss_in = p.norm_ss_in; t.norm_ss_chunks = p.norm_ss_chunks; t.norm_eps = p.norm_eps;
  t.ss_out = p.ss_out; t.ss_out_chunks = (p.N + 31) / 32;
#ifdef YMT3_GEMM_TRACE
  t.trace = g_chain_trace;
#endif
  t.argmax_out = p.argmax_out; t.argmax_n = p.argmax_n;
  t.row_scale = p.row_scale; t.group_offsets = p.group_offsets; t.num_groups = groups; t.out_f32 = out_dtype == YMT3_F32;
  t.conv_T = cg.T; t.conv_F = cg.F; t.conv_cblocks = CONV ? cg.Cin / BK : 0;
  static bool attr_set[64] = {false};   // per device (the attribute is per device and function)
  static int max_clusters[64] = {0};    // co-resident clusters of this instantiation (one CTA per SM)
  int dev = 0;
  YMT3_CUDA_CHECK(cudaGetDevice(&dev));
  const int sms = ymt3_num_sms();
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    YMT3_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tc_kernel<BN, CONV, EPI, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SmemLayout<BN, CL>::TOTAL));
    int mc = sms / CL;
    if constexpr (CL > 1) {
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3((unsigned)(sms / CL * CL));
      q.blockDim = dim3(THREADS);
      q.dynamicSmemBytes = SmemLayout<BN, CL>::TOTAL;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, gemm_bf16_tc_kernel<BN, CONV, EPI, CL>, &q) == cudaSuccess && n > 0) mc = n < mc ? n : mc;
      else (void)cudaGetLastError();
    }
    if (dev >= 0 && dev < 64) { attr_set[dev] = true; max_clusters[dev] = mc; }
  }
  const int mclusters = (dev >= 0 && dev < 64 && max_clusters[dev] > 0) ? max_clusters[dev] : sms / CL;
  // persistent: one CTA per SM (or per tile when there are fewer tiles than SMs); in grouped mode the tile
  // count depends on device-side offsets, so all SMs are launched and idle CTAs exit after setup
  const int64_t tiles = (int64_t)ymt3_div_up(p.M, BM * CL) * ymt3_div_up(p.N, BN);   // (super-)tiles
  const int grid = p.group_offsets ? sms : (int)(tiles < mclusters ? tiles : mclusters) * CL;
  YMT3_REQUIRE(groups <= 32, "gemm_bf16_tc: at most 32 groups");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SmemLayout<BN, CL>::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (ymt3_pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if constexpr (CL > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CL; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  YMT3_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_bf16_tc_kernel<BN, CONV, EPI, CL>, mapA, mapW, mapC, t));
  return YMT3_OK;
}

// CTA pair (cta_group::2) or single CTA?  MEASURED on B200 (profiles/r02_ab_gemm_pair.txt, r02_ab_gemm_cluster.txt):
//   8192^3: 1152 TFLOP/s single -> 1376 TFLOP/s as a pair with a 5-deep ring (85 % of the measured cuBLAS burst, 99.5 % of
//   its sustained figure); the same cluster with cta_group::1 MMAs and the W tile multicast: 1159 (neutral);
//   every GEMM / convolution OF THIS MODEL (K <= 1536, one to four tiles per SM): the pair is 4-25 % SLOWER (decode-step
//   GEMMs at M = 9464: 26.5 -> 30.1 us, conv pre-encoder 22.6 -> 24.0 ms, K = 128 projections 66 -> 82 us): two CTAs
//   that must stay in lock step pay for every epilogue twice as long as it takes the slower one.
// So the pair is chosen for long-K, many-tile GEMMs only; YMT3_GEMM_CLUSTER=1 / 2 forces either (A/B aid).
int cluster_size(const GemmParams& p, int64_t m_tiles) {
  static const int want = getenv("YMT3_GEMM_CLUSTER") ? atoi(getenv("YMT3_GEMM_CLUSTER")) : 0;
  if (p.group_offsets || m_tiles < 2 || want == 1) return 1;
  if (want == 2) return 2;
  return (p.K >= 2048 && m_tiles >= 16 && p.N >= 512) ? 2 : 1;
}

}  // namespace

// Host side of the GEMM chain (see gemm_chain_kernel).  phases[i]: an ordinary GemmParams (bf16 in / out, no groups,
// no arg-max, no row scale); phase i > 0 depends on phase i - 1 per M tile.  done: device ints, CHAIN_MAX_PHASES x
// ceil(M / 128), zeroed by the caller before the first launch of a run; launch ordinal = *epoch_ptr * epoch_mul +
// epoch_add must count the launches that used `done` since then (0, 1, 2, ...).
void gemm_chain_set_trace(unsigned long long* buf) { g_chain_trace = buf; }

int gemm_chain_bf16(const GemmParams* phases, int n_phases, int* done, const int* epoch_ptr, int epoch_mul, int epoch_add,
                    cudaStream_t stream) {
  constexpr int BN = 256;
  YMT3_REQUIRE(phases && n_phases >= 1 && n_phases <= CHAIN_MAX_PHASES && done, "gemm_chain: bad arguments");
  const int M = phases[0].M;
  if (M <= 0) return YMT3_OK;
  ChainArgs a;
  memset(&a, 0, sizeof(a));
  a.n_phases = n_phases;
  a.m_tiles = ymt3_div_up(M, BM);
  a.done = done;
  a.epoch_ptr = epoch_ptr; a.epoch_mul = epoch_mul; a.epoch_add = epoch_add;
  a.trace = g_chain_trace;
  int tile0 = 0, rc;
  for (int i = 0; i < n_phases; ++i) {
    const GemmParams& p = phases[i];
    YMT3_REQUIRE(p.A && p.W && p.C && p.M == M && p.N > 0 && p.K > 0, "gemm_chain: phase %d: bad shape", i);
    YMT3_REQUIRE(p.K % 8 == 0 && p.lda % 8 == 0 && p.ldw % 8 == 0 && p.ldc % 8 == 0 && (!p.residual || p.ldr % 8 == 0) &&
                     p.N % (p.gated ? 16 : 8) == 0,
                 "gemm_chain: phase %d: alignment", i);
    YMT3_REQUIRE(!p.group_offsets && !p.argmax_out && !p.row_scale, "gemm_chain: phase %d: unsupported epilogue option", i);
    // every phase runs a lean epilogue instance (lean_tile): full 32-column chunks, one scale factor
    YMT3_REQUIRE(p.N % 32 == 0 && p.out_scale == 1.0f, "gemm_chain: phase %d: needs N %% 32 == 0 and out_scale 1", i);
    const int flags = (p.bias ? 1 : 0) | (p.residual ? 2 : 0) | (p.ss_out ? 4 : 0);
    if (p.gated)
      YMT3_REQUIRE(p.act == YMT3_ACT_GELU_NEW && flags == 0, "gemm_chain: phase %d: gated phases are gelu_new without bias / residual", i);
    else
      YMT3_REQUIRE(p.act == YMT3_ACT_NONE && (flags == 0 || flags == 1 || flags == 6 || flags == 7),
                   "gemm_chain: phase %d: unsupported bias / residual / sum-of-squares combination %d", i, flags);
    YMT3_REQUIRE(!p.ss_out || (!p.gated && p.N % 32 == 0), "gemm_chain: phase %d: sum-of-squares output needs N %% 32 == 0", i);
    YMT3_REQUIRE(!p.norm_ss_in || (p.norm_ss_chunks > 0 && p.norm_ss_chunks % 4 == 0), "gemm_chain: phase %d: norm partials", i);
    ChainPhase& P = a.ph[i];
    if ((rc = make_map(&P.mapA, p.A, p.M, p.K, p.lda, BM))) return rc;
    if ((rc = make_map(&P.mapW, p.W, p.N, p.K, p.ldw, BN))) return rc;
    const int tma_store = 1;
    if ((rc = make_out_map(&P.mapC, p.C, p.M, p.gated ? p.N / 2 : p.N, p.ldc, 128))) return rc;
    P.kind = p.gated ? 80 : 64 + flags;
    TcParams& t = P.p;
    t.tma_store = tma_store;
    t.C = p.C; t.ldc = p.ldc; t.bias = p.bias; t.residual = p.residual; t.ldr = p.ldr;
    t.M = p.M; t.N = p.N; t.K = p.K; t.act = p.act; t.gated = p.gated; t.out_scale = p.out_scale;
    t.norm_ss_in = p.norm_ss_in; t.norm_ss_chunks = p.norm_ss_chunks; t.norm_eps = p.norm_eps;
    t.ss_out = p.ss_out; t.ss_out_chunks = (p.N + 31) / 32;
    t.num_groups = 1;
    P.tile0 = tile0;
    P.n_tiles = ymt3_div_up(p.N, BN);
    P.num_kb = ymt3_div_up(p.K, BK);
    P.dep = i > 0;
    tile0 += a.m_tiles * P.n_tiles;
  }
  a.total_tiles = tile0;
  static bool attr_set[64] = {false};
  int dev = 0;
  YMT3_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    YMT3_CUDA_CHECK(cudaFuncSetAttribute(gemm_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<BN, 1>::TOTAL));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int sms = ymt3_num_sms();
  // one CTA per SM, all co-resident (the dependency waits rely on it): never more CTAs than SMs.  `split` CTAs share a
  // row tile: as many as the SM count allows, at most the widest phase's N tiles.
  int max_nt = 1;
  for (int i = 0; i < n_phases; ++i) max_nt = a.ph[i].n_tiles > max_nt ? a.ph[i].n_tiles : max_nt;
  int split = sms / a.m_tiles;
  if (const char* e = getenv("YMT3_GEMM_CHAIN_SPLIT")) split = atoi(e);
  split = split < 1 ? 1 : (split > max_nt ? max_nt : split);
  if (split > sms) split = sms;
  a.split = split;
  const int groups = a.m_tiles < sms / split ? a.m_tiles : sms / split;
  const int grid = groups * split;
  YMT3_CUDA_CHECK(ymt3_launch_pdl(gemm_chain_kernel, dim3(grid), dim3(THREADS), SmemLayout<BN, 1>::TOTAL, stream, a));
  return YMT3_OK;
}

int gemm_chain_counters(int M) { return CHAIN_MAX_PHASES * ymt3_div_up(M, BM); }

// Fused MoE expert MLP (moe_expert_fused_kernel): xs (S, 128) bf16 expert-sorted, w13 (E, 1024, 128), w2 (E, 128, 512),
// offsets (E + 1) device ints, slot_w (S) fp32, ys (S, 128) bf16.  Returns YMT3_ERR_* ; shapes other than d_model 128 /
// hidden 512 / SiLU or gelu_new are the caller's business (two grouped GEMMs).
int moe_expert_fused(const void* xs, int64_t S, const void* w13, const void* w2, const int* offsets, int E,
                     const float* slot_w, void* ys, int act, cudaStream_t stream) {
  using namespace moefused;
  YMT3_REQUIRE(xs && w13 && w2 && offsets && slot_w && ys, "moe_expert_fused: null pointer");
  YMT3_REQUIRE(E >= 1 && E <= MAX_E && S >= 0 && S < (1ll << 31), "moe_expert_fused: bad shape");
  YMT3_REQUIRE(act == YMT3_ACT_SILU || act == YMT3_ACT_GELU_NEW, "moe_expert_fused: activation %d", act);
  if (S == 0) return YMT3_OK;
  CUtensorMap mapX, mapW13, mapW2;
  int rc;
  if ((rc = make_map(&mapX, xs, S, D, D, BM))) return rc;
  if ((rc = make_map(&mapW13, w13, (int64_t)E * 2 * I, D, D, 128))) return rc;
  if ((rc = make_map(&mapW2, w2, (int64_t)E * D, I, I, 128))) return rc;
  static bool attr_set[64][2] = {};
  int dev = 0;
  YMT3_CUDA_CHECK(cudaGetDevice(&dev));
  const int ai = act == YMT3_ACT_SILU ? 0 : 1;
  if (dev < 0 || dev >= 64 || !attr_set[dev][ai]) {
    if (ai == 0)
      YMT3_CUDA_CHECK(cudaFuncSetAttribute(moe_expert_fused_kernel<YMT3_ACT_SILU>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOTAL));
    else
      YMT3_CUDA_CHECK(cudaFuncSetAttribute(moe_expert_fused_kernel<YMT3_ACT_GELU_NEW>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOTAL));
    if (dev >= 0 && dev < 64) attr_set[dev][ai] = true;
  }
  const int sms = ymt3_num_sms();
  const int64_t max_tiles = S / BM + E;   // upper bound of the tile count (one ragged tile per expert)
  const int grid = (int)(max_tiles < sms ? max_tiles : sms);
  if (ai == 0)
    moe_expert_fused_kernel<YMT3_ACT_SILU><<<grid, THREADS, TOTAL, stream>>>(mapX, mapW13, mapW2, offsets, E, slot_w, (__nv_bfloat16*)ys);
  else
    moe_expert_fused_kernel<YMT3_ACT_GELU_NEW><<<grid, THREADS, TOTAL, stream>>>(mapX, mapW13, mapW2, offsets, E, slot_w, (__nv_bfloat16*)ys);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

int gemm_bf16_tc(const GemmParams& p, int out_dtype, cudaStream_t stream) {
  YMT3_REQUIRE(p.A && p.W && p.C, "gemm_bf16_tc: null pointer");
  if (p.M <= 0 || p.N <= 0) return YMT3_OK;
  YMT3_REQUIRE(p.K > 0 && p.K % 8 == 0 && p.lda % 8 == 0 && p.ldw % 8 == 0,
               "gemm_bf16_tc: K, lda, ldw must be multiples of 8 (K=%d lda=%lld ldw=%lld)", p.K, (long long)p.lda,
               (long long)p.ldw);
  YMT3_REQUIRE(p.N % (p.gated ? 16 : 8) == 0, "gemm_bf16_tc: N must be a multiple of %d (N=%d)", p.gated ? 16 : 8, p.N);
  YMT3_REQUIRE(p.ldc % 8 == 0 && (!p.residual || p.ldr % 8 == 0), "gemm_bf16_tc: ldc/ldr must be multiples of 8");
  YMT3_REQUIRE((((uintptr_t)p.A | (uintptr_t)p.W | (uintptr_t)p.C | (uintptr_t)p.residual) & 15) == 0,
               "gemm_bf16_tc: pointers must be 16-byte aligned");
  YMT3_REQUIRE(!p.group_offsets || p.strideW == (int64_t)p.N * p.ldw,
               "gemm_bf16_tc: grouped weights must be stacked contiguously");
  YMT3_REQUIRE(!p.ss_out || (out_dtype == YMT3_BF16 && !p.gated && p.N % 32 == 0),
               "gemm_bf16_tc: sum-of-squares output needs bf16, non-gated, N %% 32 == 0");
  YMT3_REQUIRE(!p.norm_ss_in || (p.norm_ss_chunks > 0 && ((uintptr_t)p.norm_ss_in & 15) == 0 && p.norm_ss_chunks % 4 == 0),
               "gemm_bf16_tc: fused-norm partials must be 16-byte aligned, a multiple of 4 per row");
  YMT3_REQUIRE(!p.argmax_out || (out_dtype == YMT3_F32 && !p.gated && !p.residual && !p.group_offsets && p.argmax_n > 0 &&
                                 p.argmax_n <= p.N),
               "gemm_bf16_tc: fused arg-max needs fp32 output, no gate / residual / groups, 0 < argmax_n <= N");
  // pick BN so that the grid fills the 148 SMs (2 CTAs/SM resident) when the problem allows
  const int64_t mt = ymt3_div_up(p.M, BM);
  const int sms = ymt3_num_sms();
  // Tile width: the mainloop of these GEMMs is bound by the L2 -> shared-memory fill (a 128 x BN tile moves
  // (128 + BN) * K * 2 bytes for 128 * BN outputs), and every SM works through ceil(tiles / SMs) tiles, so pick the BN
  // in {256, 128, 64, 32} that minimises rounds * (128 + BN) (ties: the wider tile).  With fewer tiles than SMs that is
  // one round of the narrowest tile = parallelism first; at M = 6656 (52 M-tiles) it picks 256 for every GEMM of the
  // decode step (e.g. N = 512: 104 tiles in ONE round instead of 208 tiles of 128 in two).  YMT3_GEMM_MAX_BN (A/B aid)
  // caps BN.
  static const int max_bn = getenv("YMT3_GEMM_MAX_BN") ? atoi(getenv("YMT3_GEMM_MAX_BN")) : 256;
  const int bn = gemm_choose_bn(mt, p.N, sms, max_bn);   // index_maps.h
  static const bool generic_only = getenv("YMT3_GEMM_DIRECT_STORE") || getenv("YMT3_GEMM_TMA_GATED") ||
                                   getenv("YMT3_GEMM_GENERIC");   // A/B switches act on the run-time kernel
  int code = generic_only ? -1 : p.act * 4 + (p.gated ? 2 : 0) + (out_dtype == YMT3_F32 ? 1 : 0);
  static const bool no_lean = getenv("YMT3_GEMM_NO_LEAN") != nullptr;   // A/B aid
  const bool cl2 = cluster_size(p, mt) == 2;
  // lean epilogue instances (single CTA): every chunk full, ONE scale factor (the other exactly 1), compile-time bias /
  // residual / sum-of-squares flags - the combinations the model's GEMMs use
  if (code == EPI_PLAIN_BF16 && !no_lean && p.N % 32 == 0 && !p.group_offsets && !p.argmax_out && !p.row_scale &&
      (p.out_scale == 1.0f || (!p.norm_ss_in && !p.bias))) {
    const int flags = (p.bias ? 1 : 0) | (p.residual ? 2 : 0) | (p.ss_out ? 4 : 0);
    if (flags != 4 && flags != 5) code = EPI_LEAN_BF16 + flags;
  }
  if ((code == EPI_GATED_GELU_NEW || code == EPI_GATED_SILU) && !no_lean && bn >= 64 && p.N % 32 == 0 && !p.argmax_out &&
      !p.residual && !p.ss_out)
    code = EPI_LEAN_GATED + (p.bias ? 1 : 0) + (code == EPI_GATED_SILU ? 2 : 0);
#define YMT3_TC_LAUNCH(EPI)                                                                                      \
  switch (bn) {                                                                                                  \
    case 256: return cl2 ? launch<256, false, EPI, 2>(p, out_dtype, stream) : launch<256, false, EPI, 1>(p, out_dtype, stream); \
    case 128: return cl2 ? launch<128, false, EPI, 2>(p, out_dtype, stream) : launch<128, false, EPI, 1>(p, out_dtype, stream); \
    case 64: return cl2 ? launch<64, false, EPI, 2>(p, out_dtype, stream) : launch<64, false, EPI, 1>(p, out_dtype, stream);   \
    default: return cl2 ? launch<32, false, EPI, 2>(p, out_dtype, stream) : launch<32, false, EPI, 1>(p, out_dtype, stream);  \
  }
  switch (code) {
#define YMT3_TC_LAUNCH_LEAN(EPI)                                                  \
  switch (bn) {                                                                  \
    case 256: return cl2 ? launch<256, false, EPI, 2>(p, out_dtype, stream) : launch<256, false, EPI, 1>(p, out_dtype, stream); \
    case 128: return cl2 ? launch<128, false, EPI, 2>(p, out_dtype, stream) : launch<128, false, EPI, 1>(p, out_dtype, stream); \
    case 64: return launch<64, false, EPI, 1>(p, out_dtype, stream);             \
    default: return launch<32, false, EPI, 1>(p, out_dtype, stream);             \
  }
    case EPI_LEAN_BF16 + 0: YMT3_TC_LAUNCH_LEAN(EPI_LEAN_BF16 + 0)
    case EPI_LEAN_BF16 + 1: YMT3_TC_LAUNCH_LEAN(EPI_LEAN_BF16 + 1)
    case EPI_LEAN_BF16 + 2: YMT3_TC_LAUNCH_LEAN(EPI_LEAN_BF16 + 2)
    case EPI_LEAN_BF16 + 3: YMT3_TC_LAUNCH_LEAN(EPI_LEAN_BF16 + 3)
    case EPI_LEAN_BF16 + 6: YMT3_TC_LAUNCH_LEAN(EPI_LEAN_BF16 + 6)
    case EPI_LEAN_BF16 + 7: YMT3_TC_LAUNCH_LEAN(EPI_LEAN_BF16 + 7)
#define YMT3_TC_LAUNCH_LEAN_G(EPI)                                                \
  switch (bn) {                                                                  \
    case 256: return cl2 ? launch<256, false, EPI, 2>(p, out_dtype, stream) : launch<256, false, EPI, 1>(p, out_dtype, stream); \
    case 128: return cl2 ? launch<128, false, EPI, 2>(p, out_dtype, stream) : launch<128, false, EPI, 1>(p, out_dtype, stream); \
    default: return launch<64, false, EPI, 1>(p, out_dtype, stream);             \
  }
    case EPI_LEAN_GATED + 0: YMT3_TC_LAUNCH_LEAN_G(EPI_LEAN_GATED + 0)
    case EPI_LEAN_GATED + 1: YMT3_TC_LAUNCH_LEAN_G(EPI_LEAN_GATED + 1)
    case EPI_LEAN_GATED + 2: YMT3_TC_LAUNCH_LEAN_G(EPI_LEAN_GATED + 2)
    case EPI_LEAN_GATED + 3: YMT3_TC_LAUNCH_LEAN_G(EPI_LEAN_GATED + 3)
#undef YMT3_TC_LAUNCH_LEAN_G
#undef YMT3_TC_LAUNCH_LEAN
    case EPI_PLAIN_BF16: YMT3_TC_LAUNCH(EPI_PLAIN_BF16)
    case EPI_PLAIN_F32: YMT3_TC_LAUNCH(EPI_PLAIN_F32)
    case EPI_GATED_GELU_NEW: YMT3_TC_LAUNCH(EPI_GATED_GELU_NEW)
    case EPI_GATED_SILU: YMT3_TC_LAUNCH(EPI_GATED_SILU)
    default: YMT3_TC_LAUNCH(-1)
  }
#undef YMT3_TC_LAUNCH
}

// 3x3 / stride 1 / zero-pad 1 convolution as an implicit GEMM on the tensor cores.
//   x: (B, T, F, Cin) bf16 NHWC, F % 128 == 0, Cin % 64 == 0.  W: (Cout, 9*Cin) bf16 with k = (ky*3+kx)*Cin + c.
//   y: (B*T*F, Cout) with the GemmParams epilogue (bias / act / residual / out dtype).
int conv3x3_bf16_tc(const void* x, int B, int T, int F, int Cin, const GemmParams& epi, int out_dtype,
                    cudaStream_t stream) {
  YMT3_REQUIRE(x && epi.W && epi.C, "conv3x3_bf16_tc: null pointer");
  YMT3_REQUIRE(F % BM == 0 && Cin % BK == 0, "conv3x3_bf16_tc: need F %% 128 == 0 and Cin %% 64 == 0 (F=%d Cin=%d)", F, Cin);
  const int64_t M64 = (int64_t)B * T * F;
  YMT3_REQUIRE(M64 < (1ll << 31), "conv3x3_bf16_tc: too many pixels");
  GemmParams p = epi;
  p.A = x; p.lda = Cin;
  p.M = (int)M64; p.K = 9 * Cin; p.ldw = 9 * Cin;
  p.group_offsets = nullptr; p.gated = 0;
  YMT3_REQUIRE(p.N % 8 == 0 && p.ldc % 8 == 0 && (!p.residual || p.ldr % 8 == 0), "conv3x3_bf16_tc: Cout/ldc alignment");
  ConvGeom cg;
  cg.B = B; cg.T = T; cg.F = F; cg.Cin = Cin;
  const bool cl2 = cluster_size(p, M64 / BM) == 2;
  if (p.N % 128 == 0) return cl2 ? launch<128, true, -1, 2>(p, out_dtype, stream, cg) : launch<128, true, -1, 1>(p, out_dtype, stream, cg);
  if (p.N % 64 == 0) return cl2 ? launch<64, true, -1, 2>(p, out_dtype, stream, cg) : launch<64, true, -1, 1>(p, out_dtype, stream, cg);
  return cl2 ? launch<32, true, -1, 2>(p, out_dtype, stream, cg) : launch<32, true, -1, 1>(p, out_dtype, stream, cg);
}

}  // namespace ymt3
