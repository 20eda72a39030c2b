// Small pure functions shared by the kernels and the CPU emulation harness (tests/host_emu): index maps, packings and
// host-side launch heuristics whose correctness does not need a GPU to check.  Everything here is __host__ __device__
// (or host only) and free of CUDA runtime calls, so tests/host_emu can compile it for the CPU and the CPU test tier can
// pin it against independent restatements (tests/test_index_maps_emu.py).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define YMT3_HD __host__ __device__ __forceinline__
#else
#define YMT3_HD inline
#endif

// ---- fused greedy selection (GemmParams::argmax_out) --------------------------------------------------------------
// Monotone packing of (fp32 logit, column) into one 64-bit key for atomicMax: the larger value wins, among equal
// values the SMALLER column (= torch.argmax's first-maximum rule).  -0.0 is folded into +0.0 so that equal values
// produce equal high words.  column = 0xFFFFFFFF - (key & 0xFFFFFFFF).
YMT3_HD unsigned long long argmax_key(float v, int col) {
  const float c = v + 0.0f;
  unsigned int u;
#if defined(__CUDA_ARCH__)
  u = __float_as_uint(c);
#else
  memcpy(&u, &c, 4);
#endif
  u = (u >> 31) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)col);
}

// ---- TMA-store staging of the tcgen05 GEMM epilogue ---------------------------------------------------------------
// A staging box holds 32 rows of rb bytes (rb = 32, 64 or 128 = the swizzle span of the output tensor map:
// CU_TENSOR_MAP_SWIZZLE_32B / 64B / 128B = Swizzle<1|2|3, 4, 3>: address bits [4, 4+B) are XOR-ed with bits [7, 7+B)).
// With the box base aligned to 1024 bytes those bits come from the row index alone.
YMT3_HD int tma_swizzle_xor(int row, int rb) { return (((row * rb) >> 7) & ((rb >> 4) - 1)) << 4; }
// byte offset inside the box of 16-byte unit `unit` of row `row`
YMT3_HD int tma_box_offset(int row, int unit, int rb) { return row * rb + ((unit << 4) ^ tma_swizzle_xor(row, rb)); }

// ---- swizzled shared-memory tiles of the mma.sync attention kernels (attention.cu) ---------------------------------
// 16-byte chunk c of row r.  ldmatrix reads 8 rows x 16 bytes per 8-lane phase: the 8 addresses of a phase (8 consecutive
// rows, same logical chunk) must fall into 8 different 16-byte bank groups.
// dk 16 tiles (32-byte rows, 2 chunks): chunk ^= bit 2 of the row
YMT3_HD uint32_t tc_row_off(int r, int c) { return (uint32_t)(r * 32 + ((c ^ ((r >> 2) & 1)) << 4)); }
// dk 128 tiles (256-byte rows, 16 chunks): low 3 bits of the chunk ^= row & 7
YMT3_HD uint32_t xw_off(int r, int c) { return (uint32_t)(r * 256 + ((c ^ (r & 7)) << 4)); }

// ---- tile width of the tcgen05 GEMM (host) --------------------------------------------------------------------------
// The mainloop is bound by the L2 -> shared-memory fill: a 128 x BN tile moves (128 + BN) * K * 2 bytes for 128 * BN
// outputs, and every SM walks ceil(tiles / sms) tiles, so BN in {256, 128, 64, 32} is the argmin of
// rounds * (128 + BN); ties go to the wider tile; a BN wider than N is never used (except the minimum 32).
static inline int gemm_choose_bn(int64_t m_tiles, int N, int sms, int max_bn) {
  int bn = 32;
  int64_t best = -1;
  const int cand[4] = {256, 128, 64, 32};
  for (int i = 0; i < 4; ++i) {
    const int b = cand[i];
    if (b > max_bn || (b > 32 && N < b)) continue;
    const int64_t tiles = m_tiles * ((N + b - 1) / b);
    const int64_t cost = ((tiles + sms - 1) / sms) * (int64_t)(128 + b);
    if (best < 0 || cost < best) {
      best = cost;
      bn = b;
    }
  }
  return bn;
}
