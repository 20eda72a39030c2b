// CPU side of yourmt3_b200/csrc/index_maps.h: the packings, shared-memory index maps and launch heuristics of the
// GEMM epilogue / fused greedy selection compiled for the host, plus a thread-by-thread emulation of how the epilogue
// threads reduce a row of logits (one key per (row, 32-column chunk) thread, atomicMax across threads).  Test
// infrastructure only (tests/test_index_maps_emu.py); never used by the product path.
#include "../../yourmt3_b200/csrc/index_maps.h"

#define EMU_API extern "C" __attribute__((visibility("default")))

EMU_API unsigned long long emu_argmax_key(float v, int col) { return argmax_key(v, col); }

// logits (M, ld) fp32; keys[m] = max over the epilogue threads' keys, exactly as the GEMM epilogues build them:
// tile width bn, each thread owns one row and 32-column chunks, columns >= V never contribute
EMU_API void emu_argmax_rows(const float* logits, long long M, long long ld, int N, int V, int bn, unsigned long long* keys) {
  for (long long m = 0; m < M; ++m) {
    unsigned long long row_key = 0;   // the zeroed global slot
    for (int n0 = 0; n0 < N; n0 += bn)
      for (int c = n0; c < n0 + bn && c < N; c += 32) {
        unsigned long long best = 0;  // one epilogue thread
        for (int q = 0; q < 32; ++q)
          if (c + q < V && c + q < N) {
            const unsigned long long k = argmax_key(logits[m * ld + c + q], c + q);
            best = k > best ? k : best;
          }
        if (best && best > row_key) row_key = best;   // atomicMax
      }
    keys[m] = row_key;
  }
}

EMU_API int emu_tma_box_offset(int row, int unit, int rb) { return tma_box_offset(row, unit, rb); }
EMU_API int emu_gemm_choose_bn(long long m_tiles, int N, int sms, int max_bn) { return gemm_choose_bn(m_tiles, N, sms, max_bn); }
EMU_API unsigned emu_tc_row_off(int r, int c) { return tc_row_off(r, c); }
EMU_API unsigned emu_xw_off(int r, int c) { return xw_off(r, c); }
