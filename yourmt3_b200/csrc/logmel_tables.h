// Host-side constant tables of the fused log-mel kernel (twiddles + banded mel
// filterbank). Shared by the CUDA library and the CPU emulation harness in tests/.
#pragma once
#include "logmel_core.cuh"
#include "../../include/ymt3_b200.h"
#include <math.h>
#include <vector>

struct LmHostTables {
  std::vector<float2> tw1, tw2;
  std::vector<int> first, off;
  std::vector<int2> meta;   // (first | count << 16, offset)
  std::vector<float> wts;
  std::vector<float4> rec_w;  // fixed-width filter records (LmTables::rec_w), [LM_MEL_NV][n_mels]
  int rec_ok = 0;
  int n_out = 0;
};

// returns nullptr on success, else a static error string
static inline const char* lm_build_host_tables(const ymt3_audio_cfg_t* cfg, const float* fb_host,
                                               LmHostTables& t) {
  t.tw1.resize(16 * 128);
  t.tw2.resize(8 * 16);
  const double two_pi = 6.283185307179586476925286766559;
  for (int k1 = 0; k1 < 16; ++k1)
    for (int m = 0; m < 128; ++m) {
      double a = -two_pi * (double)(m * k1) / 2048.0;
      t.tw1[k1 * 128 + m] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int n3 = 0; n3 < 8; ++n3)
    for (int k2 = 0; k2 < 16; ++k2) {
      double a = -two_pi * (double)(n3 * k2) / 128.0;
      t.tw2[n3 * 16 + k2] = make_float2((float)cos(a), (float)sin(a));
    }
  if (cfg->codec == YMT3_CODEC_MELSPEC) {
    if (!fb_host || cfg->n_mels <= 0) return "melspec codec needs fb_host and n_mels > 0";
    const int M = cfg->n_mels;
    t.first.resize(M);
    t.off.resize(M + 1);
    t.off[0] = 0;
    for (int m = 0; m < M; ++m) {
      int lo = -1, hi = -1;
      for (int k = 0; k < LM_NBINS; ++k)
        if (fb_host[(size_t)k * M + m] != 0.f) {
          if (lo < 0) lo = k;
          hi = k;
        }
      if (lo < 0) {  // empty filter (torchaudio warns, output is log(eps))
        t.first[m] = 0;
        t.off[m + 1] = t.off[m];
        continue;
      }
      t.first[m] = lo;
      for (int k = lo; k <= hi; ++k) t.wts.push_back(fb_host[(size_t)k * M + m]);
      t.off[m + 1] = t.off[m] + (hi - lo + 1);
    }
    t.n_out = M;
  } else {
    if (cfg->spec_bin0 < 0 || cfg->spec_bins <= 0 || cfg->spec_bin0 + cfg->spec_bins > LM_NBINS)
      return "spec bins outside [0, 1025)";
    t.n_out = cfg->spec_bins;
    t.first.assign(1, 0);
    t.off.assign(2, 0);
  }
  if (t.wts.empty()) t.wts.push_back(0.f);
  t.meta.resize(t.first.size());
  for (size_t m = 0; m < t.first.size(); ++m) {
    const int cnt = t.off[m + 1] - t.off[m];
    if (t.first[m] > 0xffff || cnt > 0x7fff) return "mel filter too wide";
    t.meta[m] = make_int2(t.first[m] | (cnt << 16), t.off[m]);
  }
  // fixed-width filter records (logmel_core.cuh: lm_mel_log fast path) when no filter is wider than 4 * LM_MEL_NV bins
  // and the zero-padded tail of every record stays inside the magnitude buffer
  {
    const size_t M = t.first.size();
    bool ok = cfg->codec == YMT3_CODEC_MELSPEC;
    for (size_t m = 0; ok && m < M; ++m) {
      const int cnt = t.off[m + 1] - t.off[m];
      // the walk reads whole groups of four bins: up to 3 slots past the filter's last bin, i.e. at most bins
      // 1025..1027, which map inside the magnitude buffer and are zeroed once per CTA (logmel.cu)
      if (cnt > 4 * LM_MEL_NV || t.first[m] + 4 * ((cnt + 3) / 4) > LM_NBINS + 3) ok = false;
    }
    t.rec_ok = ok ? 1 : 0;
    t.rec_w.assign(ok ? (size_t)LM_MEL_NV * M : 1, make_float4(0.f, 0.f, 0.f, 0.f));
    for (size_t m = 0; ok && m < M; ++m) {
      float w[4 * LM_MEL_NV] = {0.f};
      const int cnt = t.off[m + 1] - t.off[m];
      for (int j = 0; j < cnt; ++j) w[j] = t.wts[(size_t)t.off[m] + j];
      for (int v = 0; v < LM_MEL_NV; ++v)
        t.rec_w[(size_t)v * M + m] = make_float4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
    }
  }
  return nullptr;
}
