// CPU emulation of ymt3_logmel_kernel: runs the very same per-thread pass
// functions (yourmt3_b200/csrc/logmel_core.cuh) for tid = 0..127 sequentially,
// with a barrier == end of each tid loop. Test infrastructure only: lets the
// CPU-only test tier validate the FFT factorisation, smem index maps, pairing
// and mel banding without a GPU. Never used by the product path.
#include "../../yourmt3_b200/csrc/logmel_tables.h"
#include <vector>

extern "C" __attribute__((visibility("default"))) int lm_emu_run_wave(const ymt3_audio_cfg_t* cfg, const float* window,
                                                                     const float* fb, const float* audio, long long total,
                                                                     int B, int L, float* out);
extern "C" __attribute__((visibility("default"))) int lm_emu_run(const ymt3_audio_cfg_t* cfg, const float* window, const float* fb,
                          const float* audio, int B, int L, float* out) {
  return lm_emu_run_wave(cfg, window, fb, audio, (long long)B * L, B, L, out);
}
extern "C" __attribute__((visibility("default"))) int lm_emu_run_wave(const ymt3_audio_cfg_t* cfg, const float* window,
                                                                     const float* fb, const float* audio, long long total,
                                                                     int B, int L, float* out) {
  LmHostTables ht;
  if (lm_build_host_tables(cfg, fb, ht)) return 1;
  LmTables tb{window, ht.tw1.data(), ht.tw2.data(), ht.first.data(), ht.off.data(), ht.meta.data(), ht.wts.data(),
              ht.rec_w.data(), ht.rec_ok};
  const int T = 1 + L / cfg->hop_length, hop = cfg->hop_length, n_out = ht.n_out;
  const int pairs = (T + 1) / 2;
  std::vector<float2> bufA(LM_BUF_ELEMS), bufB(LM_BUF_ELEMS), mags(LM_MAG_ELEMS);
  std::vector<float> stage(hop + LM_NFFT + 4);
  for (int b = 0; b < B; ++b)
    for (int tp = 0; tp < pairs; ++tp) {
      const int tA = 2 * tp;
      const bool hasB = tA + 1 < T;
      const float* seg = audio + (size_t)b * L;
      const int startA = tA * hop - LM_NFFT / 2, startB = startA + hop;
      (void)startB;
      const long long remain = total - (long long)b * L;
      const int valid = remain >= L ? L : (remain > 0 ? (int)remain : 0);
      // the cooperative stage fill (edge-pair path of the kernel; the bulk-copy path stages the same samples)
      for (int tid = 0; tid < LM_THREADS; ++tid) lm_stage_fill(tid, seg, L, valid, startA, hop + LM_NFFT, stage.data());
      for (int tid = 0; tid < LM_THREADS; ++tid) {
        float w[16];
        float2 tw[16];
        for (int n1 = 0; n1 < 16; ++n1) {
          w[n1] = window[128 * n1 + tid];
          tw[n1] = tb.tw1[n1 * 128 + tid];
        }
        lm_pass1(tid, stage.data(), hop, hasB, w, tw, bufA.data());
      }
      for (int tid = 0; tid < LM_THREADS; ++tid) lm_pass2(tid, tb.tw2, bufA.data(), bufB.data());
      for (size_t i = 0; i < mags.size(); ++i) mags[i] = make_float2(-1.f, -1.f);   // poison: every bin must be written
      for (int i = lm_magaddr(LM_NBINS); i < LM_MAG_ELEMS; ++i) mags[i] = make_float2(0.f, 0.f);   // as the kernel does
      const bool spec = cfg->codec != YMT3_CODEC_MELSPEC;
      const bool take_sqrt = !spec && cfg->power_mode == 1;
      const LmOut oc = lm_out_consts(spec, cfg->power_mode, cfg->log_eps);
      for (int tid = 0; tid < LM_THREADS; ++tid) lm_pass3_mag(tid, bufB.data(), mags.data(), take_sqrt);
      for (int k = 0; k <= 1024; ++k)
        if (mags[lm_magaddr(k)].x < 0.f) return 2;
      float* outA = out + ((size_t)b * T + tA) * n_out;
      float* outB = hasB ? outA + n_out : nullptr;
      for (int tid = 0; tid < LM_THREADS; ++tid) {
        if (cfg->codec == YMT3_CODEC_MELSPEC) {
          LmMelRec rec;
          lm_mel_prefetch(tid, tb, n_out, rec);
          lm_mel_log(tid, tb, n_out, oc, rec, mags.data(), outA, outB);
        }
        else
          lm_spec_log(tid, cfg->spec_bin0, n_out, oc, mags.data(), outA, outB);
      }
    }
  return 0;
}
