// Fused log-(mel)spectrogram frontend for sm_100a. See logmel_core.cuh for the
// algorithm; DESIGN.md section 3 for the roofline reasoning.
#include "logmel_core.cuh"
#include "../../include/ymt3_b200.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "logmel_tables.h"

struct ymt3_frontend {
  ymt3_audio_cfg_t cfg;
  int n_out;            // feature width F
  float* d_window;
  float2* d_tw1;
  float2* d_tw2;
  int* d_mel_first;
  int* d_mel_off;
  int2* d_mel_meta;
  float* d_mel_w;
  // staging for the *_host entry point (grown on demand)
  float* d_stage_in;
  float* d_stage_out;
  size_t stage_in_bytes, stage_out_bytes;
};

__global__ void __launch_bounds__(LM_THREADS, 6)
ymt3_logmel_kernel(LmTables tb, const float* __restrict__ audio, int64_t total_samples, float* __restrict__ out, int L,
                   int T, int hop, int pairs_per_seg, int chunk, int total_pairs, int codec,
                   int n_out, int spec_bin0, int power_mode, float eps) {
  __shared__ __align__(16) float2 bufA[LM_BUF_ELEMS];
  __shared__ __align__(16) float2 bufB[LM_BUF_ELEMS];
  float2* mags = bufA;   // bufA is dead after pass 2; the interleaved magnitudes (LM_MAG_ELEMS <= LM_BUF_ELEMS) reuse it
  const int tid = threadIdx.x;

  float w[16];   // loop-invariant window column of this thread
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) w[n1] = __ldg(tb.window + 128 * n1 + tid);

  const bool spec = codec != YMT3_CODEC_MELSPEC;
  const bool take_sqrt = !spec && power_mode == 1;   // the linear-frequency codec never needs the square root
  const LmOut oc = lm_out_consts(spec, power_mode, eps);

  const int p0 = blockIdx.x * chunk;
  const int p1 = min(p0 + chunk, total_pairs);

  for (int p = p0; p < p1; ++p) {
    const int b = p / pairs_per_seg;
    const int tA = 2 * (p - b * pairs_per_seg);
    const bool hasB = (tA + 1) < T;
    const float* seg = audio + (size_t)b * L;
    const int64_t remain = total_samples - (int64_t)b * L;   // waveform mode: the last segment may be partial
    const int valid = remain >= L ? L : (remain > 0 ? (int)remain : 0);
    const int startA = tA * hop - LM_NFFT / 2;
    const int startB = startA + hop;

    lm_pass1(tid, seg, L, valid, startA, startB, hasB, w, tb.tw1, bufA);
    __syncthreads();
    lm_pass2(tid, tb.tw2, bufA, bufB);
    __syncthreads();
    lm_pass3_mag(tid, bufB, mags, take_sqrt);
    __syncthreads();
    float* outA = out + ((size_t)b * T + tA) * n_out;
    float* outB = hasB ? outA + n_out : nullptr;
    if (!spec)
      lm_mel_log(tid, tb, n_out, oc, mags, outA, outB);
    else
      lm_spec_log(tid, spec_bin0, n_out, oc, mags, outA, outB);
    __syncthreads();   // mags alias bufA, which the next pair's pass 1 overwrites
  }
}

extern "C" int ymt3_frontend_create(const ymt3_audio_cfg_t* cfg, const float* window_host,
                                    const float* fb_host, ymt3_frontend_t** out) {
  YMT3_REQUIRE(cfg && window_host && out, "frontend_create: null argument");
  YMT3_REQUIRE(cfg->n_fft == LM_NFFT, "frontend_create: only n_fft=2048 is supported (got %d)",
               cfg->n_fft);
  YMT3_REQUIRE(cfg->hop_length > 0 && cfg->hop_length <= LM_NFFT, "frontend_create: bad hop %d",
               cfg->hop_length);
  YMT3_REQUIRE(cfg->power_mode == 1 || cfg->power_mode == 2,
               "frontend_create: power_mode must be 1 or 2");
  YMT3_REQUIRE(cfg->codec == YMT3_CODEC_MELSPEC || cfg->codec == YMT3_CODEC_SPEC,
               "frontend_create: bad codec %d", cfg->codec);
  ymt3_frontend* fe = (ymt3_frontend*)calloc(1, sizeof(ymt3_frontend));
  YMT3_REQUIRE(fe, "frontend_create: out of host memory");
  fe->cfg = *cfg;

  LmHostTables ht;
  {
    const char* err = lm_build_host_tables(cfg, fb_host, ht);
    if (err) {
      free(fe);
      ymt3_set_error("frontend_create: %s", err);
      return YMT3_ERR_INVALID;
    }
  }
  fe->n_out = ht.n_out;
  std::vector<float2>& tw1 = ht.tw1;
  std::vector<float2>& tw2 = ht.tw2;
  std::vector<int>& first = ht.first;
  std::vector<int>& off = ht.off;
  std::vector<float>& wts = ht.wts;

#define FE_UPLOAD(dst, vec, type)                                                              \
  YMT3_CUDA_CHECK(cudaMalloc((void**)&(dst), (vec).size() * sizeof(type)));                    \
  YMT3_CUDA_CHECK(cudaMemcpy((dst), (vec).data(), (vec).size() * sizeof(type),                 \
                             cudaMemcpyHostToDevice));
  std::vector<float> win(window_host, window_host + LM_NFFT);
  FE_UPLOAD(fe->d_window, win, float);
  FE_UPLOAD(fe->d_tw1, tw1, float2);
  FE_UPLOAD(fe->d_tw2, tw2, float2);
  FE_UPLOAD(fe->d_mel_first, first, int);
  FE_UPLOAD(fe->d_mel_off, off, int);
  FE_UPLOAD(fe->d_mel_meta, ht.meta, int2);
  FE_UPLOAD(fe->d_mel_w, wts, float);
#undef FE_UPLOAD
  *out = fe;
  return YMT3_OK;
}

extern "C" int ymt3_frontend_destroy(ymt3_frontend_t* fe) {
  if (!fe) return YMT3_OK;
  cudaFree(fe->d_window);
  cudaFree(fe->d_tw1);
  cudaFree(fe->d_tw2);
  cudaFree(fe->d_mel_first);
  cudaFree(fe->d_mel_off);
  cudaFree(fe->d_mel_meta);
  cudaFree(fe->d_mel_w);
  cudaFree(fe->d_stage_in);
  cudaFree(fe->d_stage_out);
  free(fe);
  return YMT3_OK;
}

extern "C" int64_t ymt3_frontend_num_frames(const ymt3_frontend_t* fe, int64_t L) {
  return fe ? 1 + L / fe->cfg.hop_length : -1;
}
extern "C" int64_t ymt3_frontend_num_features(const ymt3_frontend_t* fe) {
  return fe ? fe->n_out : -1;
}

static int logmel_launch(ymt3_frontend_t* fe, const float* audio_dev, int64_t total_samples, int64_t B, int64_t L,
                         float* out_dev, void* stream);

extern "C" int ymt3_logmel_f32(ymt3_frontend_t* fe, const float* audio_dev, int64_t B, int64_t L,
                               float* out_dev, void* stream) {
  return logmel_launch(fe, audio_dev, B * L, B, L, out_dev, stream);
}

extern "C" int64_t ymt3_num_segments(int64_t n_samples, int64_t seg_len) {
  if (seg_len <= 0) return -1;
  const int64_t n = (n_samples + seg_len - 1) / seg_len;
  return n < 1 ? 1 : n;
}

// Waveform entry: segmentation (upstream utils/audio.py slice_padded_array: hop == length, zero-padded tail)
// is fused into the kernel's loads -- no sliced / padded copy of the audio is ever materialised.
extern "C" int ymt3_logmel_waveform_f32(ymt3_frontend_t* fe, const float* wave_dev, int64_t n_samples, int64_t seg_len,
                                        float* out_dev, void* stream) {
  YMT3_REQUIRE(n_samples >= 0 && seg_len > 0, "logmel_waveform: bad sizes");
  return logmel_launch(fe, wave_dev, n_samples, ymt3_num_segments(n_samples, seg_len), seg_len, out_dev, stream);
}

static int logmel_launch(ymt3_frontend_t* fe, const float* audio_dev, int64_t total_samples, int64_t B, int64_t L,
                         float* out_dev, void* stream) {
  YMT3_REQUIRE(fe, "logmel: null handle");
  YMT3_REQUIRE(B >= 0, "logmel: negative batch");
  if (B == 0) return YMT3_OK;   // empty batch: nothing to do (pointers may be null)
  YMT3_REQUIRE(audio_dev && out_dev, "logmel: null buffer");
  // torch.stft reflect padding needs pad < L (torch/functional.py:675-680)
  YMT3_REQUIRE(L > LM_NFFT / 2, "logmel: segment length %lld must exceed n_fft/2 = %d",
               (long long)L, LM_NFFT / 2);
  YMT3_REQUIRE(L < (1ll << 30), "logmel: segment length too large");
  const int T = (int)(1 + L / fe->cfg.hop_length);
  const int pairs_per_seg = (T + 1) / 2;
  const int64_t total64 = B * (int64_t)pairs_per_seg;
  YMT3_REQUIRE(total64 < (1ll << 31), "logmel: too many frames in one call");
  const int total = (int)total64;
  // chunk consecutive frame pairs of a segment onto one CTA (L1 reuse of the
  // 16x-overlapping audio) while keeping >= 2 waves of 6 CTAs/SM when possible.
  const int target_ctas = ymt3_num_sms() * 6 * 2;
  int chunk = total / target_ctas;
  if (chunk < 1) chunk = 1;
  if (chunk > 8) chunk = 8;
  const int grid = ymt3_div_up(total, chunk);
  LmTables tb{fe->d_window, fe->d_tw1, fe->d_tw2, fe->d_mel_first, fe->d_mel_off, fe->d_mel_meta, fe->d_mel_w};
  ymt3_logmel_kernel<<<grid, LM_THREADS, 0, (cudaStream_t)stream>>>(
      tb, audio_dev, total_samples, out_dev, (int)L, T, fe->cfg.hop_length, pairs_per_seg, chunk, total,
      fe->cfg.codec, fe->n_out, fe->cfg.spec_bin0, fe->cfg.power_mode, fe->cfg.log_eps);
  YMT3_CUDA_CHECK(cudaGetLastError());
  return YMT3_OK;
}

extern "C" int ymt3_logmel_host_f32(ymt3_frontend_t* fe, const float* audio_host, int64_t B,
                                    int64_t L, float* out_host, void* stream) {
  YMT3_REQUIRE(fe && out_host, "logmel_host: null argument");
  if (B == 0) return YMT3_OK;
  YMT3_REQUIRE(audio_host && B > 0, "logmel_host: bad input");
  const int T = (int)(1 + L / fe->cfg.hop_length);
  const size_t in_bytes = (size_t)B * L * sizeof(float);
  const size_t out_bytes = (size_t)B * T * fe->n_out * sizeof(float);
  if (in_bytes > fe->stage_in_bytes) {
    cudaFree(fe->d_stage_in);
    fe->d_stage_in = nullptr;
    fe->stage_in_bytes = 0;
    YMT3_CUDA_CHECK(cudaMalloc((void**)&fe->d_stage_in, in_bytes));
    fe->stage_in_bytes = in_bytes;
  }
  if (out_bytes > fe->stage_out_bytes) {
    cudaFree(fe->d_stage_out);
    fe->d_stage_out = nullptr;
    fe->stage_out_bytes = 0;
    YMT3_CUDA_CHECK(cudaMalloc((void**)&fe->d_stage_out, out_bytes));
    fe->stage_out_bytes = out_bytes;
  }
  cudaStream_t s = (cudaStream_t)stream;
  YMT3_CUDA_CHECK(cudaMemcpyAsync(fe->d_stage_in, audio_host, in_bytes, cudaMemcpyHostToDevice, s));
  int rc = ymt3_logmel_f32(fe, fe->d_stage_in, B, L, fe->d_stage_out, stream);
  if (rc != YMT3_OK) return rc;
  YMT3_CUDA_CHECK(cudaMemcpyAsync(out_host, fe->d_stage_out, out_bytes, cudaMemcpyDeviceToHost, s));
  YMT3_CUDA_CHECK(cudaStreamSynchronize(s));
  return YMT3_OK;
}
