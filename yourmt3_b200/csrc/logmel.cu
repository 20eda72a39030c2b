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
  float4* d_rec_w;
  int rec_ok;
  // staging for the *_host entry point (grown on demand)
  float* d_stage_in;
  float* d_stage_out;
  size_t stage_in_bytes, stage_out_bytes;
};

// ---- mbarrier / bulk-copy helpers (1-D TMA: cp.async.bulk global -> shared, completion on an mbarrier) ----
__device__ __forceinline__ uint32_t lm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lm_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lm_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void lm_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lm_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lm_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LM_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra LM_DONE;\n\t"
      "bra LM_WAIT;\n\t"
      "LM_DONE:\n\t"
      "}" ::"r"(lm_smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void lm_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   lm_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(lm_smem_u32(bar))
               : "memory");
}

// One CTA (128 threads) per run of `chunk` consecutive frame pairs.  Per pair: [staged samples ready] pass 1 ->
// barrier -> (one thread launches the bulk copy of the NEXT pair's samples) pass 2 -> barrier -> pass 3 + magnitudes
// -> barrier -> output stage -> barrier.  Shared memory: two FFT exchange buffers (static, 35 KB) + the sample
// stage (dynamic, hop + 2048 + 4 floats).  4 CTAs / SM (register twiddles: <= 128 registers per thread).
__global__ void __launch_bounds__(LM_THREADS, 4)
ymt3_logmel_kernel(LmTables tb, const float* __restrict__ audio, int64_t total_samples, float* __restrict__ out, int L,
                   int T, int hop, int pairs_per_seg, int chunk, int total_pairs, int codec,
                   int n_out, int spec_bin0, int power_mode, float eps, int bulk_ok) {
  __shared__ __align__(16) float2 bufA[LM_BUF_ELEMS];
  __shared__ __align__(16) float2 bufB[LM_BUF_ELEMS];
  __shared__ __align__(16) float2 mags[LM_MAG_ELEMS];   // own buffer: no barrier between the output stage and the next pass 1
  __shared__ __align__(8) uint64_t stage_bar;
  extern __shared__ __align__(16) float stage[];   // hop + 2048 (+ <= 3 alignment) samples
  const int tid = threadIdx.x;

  // loop-invariant per-thread constants: window column and pass-1 twiddles W_2048^(tid * k1)
  float w[16];
  float2 tw[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    w[n1] = __ldg(tb.window + 128 * n1 + tid);
    tw[n1] = __ldg(tb.tw1 + n1 * 128 + tid);
  }
  const bool spec = codec != YMT3_CODEC_MELSPEC;
  const bool take_sqrt = !spec && power_mode == 1;   // the linear-frequency codec never needs the square root
  const LmOut oc = lm_out_consts(spec, power_mode, eps);
  const int span = hop + LM_NFFT;                    // samples one pair covers

  if (tid < LM_MAG_ELEMS - lm_magaddr(LM_NBINS)) mags[lm_magaddr(LM_NBINS) + tid] = make_float2(0.f, 0.f);   // slots past bin 1024
  if (tid == 0) {
    lm_mbar_init(&stage_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int p0 = blockIdx.x * chunk;
  const int n_pairs = min(p0 + chunk, total_pairs) - p0;
  uint32_t phase = 0;

  // (segment, pair-in-segment) of the NEXT pair to request, advanced incrementally (one division per CTA)
  int nb = p0 / pairs_per_seg, ntp = p0 - nb * pairs_per_seg;
  // request the samples of pair (nb, ntp) into the stage (which must be free) and advance; returns the pair's
  // geometry for later consumption.  `bulk`: ONE aligned bulk copy (interior pair, both frames present, 16-byte
  // aligned window inside the buffer); `off` = frame A's offset in the staged window
  struct Geo { int b, tA, off; bool hasB, bulk; };
  auto request_next = [&]() {
    Geo g;
    g.b = nb;
    g.tA = 2 * ntp;
    g.hasB = (g.tA + 1) < T;
    const int64_t seg0 = (int64_t)nb * L;
    const int64_t remain = total_samples - seg0;               // waveform mode: the last segment may be partial
    const int valid = remain >= L ? L : (remain > 0 ? (int)remain : 0);
    const int startA = g.tA * hop - LM_NFFT / 2;
    const int64_t gi = seg0 + startA;                          // global sample index of frame A's first sample
    const int64_t a0 = gi & ~(int64_t)3;
    g.off = (int)(gi - a0);
    const int n4 = (g.off + span + 3) & ~3;
    g.bulk = bulk_ok && g.hasB && startA >= 0 && startA + span <= valid && a0 + n4 <= total_samples;
    if (g.bulk) {
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads / writes of the stage
        lm_mbar_expect_tx(&stage_bar, (uint32_t)n4 * 4u);
        lm_bulk_g2s(stage, audio + a0, (uint32_t)n4 * 4u, &stage_bar);
      }
    } else {
      g.off = 0;
      lm_stage_fill(tid, audio + seg0, L, valid, startA, span, stage);   // visible after the next barrier
    }
    if (++ntp == pairs_per_seg) { ntp = 0; ++nb; }
    return g;
  };

  Geo g{};
  if (n_pairs > 0) {
    g = request_next();
    __syncthreads();
  }
  for (int i = 0; i < n_pairs; ++i) {
    if (g.bulk) {
      lm_mbar_wait(&stage_bar, phase);
      phase ^= 1;
    }
    lm_pass1(tid, stage + g.off, hop, g.hasB, w, tw, bufA);
    __syncthreads();                      // bufA complete; the stage is free again
    const Geo cur = g;
    if (i + 1 < n_pairs) g = request_next();
    lm_pass2(tid, tb.tw2, bufA, bufB);
    __syncthreads();
    lm_pass3_mag(tid, bufB, mags, take_sqrt);
    LmMelRec rec;
    if (!spec) lm_mel_prefetch(tid, tb, n_out, rec);   // table loads in flight across the barrier
    __syncthreads();
    float* outA = out + ((size_t)cur.b * T + cur.tA) * n_out;
    float* outB = cur.hasB ? outA + n_out : nullptr;
    if (!spec)
      lm_mel_log(tid, tb, n_out, oc, rec, mags, outA, outB);
    else
      lm_spec_log(tid, spec_bin0, n_out, oc, mags, outA, outB);
    // no barrier here: the next pair's pass 1 / pass 2 touch the stage, bufA and bufB only, and two barriers separate
    // this output stage from the next write of `mags` (and a cooperative stage fill from its first read)
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
  FE_UPLOAD(fe->d_rec_w, ht.rec_w, float4);
  fe->rec_ok = ht.rec_ok;
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
  cudaFree(fe->d_rec_w);
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
  // a CTA walks `chunk` consecutive frame pairs (per-thread constants and the staging pipeline are amortised over
  // the run).  All CTAs cost the same, so the grid is sized to a WHOLE number of waves of the 4 x SMs resident CTAs:
  // k waves of at most 20 pairs each (728 segments x 55 pairs: 2503 CTAs of 16 pairs were 4.2 waves = 5 rounds;
  // 2356 CTAs of 17 are 3.98)
  const int slots = ymt3_num_sms() * 4;
  const int waves = ymt3_div_up(total, (int64_t)slots * 20);
  int chunk = ymt3_div_up(total, (int64_t)slots * waves);
  if (chunk < 1) chunk = 1;
  const int grid = ymt3_div_up(total, chunk);
  const size_t stage_bytes = (size_t)(fe->cfg.hop_length + LM_NFFT + 4) * sizeof(float);
  {   // static 43.8 KB + dynamic stage exceed the 48 KB default: opt in (per device, grown with the hop)
    static int attr_hop[64] = {0};
    int dev = 0;
    YMT3_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || attr_hop[dev] < fe->cfg.hop_length) {
      YMT3_CUDA_CHECK(cudaFuncSetAttribute(ymt3_logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
      if (dev >= 0 && dev < 64) attr_hop[dev] = fe->cfg.hop_length;
    }
  }
  // bulk copies need 16-byte aligned global addresses: the window start is rounded down to a multiple of 4 samples
  // relative to the buffer, so the buffer itself must be 16-byte aligned (else every pair takes the cooperative fill)
  const int bulk_ok = (((uintptr_t)audio_dev) & 15) == 0;
  LmTables tb{fe->d_window, fe->d_tw1, fe->d_tw2, fe->d_mel_first, fe->d_mel_off, fe->d_mel_meta, fe->d_mel_w,
              fe->d_rec_w, fe->rec_ok};
  ymt3_logmel_kernel<<<grid, LM_THREADS, stage_bytes, (cudaStream_t)stream>>>(
      tb, audio_dev, total_samples, out_dev, (int)L, T, fe->cfg.hop_length, pairs_per_seg, chunk, total,
      fe->cfg.codec, fe->n_out, fe->cfg.spec_bin0, fe->cfg.power_mode, fe->cfg.log_eps, bulk_ok);
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
