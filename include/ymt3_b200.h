/*
 * ymt3_b200.h -- C ABI of the B200-native YourMT3 inference hot path.
 *
 * Every entry point is plain C: pointers + sizes, an int status (0 = ok) and a
 * thread-local error string (ymt3_last_error).  No torch / C++ types cross the
 * boundary.  All *_dev pointers are CUDA device pointers owned by the caller;
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 * No entry point synchronises the device unless its name ends in `_host`.
 *
 * The reference (richhiey/YourMT3 -> upstream mimbres/YourMT3, amt/src/model/)
 * is pure Python and has no FFI; each function below names the reference
 * Python interface it replaces.  The mounted fork holds only README.md, so
 * upstream paths are given by name ([RECALL]) and the
 * arithmetic is cited into the installed dependencies the upstream path calls
 * (torchaudio 2.11 / transformers 5.5, abbreviated TA/ and HF/).
 */
#ifndef YMT3_B200_H
#define YMT3_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YMT3_ABI_VERSION 1

#if defined(__GNUC__)
#define YMT3_API __attribute__((visibility("default")))
#else
#define YMT3_API
#endif

/* status codes */
#define YMT3_STATUS_OK 0
#define YMT3_STATUS_INVALID 1
#define YMT3_STATUS_CUDA 2
#define YMT3_STATUS_UNSUPPORTED 3

YMT3_API const char* ymt3_last_error(void);
YMT3_API int ymt3_abi_version(void);
/* fills name (<= cap bytes), sm count, compute capability major/minor of the current device */
YMT3_API int ymt3_device_info(char* name, int cap, int* num_sms, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------- *
 * Frontend: waveform -> log-(mel)spectrogram.
 * Replaces upstream model/spectrogram.py: get_spectrogram_layer_from_audio_cfg(),
 * Melspectrogram.forward / Spectrogram.forward, i.e.
 *   TA/transforms/_transforms.py:621-631 (MelSpectrogram.forward)
 *   TA/functional/functional.py:54-145   (spectrogram: pad, stft, abs/pow)
 *   TA/transforms/_transforms.py:417     (MelScale matmul)
 *   + log(clamp(x, min=eps)).
 * ------------------------------------------------------------------------- */
#define YMT3_CODEC_MELSPEC 0
#define YMT3_CODEC_SPEC 1

typedef struct ymt3_audio_cfg {
  int32_t n_fft;       /* 2048 (only value supported by the fused kernel)           */
  int32_t hop_length;  /* 128 (melspec / MT3) or 300 (spec / Perceiver-TF)          */
  int32_t codec;       /* YMT3_CODEC_MELSPEC | YMT3_CODEC_SPEC                      */
  int32_t n_mels;      /* melspec: number of mel filters (512)                      */
  int32_t spec_bin0;   /* spec: first linear bin kept (1 = drop DC)                 */
  int32_t spec_bins;   /* spec: number of bins kept (1024)                          */
  int32_t power_mode;  /* 1 -> |X|, 2 -> |X|^2                                      */
  float log_eps;       /* clamp floor before log                                    */
} ymt3_audio_cfg_t;

typedef struct ymt3_frontend ymt3_frontend_t;

/* window_host: n_fft floats (module buffer `spectrogram.window`).
 * fb_host: melspec only, dense (n_fft/2+1, n_mels) row-major filterbank (module
 * buffer `mel_scale.fb`, TA/functional/functional.py:518-587); nonzeros of each
 * column must be contiguous in frequency (true for triangular filters). */
YMT3_API int ymt3_frontend_create(const ymt3_audio_cfg_t* cfg, const float* window_host,
                         const float* fb_host, ymt3_frontend_t** out);
YMT3_API int ymt3_frontend_destroy(ymt3_frontend_t* fe);
/* number of frames for L samples (center=True): 1 + L / hop */
YMT3_API int64_t ymt3_frontend_num_frames(const ymt3_frontend_t* fe, int64_t L);
/* feature width F of the output */
YMT3_API int64_t ymt3_frontend_num_features(const ymt3_frontend_t* fe);

/* audio_dev: (B, L) f32 contiguous. out_dev: (B, T, F) f32 contiguous. */
YMT3_API int ymt3_logmel_f32(ymt3_frontend_t* fe, const float* audio_dev, int64_t B, int64_t L,
                    float* out_dev, void* stream);
/* Same with HOST buffers: H2D + kernel + D2H + stream sync (the end-to-end call). */
YMT3_API int ymt3_logmel_host_f32(ymt3_frontend_t* fe, const float* audio_host, int64_t B, int64_t L,
                         float* out_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YMT3_B200_H */
