// Fused STFT(2048) -> |X|^p -> mel / linear bins -> log core, written as
// per-thread "pass" functions so the same code runs (a) inside the sm_100a
// kernel with 128 threads per frame pair and __syncthreads() between passes,
// and (b) on the host, thread by thread, in tests/host_emu (no GPU needed).
//
// Reference semantics being reproduced (installed deps of the upstream path,
// see SURVEY.md 3.2): torchaudio/functional/functional.py:119-144
// (spectrogram: reflect pad, frame, window, rFFT, abs()/pow), torchaudio/
// transforms/_transforms.py:417 (mel matmul), upstream model/spectrogram.py
// log(clamp(x, eps)) [RECALL].
//
// Algorithm: two real frames (t, t+1) are packed as re/im of one complex
// 2048-point DIF FFT, factorised 16 x 16 x 8:
//   n = 128*n1 + 8*n2 + n3,   k = k1 + 16*k2 + 256*k3
//   pass1: radix-16 over n1, twiddle W_2048^(m*k1), m = 8*n2+n3 = tid
//   pass2: radix-16 over n2, twiddle W_128^(n3*k2)
//   pass3: radix-8  over n3
// then X_A[k] = (Z[k] + conj Z[N-k])/2, X_B[k] = (Z[k] - conj Z[N-k])/(2i).
#pragma once
#include "common.cuh"

#define LM_NFFT 2048
#define LM_NBINS 1025
#define LM_THREADS 128
// shared-memory geometry (units: float2)
#define LM_S1 129                 // pass-1 output row stride   [k1][m]
#define LM_S3 274                 // pass-2 output plane stride [n3][k1*17+k2]
#define LM_SZ 272                 // pass-3 output plane stride [k3][k1*17+k2]
#define LM_BUF_ELEMS 2192         // >= max(16*129, 8*274, 8*272)

struct LmTables {
  const float*  window;     // [2048] analysis window (win_length == n_fft)
  const float2* tw1;        // [16][128]  W_2048^(m*k1)
  const float2* tw2;        // [8][16]    W_128^(n3*k2)
  const int*    mel_first;  // [n_mels]   first nonzero bin of filter m
  const int*    mel_off;    // [n_mels+1] prefix offsets into mel_w
  const float*  mel_w;      // packed nonzero weights, ascending bin order
};

YMT3_HD float2 lm_cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
YMT3_HD float2 lm_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
YMT3_HD float2 lm_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// forward DFT-4, natural order in and out
YMT3_HD void lm_fft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = lm_add(a0, a2), t1 = lm_sub(a0, a2);
  float2 t2 = lm_add(a1, a3), t3 = lm_sub(a1, a3);
  a0 = lm_add(t0, t2);
  a2 = lm_sub(t0, t2);
  a1 = make_float2(t1.x + t3.y, t1.y - t3.x);  // t1 - i*t3
  a3 = make_float2(t1.x - t3.y, t1.y + t3.x);  // t1 + i*t3
}

#define LM_C1 0.92387953251128674f   // cos(pi/8)
#define LM_S1C 0.38268343236508977f  // sin(pi/8)
#define LM_R2 0.70710678118654752f   // sqrt(1/2)

// forward DFT-16. Input v[n] natural. Output: X[a + 4*b] is left in v[4*a + b].
YMT3_HD void lm_fft16(float2 (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) lm_fft4(v[j], v[j + 4], v[j + 8], v[j + 12]);
  // twiddles W_16^(j*a) on v[j + 4a]
  const float2 w1 = make_float2(LM_C1, -LM_S1C);
  const float2 w2 = make_float2(LM_R2, -LM_R2);
  const float2 w3 = make_float2(LM_S1C, -LM_C1);
  const float2 w6 = make_float2(-LM_R2, -LM_R2);
  const float2 w9 = make_float2(-LM_C1, LM_S1C);
  v[1 + 4] = lm_cmul(v[1 + 4], w1);
  v[1 + 8] = lm_cmul(v[1 + 8], w2);
  v[1 + 12] = lm_cmul(v[1 + 12], w3);
  v[2 + 4] = lm_cmul(v[2 + 4], w2);
  v[2 + 8] = make_float2(v[2 + 8].y, -v[2 + 8].x);  // * W_16^4 = -i
  v[2 + 12] = lm_cmul(v[2 + 12], w6);
  v[3 + 4] = lm_cmul(v[3 + 4], w3);
  v[3 + 8] = lm_cmul(v[3 + 8], w6);
  v[3 + 12] = lm_cmul(v[3 + 12], w9);
#pragma unroll
  for (int a = 0; a < 4; ++a) lm_fft4(v[4 * a], v[4 * a + 1], v[4 * a + 2], v[4 * a + 3]);
}

// forward DFT-8. Input v[n] natural. Output: X[a + 4*b] is left in v[2*a + b].
YMT3_HD void lm_fft8(float2 (&v)[8]) {
  lm_fft4(v[0], v[2], v[4], v[6]);
  lm_fft4(v[1], v[3], v[5], v[7]);
  // v[j + 2a]; multiply j=1 terms by W_8^a
  v[3] = lm_cmul(v[3], make_float2(LM_R2, -LM_R2));
  v[5] = make_float2(v[5].y, -v[5].x);
  v[7] = lm_cmul(v[7], make_float2(-LM_R2, -LM_R2));
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    float2 s = lm_add(v[2 * a], v[2 * a + 1]);
    float2 d = lm_sub(v[2 * a], v[2 * a + 1]);
    v[2 * a] = s;
    v[2 * a + 1] = d;
  }
}

// torch.stft(center=True, pad_mode="reflect") index map (torch/functional.py:675-680)
YMT3_HD int lm_reflect(int i, int L) {
  if (i < 0) i = -i;
  if (i >= L) i = 2 * (L - 1) - i;
  return i;
}

// pass 1: load frame A (real) and frame B (imag), window, radix-16 over n1, twiddle.
// seg: this segment's L samples. startA/startB: index of sample n=0 of each frame
// in un-padded coordinates (may be negative / beyond L -> reflect).
YMT3_HD void lm_pass1(int tid, const float* __restrict__ seg, int L, int startA, int startB,
                      bool hasB, const float (&w)[16], const float2* __restrict__ tw1,
                      float2* __restrict__ bufA) {
  float2 v[16];
  const bool interior = (startA >= 0) && (startB + LM_NFFT <= L);
  if (interior) {
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      int n = 128 * n1 + tid;
      float xa = seg[startA + n];
      float xb = hasB ? seg[startB + n] : 0.f;
      v[n1] = make_float2(xa * w[n1], xb * w[n1]);
    }
  } else {
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      int n = 128 * n1 + tid;
      float xa = seg[lm_reflect(startA + n, L)];
      float xb = hasB ? seg[lm_reflect(startB + n, L)] : 0.f;
      v[n1] = make_float2(xa * w[n1], xb * w[n1]);
    }
  }
  lm_fft16(v);
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    int k1 = (r >> 2) + 4 * (r & 3);
    float2 o = v[r];
    if (k1 != 0) o = lm_cmul(o, tw1[k1 * 128 + tid]);
    bufA[k1 * LM_S1 + tid] = o;
  }
}

// pass 2: thread q -> (k1 = q & 15, n3 = q >> 4); radix-16 over n2; twiddle W_128^(n3*k2)
YMT3_HD void lm_pass2(int tid, const float2* __restrict__ tw2, const float2* __restrict__ bufA,
                      float2* __restrict__ bufB) {
  const int k1 = tid & 15, n3 = tid >> 4;
  float2 v[16];
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) v[n2] = bufA[k1 * LM_S1 + 8 * n2 + n3];
  lm_fft16(v);
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    int k2 = (r >> 2) + 4 * (r & 3);
    float2 o = v[r];
    if (k2 != 0) o = lm_cmul(o, tw2[n3 * 16 + k2]);
    bufB[n3 * LM_S3 + k1 * 17 + k2] = o;
  }
}

// pass 3: thread j handles c = j and j + 128, c = k1*16 + k2; radix-8 over n3.
// Z[k1 + 16*k2 + 256*k3] is stored at bufA[k3*LM_SZ + k1*17 + k2].
YMT3_HD void lm_pass3(int tid, const float2* __restrict__ bufB, float2* __restrict__ bufA) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = tid + 128 * h;
    const int k1 = c >> 4, k2 = c & 15;
    float2 v[8];
#pragma unroll
    for (int n3 = 0; n3 < 8; ++n3) v[n3] = bufB[n3 * LM_S3 + k1 * 17 + k2];
    lm_fft8(v);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      int k3 = (r >> 1) + 4 * (r & 1);
      bufA[k3 * LM_SZ + k1 * 17 + k2] = v[r];
    }
  }
}

YMT3_HD int lm_zaddr(int k) { return (k >> 8) * LM_SZ + (k & 15) * 17 + ((k >> 4) & 15); }

// magnitude / power of both packed frames for bins 0..1024
// power_mode: 1 -> |X|, 2 -> |X|^2 (torchaudio functional.py:141-144)
YMT3_HD void lm_mag(int tid, const float2* __restrict__ bufA, float* __restrict__ magA,
                    float* __restrict__ magB, int power_mode) {
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    int k = tid + 128 * i;
    if (k > 1024) break;
    float2 z = bufA[lm_zaddr(k)];
    float2 zc = bufA[lm_zaddr((LM_NFFT - k) & (LM_NFFT - 1))];
    // X_A = ((a+c) + i(b-d))/2 ; X_B = ((b+d) - i(a-c))/2 with z=a+ib, zc=c+id
    float ar = 0.5f * (z.x + zc.x), ai = 0.5f * (z.y - zc.y);
    float br = 0.5f * (z.y + zc.y), bi = 0.5f * (z.x - zc.x);
    float pa = ar * ar + ai * ai, pb = br * br + bi * bi;
    if (power_mode == 1) {
      pa = sqrtf(pa);
      pb = sqrtf(pb);
    }
    magA[k] = pa;
    magB[k] = pb;
  }
}

// mel projection + log for both frames. Filter m: sum_j mel_w[off[m]+j] * mag[first[m]+j].
YMT3_HD void lm_mel_log(int tid, const LmTables& tb, int n_mels, float eps,
                        const float* __restrict__ magA, const float* __restrict__ magB,
                        float* __restrict__ outA, float* __restrict__ outB) {
  for (int m = tid; m < n_mels; m += LM_THREADS) {
    const int first = tb.mel_first[m];
    const int o0 = tb.mel_off[m], o1 = tb.mel_off[m + 1];
    float sa = 0.f, sb = 0.f;
    for (int j = o0; j < o1; ++j) {
      float wgt = tb.mel_w[j];
      int k = first + (j - o0);
      sa = fmaf(wgt, magA[k], sa);
      sb = fmaf(wgt, magB[k], sb);
    }
    outA[m] = logf(fmaxf(sa, eps));
    if (outB) outB[m] = logf(fmaxf(sb, eps));
  }
}

// linear-frequency ("spec" codec) log output: bins [bin0, bin0 + n_out)
YMT3_HD void lm_spec_log(int tid, int bin0, int n_out, float eps, const float* __restrict__ magA,
                         const float* __restrict__ magB, float* __restrict__ outA,
                         float* __restrict__ outB) {
  for (int f = tid; f < n_out; f += LM_THREADS) {
    outA[f] = logf(fmaxf(magA[bin0 + f], eps));
    if (outB) outB[f] = logf(fmaxf(magB[bin0 + f], eps));
  }
}
