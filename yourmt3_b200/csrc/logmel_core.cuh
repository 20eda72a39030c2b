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
//
// v2 (LSU diet, ncu showed the LSU data pipe at 74 % and issue slots at 64 %): with hop == 128 frame t+1
// is frame t shifted by one 128-sample block, so 17 loads feed both frames; pass 3 gives each thread BOTH members of every (k, N-k) pair, so magnitudes are formed in registers
// and the spectrum is never written back (one barrier and 32 KB of smem traffic less per pair); magnitudes
// of the two frames are stored interleaved (float2) for the banded mel projection.
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
#define LM_MAG_ELEMS 1092         // interleaved (magA, magB) float2 per bin, padded: addr(bin) = bin + (bin >> 4)

struct LmTables {
  const float*  window;     // [2048] analysis window (win_length == n_fft)
  const float2* tw1;        // [16][128]  W_2048^(m*k1)
  const float2* tw2;        // [8][16]    W_128^(n3*k2)
  const int*    mel_first;  // [n_mels]   first nonzero bin of filter m
  const int*    mel_off;    // [n_mels+1] prefix offsets into mel_w
  const int2*   mel_meta;   // [n_mels]   (first | count << 16, offset)
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

// powers 1..15 of a unit twiddle from its 1st, 2nd, 4th and 8th powers (table values): <= 3 products deep
YMT3_HD void lm_tw_powers(float2 w1, float2 w2, float2 w4, float2 w8, float2 (&p)[16]) {
  p[0] = make_float2(1.f, 0.f);
  p[1] = w1; p[2] = w2; p[4] = w4; p[8] = w8;
  p[3] = lm_cmul(w2, w1);
  p[5] = lm_cmul(w4, w1);
  p[6] = lm_cmul(w4, w2);
  p[7] = lm_cmul(w4, p[3]);
  p[9] = lm_cmul(w8, w1);
  p[10] = lm_cmul(w8, w2);
  p[11] = lm_cmul(w8, p[3]);
  p[12] = lm_cmul(w8, w4);
  p[13] = lm_cmul(w8, p[5]);
  p[14] = lm_cmul(w8, p[6]);
  p[15] = lm_cmul(w8, p[7]);
}

// pass 1: load frame A (real) and frame B (imag), window, radix-16 over n1, twiddle.
// seg: this segment's L samples. startA/startB: index of sample n=0 of each frame
// in un-padded coordinates (may be negative / beyond L -> reflect).
// tw1: [16][128] table W_2048^(m*k1) (L1-resident; rebuilding the powers in registers was measured slower:
// +130 FP instructions and +16 registers per thread cost more than the 15 loads they save)
// valid <= L: number of samples of this segment that exist (waveform tail); samples in [valid, L) read as 0
// (= slice_padded_array zero padding fused into the load), reflection still happens at the segment length L.
YMT3_HD float lm_ld(const float* __restrict__ seg, int i, int L, int valid) {
  i = lm_reflect(i, L);
  return i < valid ? seg[i] : 0.f;
}
YMT3_HD void lm_pass1(int tid, const float* __restrict__ seg, int L, int valid, int startA, int startB,
                      bool hasB, const float (&w)[16], const float2* __restrict__ tw1, float2* __restrict__ bufA) {
  float2 v[16];
  const bool interior = (startA >= 0) && (startB + LM_NFFT <= valid);
  if (startB - startA == 128) {
    // hop == 128: frame B sample (n1) == frame A sample (n1 + 1): 17 loads feed both frames
    float xs[17];
    if (interior) {
#pragma unroll
      for (int n1 = 0; n1 < 17; ++n1) xs[n1] = seg[startA + 128 * n1 + tid];
    } else {
#pragma unroll
      for (int n1 = 0; n1 < 17; ++n1) xs[n1] = lm_ld(seg, startA + 128 * n1 + tid, L, valid);
    }
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) v[n1] = make_float2(xs[n1] * w[n1], xs[n1 + 1] * w[n1]);
    if (!hasB) {   // odd frame count: the last pair has no second frame (uniform, rare)
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) v[n1].y = 0.f;
    }
  } else if (interior) {
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
      float xa = lm_ld(seg, startA + n, L, valid);
      float xb = hasB ? lm_ld(seg, startB + n, L, valid) : 0.f;
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

YMT3_HD int lm_magaddr(int bin) { return bin + (bin >> 4); }

// |X_A|^p, |X_B|^p of one (k, N-k) pair from z = Z[k], zc = Z[N-k]
YMT3_HD float2 lm_pair_mag(float2 z, float2 zc, int power_mode) {
  // X_A = ((a+c) + i(b-d))/2 ; X_B = ((b+d) - i(a-c))/2 with z=a+ib, zc=c+id
  float ar = 0.5f * (z.x + zc.x), ai = 0.5f * (z.y - zc.y);
  float br = 0.5f * (z.y + zc.y), bi = 0.5f * (z.x - zc.x);
  float pa = ar * ar + ai * ai, pb = br * br + bi * bi;
  if (power_mode == 1) {
    pa = sqrtf(pa);
    pb = sqrtf(pb);
  }
  return make_float2(pa, pb);
}

// pass 3 + magnitudes: every thread owns TWO (k1, k2) columns chosen so that Z[k] and Z[N-k] of all its
// bins are in its own registers after the radix-8:
//   tid   0..111 : (k1 = 1 + tid/16, k2 = tid%16)  with partner (16-k1, 15-k2)      [k3 <-> 7-k3]
//   tid 112..119 : (8, tid-112)                    with partner (8, 15-k2)
//   tid 120..126 : (0, tid-119)                    with partner (0, 16-k2)
//   tid 127      : (0, 0) [k3 <-> (8-k3)%8] and (0, 8) [k3 <-> 7-k3], both self-paired
// mags[lm_magaddr(bin)] = (|X_A[bin]|^p, |X_B[bin]|^p), bin = min(k, N-k) in [0, 1024].
YMT3_HD void lm_pass3_mag(int tid, const float2* __restrict__ bufB, float2* __restrict__ mags, int power_mode) {
  int k1a, k2a, k1b, k2b;
  if (tid < 112) { k1a = 1 + (tid >> 4); k2a = tid & 15; k1b = 16 - k1a; k2b = 15 - k2a; }
  else if (tid < 120) { k1a = 8; k2a = tid - 112; k1b = 8; k2b = 15 - k2a; }
  else if (tid < 127) { k1a = 0; k2a = tid - 119; k1b = 0; k2b = 16 - k2a; }
  else { k1a = 0; k2a = 0; k1b = 0; k2b = 8; }
  float2 va[8], vb[8];
#pragma unroll
  for (int n3 = 0; n3 < 8; ++n3) {
    va[n3] = bufB[n3 * LM_S3 + k1a * 17 + k2a];
    vb[n3] = bufB[n3 * LM_S3 + k1b * 17 + k2b];
  }
  lm_fft8(va);
  lm_fft8(vb);
  // register index r = 2a + b holds k3 = a + 4b  ->  k3 lives at r(k3) = 2*(k3 & 3) + (k3 >> 2)
  if (tid < 127) {
    const int ka0 = k1a + 16 * k2a;
#pragma unroll
    for (int k3 = 0; k3 < 8; ++k3) {
      const int ra = 2 * (k3 & 3) + (k3 >> 2);
      const int k3p = 7 - k3;
      const int rb = 2 * (k3p & 3) + (k3p >> 2);
      const int k = ka0 + 256 * k3;
      const int bin = k <= 1024 ? k : LM_NFFT - k;
      mags[lm_magaddr(bin)] = lm_pair_mag(va[ra], vb[rb], power_mode);
    }
  } else {
    // (0,0): k = 256*k3, partner k3' = (8-k3)%8 -> bins 0, 256, 512, 768, 1024
#pragma unroll
    for (int k3 = 0; k3 <= 4; ++k3) {
      const int k3p = (8 - k3) & 7;
      const int ra = 2 * (k3 & 3) + (k3 >> 2), rb = 2 * (k3p & 3) + (k3p >> 2);
      mags[lm_magaddr(256 * k3)] = lm_pair_mag(va[ra], va[rb], power_mode);
    }
    // (0,8): k = 128 + 256*k3, partner 7-k3 -> bins 128, 384, 640, 896
#pragma unroll
    for (int k3 = 0; k3 < 4; ++k3) {
      const int k3p = 7 - k3;
      const int ra = 2 * (k3 & 3) + (k3 >> 2), rb = 2 * (k3p & 3) + (k3p >> 2);
      mags[lm_magaddr(128 + 256 * k3)] = lm_pair_mag(vb[ra], vb[rb], power_mode);
    }
  }
}

// mel projection + log for both frames. Filter m: sum_j mel_w[off+j] * mag[first+j] (banded filterbank).
YMT3_HD void lm_mel_log(int tid, const LmTables& tb, int n_mels, float eps,
                        const float2* __restrict__ mags, float* __restrict__ outA, float* __restrict__ outB) {
  for (int m = tid; m < n_mels; m += LM_THREADS) {
    const int2 meta = tb.mel_meta[m];
    const int first = meta.x & 0xffff, cnt = meta.x >> 16, o0 = meta.y;
    float sa = 0.f, sb = 0.f;
    for (int j = 0; j < cnt; ++j) {
      const float wgt = tb.mel_w[o0 + j];
      const float2 mg = mags[lm_magaddr(first + j)];
      sa = fmaf(wgt, mg.x, sa);
      sb = fmaf(wgt, mg.y, sb);
    }
    outA[m] = logf(fmaxf(sa, eps));
    if (outB) outB[m] = logf(fmaxf(sb, eps));
  }
}

// linear-frequency ("spec" codec) log output: bins [bin0, bin0 + n_out)
YMT3_HD void lm_spec_log(int tid, int bin0, int n_out, float eps, const float2* __restrict__ mags,
                         float* __restrict__ outA, float* __restrict__ outB) {
  for (int f = tid; f < n_out; f += LM_THREADS) {
    const float2 mg = mags[lm_magaddr(bin0 + f)];
    outA[f] = logf(fmaxf(mg.x, eps));
    if (outB) outB[f] = logf(fmaxf(mg.y, eps));
  }
}
