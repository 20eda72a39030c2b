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
#define LM_MEL_NV 3               // float4 weight vectors per fixed-width mel filter record (<= 12 bins per filter)
#define LM_MAG_ELEMS 1092         // interleaved (magA, magB) float2 per bin, padded: addr(bin) = bin + (bin >> 4)

struct LmTables {
  const float*  window;     // [2048] analysis window (win_length == n_fft)
  const float2* tw1;        // [16][128]  W_2048^(m*k1)
  const float2* tw2;        // [8][16]    W_128^(n3*k2)
  const int*    mel_first;  // [n_mels]   first nonzero bin of filter m
  const int*    mel_off;    // [n_mels+1] prefix offsets into mel_w
  const int2*   mel_meta;   // [n_mels]   (first | count << 16, offset)
  const float*  mel_w;      // packed nonzero weights, ascending bin order
  // fixed-width filter records for banks whose filters span <= 4 * LM_MEL_NV bins (the 512-band HTK bank: <= 11):
  // rec_w[v][m] = weights 4v..4v+3 of filter m (zero padded), so the weight loads do not depend on the meta load and
  // a thread issues the loads of two filters at once.  rec_ok == 0: wider bank, generic per-filter loop.
  const float4* rec_w;      // [LM_MEL_NV][n_mels]
  int           rec_ok;
};

// Complex arithmetic on (re, im) register pairs.  On sm_100a every helper is ONE packed fp32 instruction
// (FADD2 / FMUL2 / FFMA2: add.f32x2 / mul.f32x2 / fma.f32x2), the +-i rotations and the scalar broadcasts ride on the
// instructions' operand modifiers (.LO_HI swap, per-half negate, .F32 broadcast), so a complex add is one issue slot and
// a complex multiply two - the r01 kernel was issue-bound with 55 % of its slots in scalar FADD / FMUL / FFMA
// (profiles/r01_logmel_default_workload_ncu_full.txt).  Same IEEE round-to-nearest operations as the scalar forms, so
// the host emulation (tests/host_emu) mirrors them with fmaf.
YMT3_HD float2 lm_add(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  return __fadd2_rn(a, b);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
YMT3_HD float2 lm_sub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  return __fadd2_rn(a, make_float2(-b.x, -b.y));
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// a - i*b = (a.x + b.y, a.y - b.x)
YMT3_HD float2 lm_add_mi(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  return __fadd2_rn(a, make_float2(b.y, -b.x));
#else
  return make_float2(a.x + b.y, a.y - b.x);
#endif
}
// a + i*b = (a.x - b.y, a.y + b.x)
YMT3_HD float2 lm_add_pi(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  return __fadd2_rn(a, make_float2(-b.y, b.x));
#else
  return make_float2(a.x - b.y, a.y + b.x);
#endif
}
// a * w = (a.x w.x - a.y w.y, a.x w.y + a.y w.x) = w.x * (a.x, a.y) + w.y * (-a.y, a.x): in THIS form ptxas folds the
// swap + half negation of `a` and both scalar broadcasts into operand modifiers (FMUL2 a, w.x.F32; FFMA2
// -a.LO_HI.NP, w.y.F32, t): exactly two instructions; the forms that swizzle the twiddle instead cost a scalar FADD
// and a MOV per product (checked in SASS).
YMT3_HD float2 lm_cmul(float2 a, float2 w) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  return __ffma2_rn(make_float2(-a.y, a.x), make_float2(w.y, w.y), __fmul2_rn(a, make_float2(w.x, w.x)));
#else
  return make_float2(fmaf(-a.y, w.y, a.x * w.x), fmaf(a.x, w.y, a.y * w.x));
#endif
}
// (x * s, y * s)
YMT3_HD float2 lm_scale2(float x, float y, float s) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  return __fmul2_rn(make_float2(x, y), make_float2(s, s));
#else
  return make_float2(x * s, y * s);
#endif
}

// forward DFT-4, natural order in and out
YMT3_HD void lm_fft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = lm_add(a0, a2), t1 = lm_sub(a0, a2);
  float2 t2 = lm_add(a1, a3), t3 = lm_sub(a1, a3);
  a0 = lm_add(t0, t2);
  a2 = lm_sub(t0, t2);
  a1 = lm_add_mi(t1, t3);  // t1 - i*t3
  a3 = lm_add_pi(t1, t3);  // t1 + i*t3
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
  v[5] = make_float2(v[5].y, -v[5].x);   // * (-i): folds into the operand modifiers of the add / sub below
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

// Staging: the hop + 2048 samples [startA, startA + hop + 2048) that frames A and B of a pair cover are staged ONCE
// in shared memory.  Interior pairs (no padding involved): one elected thread issues a 1-D bulk async copy
// (cp.async.bulk, TMA engine) for the NEXT pair right after the pass-1 barrier, so the global-memory latency hides
// under passes 2-3 of the current pair (logmel.cu).  Edge pairs (reflect padding at the segment ends, zero-padded
// waveform tail, misaligned buffers): the CTA fills the buffer itself with this function - also the path the host
// emulation takes for every pair.
// seg: this segment's L samples. startA: index of sample n = 0 of frame A in un-padded coordinates (may be negative /
// beyond L -> reflect, torch/functional.py:675-680).  valid <= L: number of samples of this segment that exist
// (waveform tail); samples in [valid, L) read as 0 (= slice_padded_array zero padding fused into the load),
// reflection still happens at the segment length L.
YMT3_HD float lm_ld(const float* __restrict__ seg, int i, int L, int valid) {
  i = lm_reflect(i, L);
  return (i >= 0 && i < valid) ? seg[i] : 0.f;
}
YMT3_HD void lm_stage_fill(int tid, const float* __restrict__ seg, int L, int valid, int startA, int count,
                           float* __restrict__ stage) {
  for (int i = tid; i < count; i += LM_THREADS) stage[i] = lm_ld(seg, startA + i, L, valid);
}

// pass 1: frame A (real) and frame B (imag) from the staged samples (fa[n] = frame A sample n, frame B = fa + hop),
// window, radix-16 over n1, twiddle W_2048^(tid*k1) from the thread's REGISTERS (tw[k1], loop-invariant: the r01 / v3
// kernels re-read 15 table entries per pair through an L1 that the 200+ KB of shared memory leaves at 18 KB).
YMT3_HD void lm_pass1(int tid, const float* __restrict__ fa, int hop, bool hasB, const float (&w)[16],
                      const float2 (&tw)[16], float2* __restrict__ bufA) {
  float2 v[16];
  if (hop == 128) {
    // frame B sample (n1) == frame A sample (n1 + 1): 17 loads feed both frames
    float xs[17];
#pragma unroll
    for (int n1 = 0; n1 < 17; ++n1) xs[n1] = fa[128 * n1 + tid];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) v[n1] = lm_scale2(xs[n1], xs[n1 + 1], w[n1]);
  } else {
    const float* fb = fa + hop;
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) v[n1] = lm_scale2(fa[128 * n1 + tid], fb[128 * n1 + tid], w[n1]);
  }
  if (!hasB) {   // odd frame count: the last pair has no second frame (uniform, rare)
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) v[n1].y = 0.f;
  }
  lm_fft16(v);
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int k1 = (r >> 2) + 4 * (r & 3);
    float2 o = v[r];
    if (k1 != 0) o = lm_cmul(o, tw[k1]);
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

// One (k, N-k) pair, z = Z[k], zc = Z[N-k]:  X_A = ((a+c) + i(b-d))/2, X_B = ((b+d) - i(a-c))/2 with z = a+ib,
// zc = c+id, so with S = z + zc, D = z - zc:  4|X_A|^2 = S.x^2 + D.y^2,  4|X_B|^2 = S.y^2 + D.x^2  - four packed
// instructions for both frames.  The factor 4 (and the square root of power = 1) is folded into the output stage:
// take_sqrt = false -> (4|X_A|^2, 4|X_B|^2);  true -> (2|X_A|, 2|X_B|)  (sqrt.approx, 1 ulp-class).
YMT3_HD float lm_fast_sqrt(float x) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return sqrtf(x);
#endif
}
YMT3_HD float lm_fast_log2(float x) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  float y;             // lg2.approx.ftz: relative error 2^-22; the argument is clamped above any denormal
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return log2f(x);
#endif
}
YMT3_HD float2 lm_pair_mag(float2 z, float2 zc, bool take_sqrt) {
  const float2 S = lm_add(z, zc), D = lm_sub(z, zc);
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  const float2 dq = __fmul2_rn(D, D);
  float2 q = __ffma2_rn(S, S, make_float2(dq.y, dq.x));
#else
  float2 q = make_float2(fmaf(S.x, S.x, D.y * D.y), fmaf(S.y, S.y, D.x * D.x));
#endif
  if (take_sqrt) q = make_float2(lm_fast_sqrt(q.x), lm_fast_sqrt(q.y));
  return q;
}

// pass 3 + magnitudes: every thread owns TWO (k1, k2) columns chosen so that Z[k] and Z[N-k] of all its
// bins are in its own registers after the radix-8:
//   tid   0..111 : (k1 = 1 + tid/16, k2 = tid%16)  with partner (16-k1, 15-k2)      [k3 <-> 7-k3]
//   tid 112..119 : (8, tid-112)                    with partner (8, 15-k2)
//   tid 120..126 : (0, tid-119)                    with partner (0, 16-k2)
//   tid 127      : (0, 0) [k3 <-> (8-k3)%8] and (0, 8) [k3 <-> 7-k3], both self-paired
// mags[lm_magaddr(bin)] = lm_pair_mag of the two frames (see there for the folded scale), bin = min(k, N-k) in [0, 1024].
// For tid < 127 the first column has ka0 = k1a + 16*k2a <= 247, so k = ka0 + 256*k3 is a bin itself for k3 < 4 and
// mirrors to bin 2048 - k for k3 >= 4: both shared-memory addresses are a per-thread base plus a compile-time
// multiple of 272 (= 256 + 256/16).
YMT3_HD void lm_pass3_mag(int tid, const float2* __restrict__ bufB, float2* __restrict__ mags, bool take_sqrt) {
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
    float2* lo = mags + lm_magaddr(ka0);              // bins ka0 + 256*k3,           k3 = 0..3
    float2* hi = mags + lm_magaddr(256 - ka0);        // bins (256 - ka0) + 256*(7-k3), k3 = 4..7
#pragma unroll
    for (int k3 = 0; k3 < 8; ++k3) {
      const int ra = 2 * (k3 & 3) + (k3 >> 2);
      const int k3p = 7 - k3;
      const int rb = 2 * (k3p & 3) + (k3p >> 2);
      float2* dst = k3 < 4 ? lo + 272 * k3 : hi + 272 * (7 - k3);
      *dst = lm_pair_mag(va[ra], vb[rb], take_sqrt);
    }
  } else {
    // (0,0): k = 256*k3, partner k3' = (8-k3)%8 -> bins 0, 256, 512, 768, 1024
#pragma unroll
    for (int k3 = 0; k3 <= 4; ++k3) {
      const int k3p = (8 - k3) & 7;
      const int ra = 2 * (k3 & 3) + (k3 >> 2), rb = 2 * (k3p & 3) + (k3p >> 2);
      mags[lm_magaddr(256 * k3)] = lm_pair_mag(va[ra], va[rb], take_sqrt);
    }
    // (0,8): k = 128 + 256*k3, partner 7-k3 -> bins 128, 384, 640, 896
#pragma unroll
    for (int k3 = 0; k3 < 4; ++k3) {
      const int k3p = 7 - k3;
      const int ra = 2 * (k3 & 3) + (k3 >> 2), rb = 2 * (k3p & 3) + (k3p >> 2);
      mags[lm_magaddr(128 + 256 * k3)] = lm_pair_mag(vb[ra], vb[rb], take_sqrt);
    }
  }
}

// Output stage.  The stored values carry a folded scale (lm_pair_mag): with m = the stored value and
//   |X|^p = mag_scale * m,   log(max(|X|^p [summed over a mel band], eps)) = ln2 * log2(max(sum, eps / mag_scale)) + ln(mag_scale)
// and for the linear-frequency codec with power 1 the square root is never taken at all:
//   log(max(|X|, eps)) = 0.5 * ln2 * log2(max(4|X|^2, 4 eps^2)) - ln2.
// LmOut carries (floor, c1, c0): out = c1 * log2(max(v, floor)) + c0   (lg2.approx + one FFMA per output).
struct LmOut { float floor, c1, c0; };
YMT3_HD LmOut lm_out_consts(int codec_spec, int power_mode, float eps) {
  const float ln2 = 0.69314718055994531f;
  LmOut o;
  if (codec_spec && power_mode == 1) { o.floor = 4.f * eps * eps; o.c1 = 0.5f * ln2; o.c0 = -ln2; }   // v = 4|X|^2
  else if (power_mode == 1) { o.floor = 2.f * eps; o.c1 = ln2; o.c0 = -ln2; }                         // v = sum w * 2|X|
  else { o.floor = 4.f * eps; o.c1 = ln2; o.c0 = -2.f * ln2; }                                        // v = (sum w *) 4|X|^2
  return o;
}
YMT3_HD float lm_out(float v, const LmOut& o) { return fmaf(o.c1, lm_fast_log2(fmaxf(v, o.floor)), o.c0); }

// Mel projection (banded filterbank: sum_j mel_w[off+j] * mag[first+j]) + log for both frames, a thread per filter.
// Fast path (rec_ok): fixed-width records, the loads of TWO filters (2 meta + 6 float4, mutually independent) are
// issued before either is consumed, the bins are walked in groups of four with one (near warp-uniform) test per group;
// the r01 loop had a dependent meta -> weight -> accumulate chain per bin and ~450 warp instructions per pair.
YMT3_HD float2 lm_mel_filter(int meta, const float4 (&wv)[LM_MEL_NV], const float2* __restrict__ mags) {
  const int first = meta & 0xffff, cnt = meta >> 16;
  const float2* pa = mags + lm_magaddr(first);
  const float2* pb = pa + 1;                // lm_magaddr inserts one pad slot after every 16 bins
  const int skip = 16 - (first & 15);       // bins j >= skip sit one slot further
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int v = 0; v < LM_MEL_NV; ++v) {
    if (4 * v < cnt) {
      const float wq[4] = {wv[v].x, wv[v].y, wv[v].z, wv[v].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = 4 * v + q;            // weights beyond cnt are zero; the slots read stay inside `mags`
        const float2 mg = (j < skip ? pa : pb)[j];
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
        acc = __ffma2_rn(make_float2(wq[q], wq[q]), mg, acc);
#else
        acc = make_float2(fmaf(wq[q], mg.x, acc.x), fmaf(wq[q], mg.y, acc.y));
#endif
      }
    }
  }
  return acc;
}
// Records of the first LM_MEL_PRE filters of a thread (m = tid + 128 i; 2 -> no register spills): loaded BEFORE the barrier that ends pass 3
// (lm_mel_prefetch), so the table latency hides under the barrier wait; consumed after it (lm_mel_log).
#define LM_MEL_PRE 2
struct LmMelRec {
  int meta[LM_MEL_PRE];
  float4 w[LM_MEL_PRE][LM_MEL_NV];
};
YMT3_HD void lm_mel_prefetch(int tid, const LmTables& tb, int n_mels, LmMelRec& rec) {
  if (!tb.rec_ok) return;
#pragma unroll
  for (int i = 0; i < LM_MEL_PRE; ++i) {
    const int m = tid + LM_THREADS * i;
    const bool on = m < n_mels;
    rec.meta[i] = on ? tb.mel_meta[m].x : 0;
#pragma unroll
    for (int v = 0; v < LM_MEL_NV; ++v) rec.w[i][v] = on ? tb.rec_w[v * n_mels + m] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
YMT3_HD void lm_mel_log(int tid, const LmTables& tb, int n_mels, const LmOut& oc, const LmMelRec& rec,
                        const float2* __restrict__ mags, float* __restrict__ outA, float* __restrict__ outB) {
  int m_next = tid;
  if (tb.rec_ok) {
#pragma unroll
    for (int i = 0; i < LM_MEL_PRE; ++i) {
      const int m = tid + LM_THREADS * i;
      if (m < n_mels) {
        const float2 a = lm_mel_filter(rec.meta[i], rec.w[i], mags);
        outA[m] = lm_out(a.x, oc);
        if (outB) outB[m] = lm_out(a.y, oc);
      }
    }
    // the remaining filters, two at a time (their 8 loads are mutually independent: one exposed latency per pair)
    for (int m0 = tid + LM_THREADS * LM_MEL_PRE; m0 < n_mels; m0 += 2 * LM_THREADS) {
      const int m1 = m0 + LM_THREADS;
      const bool two = m1 < n_mels;
      const int meta0 = tb.mel_meta[m0].x, meta1 = two ? tb.mel_meta[m1].x : 0;
      float4 w0[LM_MEL_NV], w1[LM_MEL_NV];
#pragma unroll
      for (int v = 0; v < LM_MEL_NV; ++v) {
        w0[v] = tb.rec_w[v * n_mels + m0];
        w1[v] = two ? tb.rec_w[v * n_mels + m1] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float2 a0 = lm_mel_filter(meta0, w0, mags);
      outA[m0] = lm_out(a0.x, oc);
      if (outB) outB[m0] = lm_out(a0.y, oc);
      if (two) {
        const float2 a1 = lm_mel_filter(meta1, w1, mags);
        outA[m1] = lm_out(a1.x, oc);
        if (outB) outB[m1] = lm_out(a1.y, oc);
      }
    }
    return;
  }
  for (int m = m_next; m < n_mels; m += LM_THREADS) {   // generic bank (filters wider than 4 * LM_MEL_NV bins)
    const int2 meta = tb.mel_meta[m];
    const int first = meta.x & 0xffff, cnt = meta.x >> 16, o0 = meta.y;
    float2 acc = make_float2(0.f, 0.f);
    for (int j = 0; j < cnt; ++j) {
      const float wgt = tb.mel_w[o0 + j];
      const float2 mg = mags[lm_magaddr(first + j)];
      acc = make_float2(fmaf(wgt, mg.x, acc.x), fmaf(wgt, mg.y, acc.y));
    }
    outA[m] = lm_out(acc.x, oc);
    if (outB) outB[m] = lm_out(acc.y, oc);
  }
}

// linear-frequency ("spec" codec) log output: bins [bin0, bin0 + n_out)
YMT3_HD void lm_spec_log(int tid, int bin0, int n_out, const LmOut& oc, const float2* __restrict__ mags,
                         float* __restrict__ outA, float* __restrict__ outB) {
#pragma unroll 8
  for (int f = tid; f < n_out; f += LM_THREADS) {
    const float2 mg = mags[lm_magaddr(bin0 + f)];
    outA[f] = lm_out(mg.x, oc);
    if (outB) outB[f] = lm_out(mg.y, oc);
  }
}
