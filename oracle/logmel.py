"""numpy restatement of the log-(mel)spectrogram frontend (oracle; test infra only).

Follows, line by line, the installed dependency code the upstream
``amt/src/model/spectrogram.py`` wrappers call (SP = site-packages):

* SP/torchaudio/functional/functional.py:119-144  spectrogram(): reshape, torch.stft
  (center=True, pad_mode="reflect", onesided), abs() / abs().pow(power)
* SP/torch/functional.py:675-691                   stft(): reflect pad n_fft//2, frame, window, rFFT
* SP/torchaudio/functional/functional.py:425-587   _hz_to_mel/_mel_to_hz/_create_triangular_filterbank/
                                                   melscale_fbanks (HTK scale, norm=None)
* SP/torchaudio/transforms/_transforms.py:417      MelScale.forward: matmul(spec^T, fb)^T
* upstream model/spectrogram.py [RECALL]           log(clamp(x, min=eps)), output (B, T, F)

The FFT is evaluated in float64 and rounded to float32, i.e. this oracle is at
least as accurate as the fp32 reference path; tests pin it to torchaudio's own
CPU output within 1e-4 (tests/test_oracle_logmel.py).
"""
from __future__ import annotations

import math

import numpy as np


def torch_linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """torch.linspace(start, end, steps) in float32.

    ATen fills symmetrically from both ends with a fused multiply-add
    (start + step*i for i < steps/2, end - step*(steps-1-i) after); the fma is
    emulated by evaluating in float64 and rounding once (bit-exact vs torch in tests).
    """
    start32, end32 = np.float32(start), np.float32(end)
    if steps == 1:
        return np.array([start32], dtype=np.float32)
    step = np.float32((end32 - start32) / np.float32(steps - 1))
    i = np.arange(steps)
    half = steps // 2
    lo = (np.float64(start32) + np.float64(step) * i).astype(np.float32)
    hi = (np.float64(end32) - np.float64(step) * (steps - 1 - i)).astype(np.float32)
    return np.where(i < half, lo, hi).astype(np.float32)


def hann_window_periodic(n: int) -> np.ndarray:
    """torch.hann_window(n) (periodic=True) in float32."""
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * math.pi * k / n)).astype(np.float32)


def hz_to_mel_htk(freq: float) -> float:
    # functional.py:439-440
    return 2595.0 * math.log10(1.0 + (freq / 700.0))


def melscale_fbanks(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> np.ndarray:
    """functional.py:518-587 with mel_scale='htk', norm=None. Returns (n_freqs, n_mels) float32."""
    all_freqs = torch_linspace_f32(0, sample_rate // 2, n_freqs)
    m_pts = torch_linspace_f32(hz_to_mel_htk(f_min), hz_to_mel_htk(f_max), n_mels + 2)
    # functional.py:477  700 * (10 ** (mels / 2595) - 1)  in float32
    f_pts = (np.float32(700.0) * (np.power(np.float32(10.0), m_pts / np.float32(2595.0)) - np.float32(1.0))).astype(np.float32)
    # functional.py:505-513
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (np.float32(-1.0) * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(np.float32(0.0), np.minimum(down, up)).astype(np.float32)


def stft_power(audio: np.ndarray, n_fft: int, hop: int, window: np.ndarray, power: float) -> np.ndarray:
    """(B, L) float32 -> (B, n_fft//2+1, T) float32, |STFT|**power, center/reflect."""
    audio = np.asarray(audio, dtype=np.float32)
    B, L = audio.shape
    pad = n_fft // 2
    if not pad < L:
        raise ValueError("reflect padding needs n_fft//2 < L (torch/functional.py:675-680)")
    x = np.pad(audio.astype(np.float64), ((0, 0), (pad, pad)), mode="reflect")
    T = 1 + L // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(T)[:, None]
    frames = x[:, idx] * window.astype(np.float64)[None, None, :]  # (B, T, n_fft)
    spec = np.fft.rfft(frames, n=n_fft, axis=-1)                   # float64
    re = spec.real.astype(np.float32)
    im = spec.imag.astype(np.float32)
    mag2 = re * re + im * im
    if power == 1.0:
        out = np.sqrt(mag2)
    elif power == 2.0:
        out = mag2
    else:
        out = np.power(np.sqrt(mag2), np.float32(power))
    return np.transpose(out, (0, 2, 1)).astype(np.float32)


def log_melspectrogram(audio: np.ndarray, *, sample_rate=16000, n_fft=2048, hop_length=128, f_min=50.0,
                       f_max=8000.0, n_mels=512, power=1.0, eps=1e-5, window=None, fb=None) -> np.ndarray:
    """(B, L) or (B, 1, L) -> (B, T, n_mels): log(clamp(mel(|STFT|^p), eps))."""
    a = np.asarray(audio, dtype=np.float32)
    if a.ndim == 3:
        a = a[:, 0, :]
    window = hann_window_periodic(n_fft) if window is None else window
    fb = melscale_fbanks(n_fft // 2 + 1, f_min, f_max, n_mels, sample_rate) if fb is None else fb
    spec = stft_power(a, n_fft, hop_length, window, power)              # (B, F, T)
    mel = np.einsum("bft,fm->btm", spec.astype(np.float64), fb.astype(np.float64)).astype(np.float32)
    return np.log(np.maximum(mel, np.float32(eps))).astype(np.float32)


def log_spectrogram(audio: np.ndarray, *, n_fft=2048, hop_length=300, power=1.0, eps=1e-5, bin0=1, n_bins=1024,
                    window=None) -> np.ndarray:
    """(B, L) or (B, 1, L) -> (B, T, n_bins): log(clamp(|STFT|^p, eps))[bin0 : bin0+n_bins]."""
    a = np.asarray(audio, dtype=np.float32)
    if a.ndim == 3:
        a = a[:, 0, :]
    window = hann_window_periodic(n_fft) if window is None else window
    spec = stft_power(a, n_fft, hop_length, window, power)              # (B, F, T)
    spec = spec[:, bin0:bin0 + n_bins, :]
    return np.log(np.maximum(np.transpose(spec, (0, 2, 1)), np.float32(eps))).astype(np.float32)
