"""Shared helpers for tests: synthetic audio and error metrics."""
import numpy as np
import torch

GOLDEN_DIR = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden")


def synth_noise(B, L=32767, seed=1234, scale=0.1):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, L, generator=g) * scale).numpy().astype(np.float32)


def synth_multitrack(B, L=32767, sr=16000, seed=1234):
    """4-8 harmonic-plus-decay note streams, Poisson onsets ~8/s, -40 dB noise."""
    rng = np.random.default_rng(seed)
    t = np.arange(L) / sr
    out = np.zeros((B, L), np.float64)
    for b in range(B):
        for _ in range(rng.integers(4, 9)):
            n_on = rng.poisson(8 * L / sr / 4) + 1
            for on in rng.uniform(0, L / sr, n_on):
                f0 = 440.0 * 2 ** ((rng.integers(36, 97) - 69) / 12)
                env = np.where(t >= on, np.exp(-(t - on) * rng.uniform(2, 8)), 0.0)
                for h in range(1, 5):
                    if f0 * h < sr / 2:
                        out[b] += env * np.sin(2 * np.pi * f0 * h * (t - on)) / h * 0.1
        out[b] += rng.standard_normal(L) * 0.01 * np.abs(out[b]).max()
        out[b] /= max(1e-9, np.abs(out[b]).max())
    return (out * 0.8).astype(np.float32)


def rel_err(a, b):
    """max |a-b| / max(|b|, 1): relative where |b|>1, absolute near zero (log domain)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))) if a.size else 0.0


def assert_frontend_close(got, ref, tol=1e-4, strict_everywhere=True):
    """North-star bar: within `tol` relative in the log domain.

    mel outputs (strict_everywhere=True): every bin.  Linear-frequency ("spec") outputs have
    isolated bins ~60 dB under the frame peak where the fp32 FFT round-off of ANY fp32
    implementation (the reference's included) is amplified by the log; there the bar is applied
    to bins within 40 dB of the frame peak, and every bin must be within 2e-6 of the frame
    peak in the linear domain (fp32 FFT accuracy)."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape
    if got.size == 0:
        return
    peak = np.exp(ref).max(axis=-1, keepdims=True)
    lin = np.abs(np.exp(got) - np.exp(ref)) / peak
    assert lin.max() < 2e-6, f"linear-domain error {lin.max():.3e} of frame peak"
    d = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    if strict_everywhere:
        assert d.max() < tol, f"log-domain rel err {d.max():.3e}"
    else:
        big = np.exp(ref) >= 1e-2 * peak
        assert big.mean() > 0.99
        assert d[big].max() < tol, f"log-domain rel err {d[big].max():.3e} on bins within 40 dB of peak"
        assert d.max() < 1e-2


def fill_by_name_(named_tensors, std=0.08, norm_jitter=0.1):
    """Deterministic weights that depend only on the parameter NAME and shape (seed = crc32(name)), so the HF-side
    generator of tests/golden/t5_hf_*.npz (tools/make_golden_t5.py) and the native-side tests build identical
    weights without storing them.  >= 2-D: N(0, std); 1-D (norm scales): 1 + N(0, norm_jitter)."""
    import zlib
    with torch.no_grad():
        for name, p in named_tensors:
            g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
            if p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g) * std)
            else:
                p.copy_(1.0 + norm_jitter * torch.randn(p.shape, generator=g))
