"""Pin the numpy frontend oracle (oracle/logmel.py) against the installed reference
dependency (torchaudio, live) and the committed golden vectors. CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import logmel as O
from tests.util import GOLDEN_DIR, rel_err, synth_noise, synth_multitrack

TOL = 1e-4   # north-star: log-mel within 1e-4 relative (fp32)


def _g(name):
    return np.load(os.path.join(GOLDEN_DIR, name))


def test_linspace_bit_exact_vs_torch():
    for a, b, n in [(0, 8000, 1025), (77.75, 2840.02, 514), (0.3, 9.1, 7), (1.0, 1.0, 1)]:
        assert np.array_equal(O.torch_linspace_f32(a, b, n), torch.linspace(a, b, n).numpy())


def test_filterbank_vs_golden_and_torchaudio():
    g = _g("melscale_fbanks_512.npz")
    fb_g = np.zeros((1025, 512), np.float32)
    fb_g[g["rows"], g["cols"]] = g["vals"]
    fb = O.melscale_fbanks(1025, 50.0, 8000.0, 512, 16000)
    # pow() differs by <= 1 ulp between numpy and ATen -> weights agree to ~1e-4 abs at the narrow low filters
    assert np.abs(fb - fb_g).max() < 2e-4
    assert ((fb != 0) == (fb_g != 0)).mean() > 0.9999
    torchaudio = pytest.importorskip("torchaudio")
    fb_ta = torchaudio.functional.melscale_fbanks(1025, 50.0, 8000.0, 512, 16000).numpy()
    assert np.array_equal(fb_ta, fb_g)


def test_product_filterbank_bit_identical_to_torchaudio_golden():
    from yourmt3_b200.spectrogram import mel_filterbank
    g = _g("melscale_fbanks_512.npz")
    fb_g = np.zeros((1025, 512), np.float32)
    fb_g[g["rows"], g["cols"]] = g["vals"]
    assert np.array_equal(mel_filterbank(1025, 50.0, 8000.0, 512, 16000).numpy(), fb_g)


def test_window():
    assert np.abs(O.hann_window_periodic(2048) - torch.hann_window(2048).numpy()).max() < 5e-7


@pytest.mark.parametrize("name,kw", [
    ("logmel_melspec_hop128.npz", dict(kind="mel", hop_length=128, power=1.0)),
    ("logmel_melspec_power2.npz", dict(kind="mel", hop_length=128, power=2.0)),
    ("logmel_spec_hop300.npz", dict(kind="spec", hop_length=300, power=1.0)),
])
def test_oracle_vs_golden(name, kw):
    g = _g(name)
    kind = kw.pop("kind")
    if kind == "mel":
        g_fb = _g("melscale_fbanks_512.npz")
        fb = np.zeros((1025, 512), np.float32)
        fb[g_fb["rows"], g_fb["cols"]] = g_fb["vals"]
        got = O.log_melspectrogram(g["audio"], fb=fb, window=torch.hann_window(2048).numpy(), **kw)
    else:
        got = O.log_spectrogram(g["audio"], window=torch.hann_window(2048).numpy(), **kw)
    assert got.shape == g["out"].shape
    assert rel_err(got, g["out"]) < TOL


def test_oracle_vs_torchaudio_live_full_segment():
    torchaudio = pytest.importorskip("torchaudio")
    x = np.concatenate([synth_noise(1), synth_multitrack(1)], 0)
    ms = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048, hop_length=128, f_min=50.0,
                                              f_max=8000.0, n_mels=512, power=1.0)
    ref = torch.log(torch.clamp(ms(torch.from_numpy(x)), min=1e-5)).transpose(1, 2).numpy()
    got = O.log_melspectrogram(x, fb=ms.mel_scale.fb.numpy(), window=ms.spectrogram.window.numpy())
    assert got.shape == (2, 256, 512)
    # noise segment: every bin well above the fp32 FFT noise floor -> strict bar
    assert rel_err(got[0], ref[0]) < TOL
    # tonal segment: bins ~100 dB below the peaks carry fp32 round-off of the *reference* FFT
    # itself (pocketfft fp32 vs the oracle's float64) -> compare in the linear domain against
    # the frame peak
    lin = np.abs(np.exp(got[1]) - np.exp(ref[1])) / np.exp(ref[1]).max(axis=-1, keepdims=True)
    assert lin.max() < 1e-5


def test_empty_and_short_inputs():
    assert O.log_melspectrogram(np.zeros((0, 4096), np.float32)).shape == (0, 33, 512)
    with pytest.raises(ValueError):
        O.log_melspectrogram(np.zeros((1, 1024), np.float32))
    # silence clamps to log(eps)
    y = O.log_melspectrogram(np.zeros((1, 2049), np.float32))
    assert np.allclose(y, np.log(np.float32(1e-5)))


def test_reference_fp32_itself_misses_1e4_on_quiet_linear_bins():
    """Why the `spec` codec's 1e-4 bar is applied to bins within 40 dB of the frame peak (tests/util.py
    assert_frontend_close, strict_everywhere=False): the REFERENCE'S OWN fp32 path (torchaudio Spectrogram + log on
    CPU) differs from the float64 oracle by more than 1e-4 on isolated bins 40+ dB under the peak of harmonic audio -
    fp32 FFT round-off amplified by the log - while it holds 1e-5 on every bin inside the 40 dB window.  Any fp32
    implementation shares that floor; the GPU tests therefore bound the quiet bins in the LINEAR domain (2e-6 of the
    frame peak) and apply the 1e-4 log-domain bar where the reference itself meets it."""
    import torch
    import torchaudio
    from tests.util import synth_multitrack
    x = synth_multitrack(4, seed=2)
    sp = torchaudio.transforms.Spectrogram(n_fft=2048, hop_length=300, power=1.0)
    ta = torch.log(torch.clamp(sp(torch.from_numpy(x))[:, 1:1025], min=1e-5)).transpose(1, 2).numpy().astype(np.float64)
    ref = O.log_spectrogram(x, hop_length=300, window=torch.hann_window(2048).numpy(), bin0=1, n_bins=1024).astype(np.float64)
    d = np.abs(ta - ref) / np.maximum(np.abs(ref), 1.0)
    peak = np.exp(ref).max(axis=-1, keepdims=True)
    big = np.exp(ref) >= 1e-2 * peak
    lin = np.abs(np.exp(ta) - np.exp(ref)) / peak
    assert d[big].max() < 1e-5                     # inside the window the reference is 10x better than the bar
    assert d.max() > 1e-4                          # outside it the reference itself is beyond the bar ...
    assert d.max() < 1e-3 and lin.max() < 2e-6     # ... by a bounded amount: the linear-domain error is tiny
