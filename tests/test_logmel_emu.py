"""CPU tier: run the kernel's own per-thread pass functions (logmel_core.cuh) through
the host emulation harness and compare with the oracle + golden vectors. This checks
the FFT factorisation, shared-memory index maps, frame pairing, reflect padding and the
banded mel projection without a GPU; the CUDA launch itself is covered by -m gpu tests."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import logmel as O
from tests.util import GOLDEN_DIR, assert_frontend_close, rel_err, synth_noise
from yourmt3_b200 import spectrogram as S

TOL = 1e-4


def emu_run(emu, layer, audio):
    audio = np.ascontiguousarray(audio, np.float32)
    B, L = audio.shape
    out = np.zeros((B, layer.num_frames(L), layer.num_features), np.float32)
    win = layer._window().numpy().copy()
    fb = None if layer._fb() is None else layer._fb().numpy().copy()
    cfg = layer._cfg()
    emu.lm_emu_run.restype = C.c_int
    rc = emu.lm_emu_run(C.byref(cfg), win.ctypes.data_as(C.c_void_p),
                        None if fb is None else fb.ctypes.data_as(C.c_void_p),
                        audio.ctypes.data_as(C.c_void_p), B, L, out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def oracle_run(layer, audio):
    win = layer._window().numpy()
    if layer.codec == "melspec":
        return O.log_melspectrogram(audio, hop_length=layer.hop_length, power=layer.power, eps=layer.eps,
                                    window=win, fb=layer._fb().numpy())
    return O.log_spectrogram(audio, hop_length=layer.hop_length, power=layer.power, eps=layer.eps, window=win,
                             bin0=layer.bin0, n_bins=layer.num_features)


@pytest.mark.parametrize("L", [1025, 2047, 2048, 4100, 32767])
@pytest.mark.parametrize("codec,hop", [("melspec", 128), ("spec", 300), ("melspec", 300), ("spec", 77)])
def test_emu_vs_oracle(emu_lib, L, codec, hop):
    layer = S.Melspectrogram(hop_length=hop) if codec == "melspec" else S.Spectrogram(hop_length=hop)
    x = synth_noise(2, L, seed=L + hop)
    got, ref = emu_run(emu_lib, layer, x), oracle_run(layer, x)
    assert_frontend_close(got, ref, TOL, strict_everywhere=(codec == "melspec"))


def test_emu_power2(emu_lib):
    layer = S.Melspectrogram(power=2.0)
    x = synth_noise(1, 3000, seed=5)
    assert rel_err(emu_run(emu_lib, layer, x), oracle_run(layer, x)) < TOL


@pytest.mark.parametrize("name,layer", [
    ("logmel_melspec_hop128.npz", lambda: S.Melspectrogram()),
    ("logmel_melspec_power2.npz", lambda: S.Melspectrogram(power=2.0)),
    ("logmel_spec_hop300.npz", lambda: S.Spectrogram()),
])
def test_emu_vs_golden(emu_lib, name, layer):
    g = np.load(os.path.join(GOLDEN_DIR, name))
    got = emu_run(emu_lib, layer(), g["audio"])
    assert got.shape == g["out"].shape
    assert rel_err(got, g["out"]) < TOL


def test_emu_impulse_and_dc(emu_lib):
    """Known answers: DC input -> only the lowest bins; silence -> log(eps)."""
    layer = S.Spectrogram(hop_length=128, drop_dc=False)
    x = np.zeros((2, 4096), np.float32)
    x[1] = 0.5
    y = emu_run(emu_lib, layer, x)
    assert np.allclose(y[0], np.log(np.float32(1e-5)))
    # hann window sums to n_fft/2 -> |X[0]| = 0.5 * 1024
    assert np.allclose(y[1, :, 0], np.log(512.0), atol=1e-5)
    assert np.allclose(y[1, :, 1], np.log(256.0), atol=1e-5)          # hann first side lobe = half
    assert np.all(y[1, :, 3:] < np.log(1e-3))


def test_emu_waveform_mode_equals_slice_padded_array(emu_lib):
    """Segmentation fused into the loads == slice_padded_array + per-segment frontend."""
    from yourmt3_b200.audio_utils import slice_padded_array
    layer = S.Melspectrogram()
    L = 4100
    wave = synth_noise(1, 3 * L - 1234, seed=3)[0]
    segs = np.ascontiguousarray(slice_padded_array(wave, L, L)[:, 0, :])
    ref = emu_run(emu_lib, layer, segs)
    out = np.zeros_like(ref)
    win, fb, cfg = layer._window().numpy().copy(), layer._fb().numpy().copy(), layer._cfg()
    emu_lib.lm_emu_run_wave.restype = C.c_int
    rc = emu_lib.lm_emu_run_wave(C.byref(cfg), win.ctypes.data_as(C.c_void_p), fb.ctypes.data_as(C.c_void_p),
                                 np.ascontiguousarray(wave).ctypes.data_as(C.c_void_p), C.c_longlong(wave.size), 3, L,
                                 out.ctypes.data_as(C.c_void_p))
    assert rc == 0 and segs.shape[0] == 3
    assert np.array_equal(out, ref)
