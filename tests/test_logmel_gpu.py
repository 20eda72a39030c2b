"""GPU parity tests for the fused log-(mel)spectrogram kernel, through the C ABI.
Bar (north-star): within 1e-4 relative of the reference fp32 path."""
import os

import numpy as np
import pytest
import torch

from oracle import logmel as O
from tests.util import GOLDEN_DIR, assert_frontend_close, rel_err, synth_multitrack, synth_noise
from yourmt3_b200 import spectrogram as S

pytestmark = pytest.mark.gpu
TOL = 1e-4


def oracle_run(layer, audio):
    win = layer._window().cpu().numpy()
    if layer.codec == "melspec":
        return O.log_melspectrogram(audio, hop_length=layer.hop_length, power=layer.power, eps=layer.eps,
                                    window=win, fb=layer._fb().cpu().numpy())
    return O.log_spectrogram(audio, hop_length=layer.hop_length, power=layer.power, eps=layer.eps, window=win,
                             bin0=layer.bin0, n_bins=layer.num_features)


def run(layer, audio, dev):
    return layer(torch.from_numpy(np.ascontiguousarray(audio)).to(dev)).cpu().numpy()


@pytest.mark.parametrize("L", [1025, 2048, 4100, 32767])
@pytest.mark.parametrize("codec,hop", [("melspec", 128), ("spec", 300), ("melspec", 300), ("spec", 77)])
def test_cuda_vs_oracle(cuda_device, native_lib, L, codec, hop):
    layer = S.Melspectrogram(hop_length=hop) if codec == "melspec" else S.Spectrogram(hop_length=hop)
    x = synth_noise(3, L, seed=L + hop)
    got, ref = run(layer, x, cuda_device), oracle_run(layer, x)
    assert_frontend_close(got, ref, TOL, strict_everywhere=(codec == "melspec"))


@pytest.mark.parametrize("name,layer", [
    ("logmel_melspec_hop128.npz", lambda: S.Melspectrogram()),
    ("logmel_melspec_power2.npz", lambda: S.Melspectrogram(power=2.0)),
    ("logmel_spec_hop300.npz", lambda: S.Spectrogram()),
])
def test_cuda_vs_golden(cuda_device, native_lib, name, layer):
    g = np.load(os.path.join(GOLDEN_DIR, name))
    got = run(layer(), g["audio"], cuda_device)
    assert got.shape == g["out"].shape
    assert rel_err(got, g["out"]) < TOL


def test_cuda_vs_torchaudio_live(cuda_device, native_lib):
    """Against the installed reference dependency itself (CPU fp32 path), 3-D input."""
    torchaudio = pytest.importorskip("torchaudio")
    x = torch.from_numpy(synth_noise(4)).unsqueeze(1)
    ms = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048, hop_length=128, f_min=50.0,
                                              f_max=8000.0, n_mels=512, power=1.0)
    ref = torch.log(torch.clamp(ms(x), min=1e-5))[:, 0].transpose(1, 2).numpy()
    got = S.Melspectrogram()(x.to(cuda_device)).cpu().numpy()
    assert got.shape == (4, 256, 512)
    assert rel_err(got, ref) < TOL


def test_multitrack_audio(cuda_device, native_lib):
    layer = S.Melspectrogram()
    x = synth_multitrack(2)
    got, ref = run(layer, x, cuda_device), oracle_run(layer, x)
    lin = np.abs(np.exp(got) - np.exp(ref)) / np.exp(ref).max(axis=-1, keepdims=True)
    assert lin.max() < 1e-5
    assert np.median(np.abs(got - ref)) < 1e-5


def test_edge_cases(cuda_device, native_lib):
    layer = S.Melspectrogram()
    # empty batch
    y = layer(torch.zeros(0, 1, 32767, device=cuda_device))
    assert y.shape == (0, 256, 512)
    # silence -> log(eps) everywhere
    y = layer(torch.zeros(2, 32767, device=cuda_device))
    assert torch.allclose(y, torch.full_like(y, float(np.log(np.float32(1e-5)))))
    # too-short segment is rejected like torch.stft's reflect pad
    with pytest.raises(RuntimeError, match="exceed n_fft/2"):
        layer(torch.zeros(1, 1024, device=cuda_device))
    with pytest.raises(TypeError):
        layer(torch.zeros(1, 4096, device=cuda_device, dtype=torch.float64))
    # non-contiguous input view
    x = torch.from_numpy(synth_noise(2, 8192)).to(cuda_device)
    xv = x[:, ::2]
    assert torch.equal(layer(xv), layer(xv.contiguous()))


def test_full_size_properties(cuda_device, native_lib):
    """B=1024 (0.67 GB of output): size-independent properties.
    (i) batch independence: every segment equals the same segment run alone;
    (ii) homogeneity: log-mel(a*x) = log-mel(x) + log(a) away from the clamp."""
    layer = S.Melspectrogram()
    g = torch.Generator(device="cpu").manual_seed(7)
    base = torch.randn(8, 32767, generator=g) * 0.1
    x = base.repeat(128, 1).to(cuda_device)                      # 1024 segments
    y = layer(x)
    assert y.shape == (1024, 256, 512)
    y8 = layer(base.to(cuda_device))
    assert torch.equal(y.view(128, 8, 256, 512), y8.unsqueeze(0).expand(128, -1, -1, -1))
    ys = layer(base.to(cuda_device) * 4.0)
    assert torch.allclose(ys, y8 + float(np.log(4.0)), atol=2e-5)
    ref = oracle_run(layer, base[:2].numpy())
    assert rel_err(y8[:2].cpu().numpy(), ref) < TOL


def test_host_entry_point(cuda_device, native_lib):
    layer = S.Spectrogram()
    x = torch.from_numpy(synth_noise(5, 20000))
    got = layer.forward_host(x)
    assert torch.equal(got, layer(x.to(cuda_device)).cpu())


def test_state_dict_buffers_drive_the_kernel(cuda_device, native_lib):
    """Checkpoint buffers (window / fb) are what the kernel uses: change them, output changes."""
    layer = S.Melspectrogram().to(cuda_device)
    assert set(layer.state_dict()) == {"mel_stft.spectrogram.window", "mel_stft.mel_scale.fb"}
    x = torch.from_numpy(synth_noise(1, 4096)).to(cuda_device)
    y0 = layer(x)
    sd = layer.state_dict()
    sd["mel_stft.mel_scale.fb"] = sd["mel_stft.mel_scale.fb"] * 2.0
    layer.load_state_dict(sd)
    y1 = layer(x)
    assert torch.allclose(y1, y0 + float(np.log(2.0)), atol=1e-5)


def test_waveform_entry_equals_sliced_segments(cuda_device, native_lib):
    """ymt3_logmel_waveform_f32 (segmentation + tail padding fused) == slice_padded_array + forward, bit for bit."""
    from yourmt3_b200.audio_utils import slice_padded_array
    for layer in (S.Melspectrogram(), S.Spectrogram()):
        for n in (32767 * 3 + 5000, 32767 * 2, 100, 32767 * 4 - 1):
            wave = torch.from_numpy(synth_noise(1, n, seed=n % 1000)[0]).to(cuda_device)
            segs = slice_padded_array(wave, 32767, 32767)
            ref = layer(segs)
            got = layer.forward_waveform(wave)
            assert got.shape == ref.shape
            assert torch.equal(got, ref)


def test_bulk_copy_staging_equals_cooperative_fill(cuda_device, native_lib):
    """Interior frame pairs are staged by 1-D bulk async copies (16-byte aligned windows), edge pairs and misaligned
    buffers by the CTA itself: the same audio at a 4-byte-misaligned address (every pair cooperative) must give
    bit-identical output; several hops (stage size = hop + 2048 samples, > 48 KB of shared memory in total for the
    largest) against the float64 oracle."""
    for layer in (S.Melspectrogram(), S.Spectrogram()):
        x = torch.from_numpy(synth_noise(5, 32767, seed=3)).to(cuda_device)
        ref = layer(x)
        for shift in (1, 2, 3):
            buf = torch.zeros(5 * 32767 + 8, device=cuda_device)
            view = buf[shift: shift + 5 * 32767].view(5, 32767)
            view.copy_(x)
            assert view.data_ptr() % 16 != 0 and view.is_contiguous()
            assert torch.equal(layer(view), ref)
    for hop in (64, 128, 300, 512, 1000, 2048):
        import yourmt3_b200 as ymt3
        layer, (T, F) = S.get_spectrogram_layer_from_audio_cfg(ymt3.get_audio_cfg(codec="spec", hop_length=hop))
        xs = synth_noise(3, 20000, seed=hop)
        got, ref = run(layer, xs, cuda_device), oracle_run(layer, xs)
        assert got.shape == ref.shape == (3, 1 + 20000 // hop, F)
        assert_frontend_close(got, ref, TOL, strict_everywhere=False)
