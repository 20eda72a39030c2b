"""Spectrogram layers with the upstream interface (amt/src/model/spectrogram.py [RECALL]):
``get_spectrogram_layer_from_audio_cfg(audio_cfg) -> (layer, (T, F))``,
``Melspectrogram`` / ``Spectrogram`` : ``(B, 1, L) f32 -> (B, T, F) f32``.

The arithmetic of the reference path is torchaudio's
(SP/torchaudio/transforms/_transforms.py:621-631, functional/functional.py:54-145,
:518-587) followed by ``log(clamp(x, eps))``.  Here ``forward`` is ONE fused
sm_100a kernel called through the C ABI (``ymt3_logmel_f32``); there is no CPU
or eager fallback.  The module keeps the buffers torchaudio's transforms
register (``window``, ``fb``) under the same state-dict keys, so a checkpoint's
frontend buffers drop in.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Tuple

import torch
from torch import nn

from . import _lib


def _hz_to_mel_htk(freq: float) -> float:
    return 2595.0 * math.log10(1.0 + (freq / 700.0))


def mel_filterbank(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """HTK triangular filterbank (n_freqs, n_mels), norm=None; same fp32 torch op
    sequence as SP/torchaudio/functional/functional.py:563-573,505-513 so the buffer
    is bit-identical to ``torchaudio.functional.melscale_fbanks`` (checked in tests)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel_htk(f_min), _hz_to_mel_htk(f_max), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


class _Buffers(nn.Module):
    """Name-only container so state-dict keys match torchaudio's nested transforms."""


class _FusedFrontend(nn.Module):
    codec = "melspec"

    def __init__(self, n_fft: int, hop_length: int, power: float, eps: float):
        super().__init__()
        if float(power) not in (1.0, 2.0):
            raise ValueError("power must be 1.0 or 2.0")
        self.n_fft, self.hop_length, self.power, self.eps = int(n_fft), int(hop_length), float(power), float(eps)
        self._handle = None
        self._handle_key = None

    # -- buffers (overridden) -------------------------------------------------
    def _window(self) -> torch.Tensor:
        raise NotImplementedError

    def _fb(self):
        return None

    def _cfg(self) -> _lib.AudioCfg:
        raise NotImplementedError

    def num_frames(self, num_samples: int) -> int:
        return 1 + num_samples // self.hop_length

    # -- native handle --------------------------------------------------------
    def _get_handle(self, device: torch.device):
        win, fb = self._window(), self._fb()
        key = (device.index, win.data_ptr(), win._version, None if fb is None else (fb.data_ptr(), fb._version))
        if self._handle is not None and key == self._handle_key:
            return self._handle
        self._free_handle()
        lib = _lib.load()
        win_h = win.detach().to("cpu", torch.float32).contiguous()
        fb_h = None if fb is None else fb.detach().to("cpu", torch.float32).contiguous()
        h = C.c_void_p()
        cfg = self._cfg()
        with torch.cuda.device(device):
            rc = lib.ymt3_frontend_create(C.byref(cfg), win_h.data_ptr(), None if fb_h is None else fb_h.data_ptr(),
                                          C.byref(h))
        _lib.check(rc, "frontend_create")
        self._handle, self._handle_key = h, key
        return h

    def _free_handle(self):
        if self._handle is not None:
            _lib.load().ymt3_frontend_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free_handle()
        except Exception:
            pass

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: (B, 1, L) or (B, L) float32 CUDA -> (B, T, F) float32."""
        if x.dim() == 3:
            if x.shape[1] != 1:
                raise ValueError("expected mono audio (B, 1, L)")
            x = x[:, 0, :]
        if not x.is_cuda:
            raise RuntimeError("yourmt3_b200 frontend runs on CUDA only (no CPU fallback)")
        if x.dtype != torch.float32:
            raise TypeError("audio must be float32")
        x = x.contiguous()
        B, L = x.shape
        h = self._get_handle(x.device)
        out = torch.empty((B, self.num_frames(L), self.num_features), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.load().ymt3_logmel_f32(h, x.data_ptr(), B, L, out.data_ptr(), _lib.current_stream_ptr())
        _lib.check(rc, "logmel")
        return out

    def forward_waveform(self, wave: torch.Tensor, segment_length: int = 32767) -> torch.Tensor:
        """wave: (n_samples,) or (1, n_samples) float32 CUDA -> (n_seg, T, F).  Equivalent to
        ``forward(slice_padded_array(wave, L, L))`` (upstream utils/audio.py) with the slicing and the zero
        padding of the tail fused into the kernel's loads."""
        if not wave.is_cuda or wave.dtype != torch.float32:
            raise RuntimeError("forward_waveform expects a float32 CUDA tensor (no CPU fallback)")
        wave = wave.reshape(-1).contiguous()
        n = wave.numel()
        lib = _lib.load()
        n_seg = lib.ymt3_num_segments(n, segment_length)
        h = self._get_handle(wave.device)
        out = torch.empty((n_seg, self.num_frames(segment_length), self.num_features), dtype=torch.float32,
                          device=wave.device)
        with torch.cuda.device(wave.device):
            rc = lib.ymt3_logmel_waveform_f32(h, wave.data_ptr(), n, segment_length, out.data_ptr(),
                                              _lib.current_stream_ptr())
        _lib.check(rc, "logmel_waveform")
        return out

    def forward_host(self, x: torch.Tensor) -> torch.Tensor:
        """End-to-end call with HOST buffers (H2D + kernel + D2H inside the C ABI)."""
        if x.dim() == 3:
            x = x[:, 0, :]
        if x.is_cuda or x.dtype != torch.float32:
            raise TypeError("forward_host expects a float32 CPU tensor")
        x = x.contiguous()
        B, L = x.shape
        dev = torch.device("cuda", torch.cuda.current_device())
        h = self._get_handle(dev)
        out = torch.empty((B, self.num_frames(L), self.num_features), dtype=torch.float32, pin_memory=True)
        rc = _lib.load().ymt3_logmel_host_f32(h, x.data_ptr(), B, L, out.data_ptr(), _lib.current_stream_ptr())
        _lib.check(rc, "logmel_host")
        return out


class Melspectrogram(_FusedFrontend):
    codec = "melspec"

    def __init__(self, audio_backend: str = "torchaudio", sample_rate: int = 16000, n_fft: int = 2048,
                 hop_length: int = 128, f_min: float = 50.0, f_max: float = 8000.0, n_mels: int = 512,
                 eps: float = 1e-5, power: float = 1.0, **kwargs):
        super().__init__(n_fft, hop_length, power, eps)
        self.sample_rate, self.f_min, self.f_max, self.n_mels = sample_rate, f_min, f_max, n_mels
        # same nesting as torchaudio.transforms.MelSpectrogram: .spectrogram.window / .mel_scale.fb
        self.mel_stft = _Buffers()
        self.mel_stft.spectrogram = _Buffers()
        self.mel_stft.mel_scale = _Buffers()
        self.mel_stft.spectrogram.register_buffer("window", torch.hann_window(n_fft))
        self.mel_stft.mel_scale.register_buffer(
            "fb", mel_filterbank(n_fft // 2 + 1, f_min, f_max, n_mels, sample_rate))
        self.num_features = n_mels

    def _window(self):
        return self.mel_stft.spectrogram.window

    def _fb(self):
        return self.mel_stft.mel_scale.fb

    def _cfg(self):
        return _lib.AudioCfg(n_fft=self.n_fft, hop_length=self.hop_length, codec=_lib.CODEC_MELSPEC,
                             n_mels=self.n_mels, spec_bin0=0, spec_bins=0, power_mode=int(self.power),
                             log_eps=self.eps)


class Spectrogram(_FusedFrontend):
    codec = "spec"

    def __init__(self, audio_backend: str = "torchaudio", n_fft: int = 2048, hop_length: int = 300,
                 eps: float = 1e-5, power: float = 1.0, drop_dc: bool = True, **kwargs):
        super().__init__(n_fft, hop_length, power, eps)
        self.stft = _Buffers()
        self.stft.register_buffer("window", torch.hann_window(n_fft))
        self.bin0 = 1 if drop_dc else 0
        self.num_features = n_fft // 2 + 1 - self.bin0

    def _window(self):
        return self.stft.window

    def _cfg(self):
        return _lib.AudioCfg(n_fft=self.n_fft, hop_length=self.hop_length, codec=_lib.CODEC_SPEC, n_mels=0,
                             spec_bin0=self.bin0, spec_bins=self.num_features, power_mode=int(self.power),
                             log_eps=self.eps)


def get_spectrogram_layer_from_audio_cfg(audio_cfg: Dict) -> Tuple[nn.Module, Tuple[int, int]]:
    """Upstream signature: returns ``(layer, (T, F))`` for ``audio_cfg['input_frames']`` samples."""
    codec = audio_cfg["codec"]
    common = dict(n_fft=audio_cfg["n_fft"], hop_length=audio_cfg["hop_length"],
                  eps=audio_cfg.get("log_eps", 1e-5), power=audio_cfg.get("power", 1.0))
    if codec == "melspec":
        layer = Melspectrogram(sample_rate=audio_cfg["sample_rate"], f_min=audio_cfg["f_min"],
                               f_max=audio_cfg["f_max"], n_mels=audio_cfg["n_mels"], **common)
    elif codec == "spec":
        layer = Spectrogram(drop_dc=audio_cfg.get("spec_drop_dc", True), **common)
    else:
        raise ValueError(f"unknown codec {codec!r}")
    return layer, (layer.num_frames(audio_cfg["input_frames"]), layer.num_features)
