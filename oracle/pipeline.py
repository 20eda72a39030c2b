"""End-to-end CPU oracle of ``YourMT3.inference`` (test infra only): frontend -> encoder ->
greedy decode, assembled from oracle/logmel.py and oracle/t5.py (and oracle/perceiver_tf.py
for the YPTF models).  Takes the product module's state dict so both sides share weights."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import logmel as OL
from . import t5 as OT


DEVICE = "cpu"   # the oracle is a CPU checker; tools/ref_gpu_bar.py sets "cuda" to time the SAME eager modules on the
                 # GPU (the "bar to beat" of SURVEY.md 2c) - never used by the product path


def _sub(sd: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    return {k[len(prefix):]: v.detach().to(DEVICE).float() for k, v in sd.items() if k.startswith(prefix)}


def frontend(sd, audio: np.ndarray, audio_cfg: Dict) -> torch.Tensor:
    a = np.asarray(audio, np.float32)
    if audio_cfg["codec"] == "melspec":
        y = OL.log_melspectrogram(a, sample_rate=audio_cfg["sample_rate"], n_fft=audio_cfg["n_fft"],
                                  hop_length=audio_cfg["hop_length"], f_min=audio_cfg["f_min"], f_max=audio_cfg["f_max"],
                                  n_mels=audio_cfg["n_mels"], power=audio_cfg.get("power", 1.0),
                                  eps=audio_cfg.get("log_eps", 1e-5),
                                  window=sd["spectrogram.mel_stft.spectrogram.window"].cpu().numpy(),
                                  fb=sd["spectrogram.mel_stft.mel_scale.fb"].cpu().numpy())
    else:
        y = OL.log_spectrogram(a, n_fft=audio_cfg["n_fft"], hop_length=audio_cfg["hop_length"],
                               power=audio_cfg.get("power", 1.0), eps=audio_cfg.get("log_eps", 1e-5),
                               bin0=1 if audio_cfg.get("spec_drop_dc", True) else 0,
                               n_bins=audio_cfg["n_fft"] // 2 + (0 if audio_cfg.get("spec_drop_dc", True) else 1),
                               window=sd["spectrogram.stft.window"].cpu().numpy())
    return torch.from_numpy(y)


def t5_encode(sd, feats: torch.Tensor, model_cfg: Dict, n_pos: int) -> torch.Tensor:
    ec = model_cfg["encoder"]["t5"]
    pos = OT.sinusoidal_positions(n_pos, ec["d_model"]).to(DEVICE) if ec.get("position_encoding_type") == "sinusoidal" else None
    return OT.t5_encoder(_sub(sd, "encoder."), feats.to(DEVICE), n_layers=ec["num_layers"], n_heads=ec["num_heads"],
                         eps=ec.get("layer_norm_epsilon", 1e-6), pos=pos)


def t5_generate(sd, enc_hs: torch.Tensor, model_cfg: Dict, n_pos: int, max_length: int, stop_at_eos=True,
                eos_id=1, pad_id=0, return_margins=False, prefix_ids=None):
    dc = model_cfg["decoder"][model_cfg["decoder_type"]]
    pos = OT.sinusoidal_positions(n_pos, dc["d_model"]).to(DEVICE) if dc.get("position_encoding_type") == "sinusoidal" else None
    embed = sd["embed_tokens.weight"].detach().to(DEVICE).float()
    lm = sd["lm_head.lm_head.weight"].detach().to(DEVICE).float()
    enc_hs = enc_hs.to(DEVICE)
    if enc_hs.dim() == 4:
        B, C, T, D = enc_hs.shape
        enc_hs = enc_hs.reshape(B * C, T, D)
    return OT.greedy_generate(_sub(sd, "decoder."), enc_hs, embed=embed, lm_head=lm, n_layers=dc["num_layers"],
                              n_heads=dc["num_heads"], max_length=max_length, eps=dc.get("layer_norm_epsilon", 1e-6),
                              prefix="", pos=pos, tie_word_embeddings=model_cfg["tie_word_embeddings"], eos_id=eos_id,
                              pad_id=pad_id, start_id=pad_id, stop_at_eos=stop_at_eos, return_margins=return_margins,
                              prefix_ids=prefix_ids)


def transcribe_t5(sd, audio: np.ndarray, audio_cfg: Dict, model_cfg: Dict, n_pos: int, max_length: int, **kw):
    with torch.no_grad():
        feats = frontend(sd, audio, audio_cfg)
        enc = t5_encode(sd, feats, model_cfg, n_pos)
        return t5_generate(sd, enc, model_cfg, n_pos, max_length, **kw)


def transcribe_from_feats(sd, feats: torch.Tensor, model_cfg, n_pos, max_length, **kw):
    """(B, T, F) log-(mel)spectrogram features -> tokens (encoder family dispatch)."""
    with torch.no_grad():
        if model_cfg["encoder_type"] == "t5":
            enc = t5_encode(sd, feats, model_cfg, n_pos)
        else:
            from . import perceiver_tf as OPTF
            enc = OPTF.encode(sd, feats, model_cfg)
        return t5_generate(sd, enc, model_cfg, n_pos, max_length, **kw)


def transcribe(sd, audio, audio_cfg, model_cfg, n_pos, max_length, **kw):
    """Dispatch on the encoder family (T5 | Perceiver-TF)."""
    if model_cfg["encoder_type"] == "t5":
        return transcribe_t5(sd, audio, audio_cfg, model_cfg, n_pos, max_length, **kw)
    with torch.no_grad():
        return transcribe_from_feats(sd, frontend(sd, audio, audio_cfg), model_cfg, n_pos, max_length, **kw)
