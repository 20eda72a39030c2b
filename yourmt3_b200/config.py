"""Dict-shaped configs with the upstream key names (amt/src/config/config.py [RECALL]).

The mounted reference has no code, so every default below is recalled from
upstream mimbres/YourMT3 / the YourMT3+ paper (arXiv 2407.04822) and is a config
FIELD, never a constant baked into a kernel.  PROVENANCE tags each one.
"""
from __future__ import annotations

import copy

audio_cfg = {
    "codec": "melspec",        # {melspec, spec}: melspec for MT3/T5, spec for Perceiver-TF
    "hop_length": 128,         # {128, 300}
    "audio_backend": "torchaudio",
    "sample_rate": 16000,
    "input_frames": 32767,     # samples per segment (~2.048 s)
    "n_fft": 2048,
    "n_mels": 512,             # melspec only
    "f_min": 50.0,
    "f_max": 8000.0,
    # not upstream keys (upstream hard-codes them in spectrogram.py [RECALL]); kept configurable
    "power": 1.0,              # magnitude spectrogram
    "log_eps": 1e-5,           # clamp floor before log
    "spec_drop_dc": True,      # spec codec keeps bins 1..1024 -> F = 1024
}

PROVENANCE = {
    "audio_cfg": "RECALL upstream config.py; n_fft/hop/T verified by probe (SURVEY 8a)",
    "audio_cfg.power/log_eps/spec_drop_dc": "RECALL upstream model/spectrogram.py (unverified)",
    "model_cfg.encoder.t5 / decoder.t5": "RECALL upstream config.py = google/t5-v1_1-small shape",
    "model_cfg.encoder.perceiver-tf": "RECALL upstream config.py + YourMT3+ paper table (YPTF.MoE+Multi flags)",
    "vocab_size": "RECALL (event codec: 3 special + 206 shift + 128 pitch + 2 vel + 1 tie + 128 program + 128 drum = 596)",
}

model_cfg = {
    "encoder_type": "t5",          # {"t5", "perceiver-tf"}
    "decoder_type": "t5",          # {"t5", "multi-t5"}
    "pre_encoder_type": "default",
    "pre_encoder_type_default": {"t5": None, "perceiver-tf": "conv"},
    "pre_decoder_type": "default",
    "pre_decoder_type_default": {
        "t5": {"t5": None},
        "perceiver-tf": {"t5": "linear", "multi-t5": "mc_shared_linear"},
    },
    "conv_out_channels": 128,
    "use_task_conditional_encoder": True,
    "use_task_conditional_decoder": True,
    "d_feat": "auto",
    "tie_word_embeddings": True,
    "vocab_size": 596,
    "num_max_positions": "auto",
    "encoder": {
        "t5": {
            "d_model": 512, "num_heads": 6, "d_kv": 64, "num_layers": 8, "dropout_rate": 0.05,
            "position_encoding_type": "sinusoidal", "ff_widening_factor": 2, "ff_layer_type": "t5_gmlp",
            "layer_norm_epsilon": 1e-6,
        },
        "perceiver-tf": {
            "num_latents": 24, "d_latent": 128, "d_model": "q", "num_blocks": 3,
            "num_local_transformers_per_block": 2, "num_temporal_transformers_per_block": 2,
            "sca_use_query_residual": False, "dropout_rate": 0.1,
            "position_encoding_type": "trainable", "attention_to_channel": True,
            "layer_norm_type": "layer_norm", "ff_layer_type": "mlp", "ff_widening_factor": 1,
            "moe_num_experts": 4, "moe_topk": 2, "hidden_act": "gelu",
            "num_cross_attention_heads": 1, "num_self_attention_heads": 8,
            "rotary_type_sca": "pixel", "rotary_type_latent": "pixel", "rotary_type_temporal": "lang",
            "rotary_apply_to_keys": False, "rotary_partial_pe": False,
            "layer_norm_eps": 1e-5,
        },
    },
    "decoder": {
        "t5": {
            "d_model": 512, "num_heads": 6, "d_kv": 64, "num_layers": 8, "dropout_rate": 0.05,
            "position_encoding_type": "sinusoidal", "ff_widening_factor": 2, "ff_layer_type": "t5_gmlp",
            "layer_norm_epsilon": 1e-6,
        },
        "multi-t5": {
            "d_model": 512, "num_heads": 6, "d_kv": 64, "num_layers": 8, "dropout_rate": 0.05,
            "position_encoding_type": "sinusoidal", "ff_widening_factor": 2, "ff_layer_type": "t5_gmlp",
            "layer_norm_epsilon": 1e-6, "num_channels": 13,
        },
    },
    "feat_length": "auto",
    "event_length": 1024,          # max decoded tokens (256 for multi-t5 / mc13)
    "init_factor": 1.0,
}


def get_audio_cfg(**overrides) -> dict:
    cfg = copy.deepcopy(audio_cfg)
    cfg.update(overrides)
    return cfg


def get_model_cfg(preset: str = "mt3_t5_small", **overrides) -> dict:
    """Named presets for the three model families in BASELINE.json."""
    cfg = copy.deepcopy(model_cfg)
    if preset == "mt3_t5_small":
        pass
    elif preset == "yptf":            # Perceiver-TF encoder + single-channel T5 decoder
        cfg["encoder_type"] = "perceiver-tf"
        cfg["decoder_type"] = "t5"
    elif preset == "yptf_moe_multi":  # YPTF.MoE+Multi (YourMT3+ best model) [RECALL flags]
        cfg["encoder_type"] = "perceiver-tf"
        cfg["decoder_type"] = "multi-t5"
        enc = cfg["encoder"]["perceiver-tf"]
        enc.update(num_latents=26, sca_use_query_residual=True, position_encoding_type="rope",
                   rotary_partial_pe=True, ff_layer_type="moe", ff_widening_factor=4, moe_num_experts=8,
                   moe_topk=2, hidden_act="silu", layer_norm_type="rms_norm", dropout_rate=0.05)
        cfg["event_length"] = 256
    else:
        raise ValueError(f"unknown preset {preset!r}")
    cfg.update(overrides)
    return cfg
