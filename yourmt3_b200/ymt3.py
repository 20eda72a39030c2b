"""``YourMT3`` with the reference's inference API surface (upstream amt/src/model/ymt3.py [RECALL]):

    model.spectrogram / .pre_encoder / .encoder / .pre_decoder / .decoder / .embed_tokens / .lm_head
    model.encode(x)                         -> encoder hidden states
    model.inference(x, task_tokens=None)    -> (B, L) | (B, C, L) token ids
    model.inference_file(bsz, audio_segments) -> list of per-batch numpy arrays

Only the inference path exists (training / Lightning hooks are out of scope).
Every stage is a native call; nothing falls back to eager PyTorch.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
from torch import nn

from . import _lib
from .config import get_audio_cfg, get_model_cfg
from .lm_head import LMHead
from .spectrogram import get_spectrogram_layer_from_audio_cfg
from .t5mod import MultiChannelT5Decoder, T5DecoderYMT3, T5EncoderYMT3
from .t5mod_helper import task_cond_dec_generate


class YourMT3(nn.Module):
    def __init__(self, audio_cfg: Optional[Dict] = None, model_cfg: Optional[Dict] = None, precision: str = "f32",
                 eos_id: int = 1, pad_id: int = 0, **unused):
        super().__init__()
        self.audio_cfg = audio_cfg = audio_cfg or get_audio_cfg()
        self.model_cfg = model_cfg = model_cfg or get_model_cfg()
        if precision not in ("f32", "bf16"):
            raise ValueError("precision must be 'f32' or 'bf16'")
        self.precision = precision
        self._prec = {"f32": _lib.DTYPE_F32, "bf16": _lib.DTYPE_BF16}[precision]
        self.eos_id, self.pad_id = eos_id, pad_id
        self.decode_lanes = 1   # >1: concurrent decode lanes (see t5mod_helper.task_cond_dec_generate)
        # bf16 multi-channel models: run the decoder's cross-attention in absorbed form on the encoder latents
        # (same function, 24x fewer cross-attention bytes per step; include/ymt3_b200.h ymt3_t5dec_generate_latent)
        self.absorb_cross_attention = True
        self.encoder_type, self.decoder_type = model_cfg["encoder_type"], model_cfg["decoder_type"]
        self.vocab_size = int(model_cfg["vocab_size"])
        self.max_token_length = int(model_cfg["event_length"])
        self.tie_word_embeddings = bool(model_cfg["tie_word_embeddings"])

        self.spectrogram, (self.feat_length, self.feat_dim) = get_spectrogram_layer_from_audio_cfg(audio_cfg)
        dec_cfg = dict(model_cfg["decoder"][self.decoder_type])
        dec_cfg["vocab_size"] = self.vocab_size
        n_pos = max(self.feat_length, self.max_token_length) + 16   # room for task-prefix tokens

        pre_enc = model_cfg["pre_encoder_type"]
        if pre_enc == "default":
            pre_enc = model_cfg["pre_encoder_type_default"][self.encoder_type]
        pre_dec = model_cfg["pre_decoder_type"]
        if pre_dec == "default":
            pre_dec = model_cfg["pre_decoder_type_default"][self.encoder_type][self.decoder_type]

        if self.encoder_type == "t5":
            enc_cfg = dict(model_cfg["encoder"]["t5"])
            if enc_cfg["d_model"] != self.feat_dim:
                raise ValueError("T5 encoder consumes the spectrogram directly: d_model must equal the feature width")
            self.pre_encoder = nn.Identity() if pre_enc is None else self._unsupported("pre_encoder", pre_enc)
            self.encoder = T5EncoderYMT3(enc_cfg, precision=precision, num_max_positions=n_pos)
            self.pre_decoder = nn.Identity() if pre_dec is None else self._unsupported("pre_decoder", pre_dec)
        elif self.encoder_type == "perceiver-tf":
            from .perceiver_mod import build_perceiver_tf_stages  # noqa: WPS433 (kept optional until built)
            self.pre_encoder, self.encoder, self.pre_decoder = build_perceiver_tf_stages(
                self, audio_cfg, model_cfg, pre_enc, pre_dec, precision)
        else:
            raise NotImplementedError(f"encoder_type={self.encoder_type!r} (conformer is out of scope, SURVEY 2b)")

        if self.decoder_type == "t5":
            self.decoder = T5DecoderYMT3(dec_cfg, num_max_positions=n_pos)
        elif self.decoder_type == "multi-t5":
            self.decoder = MultiChannelT5Decoder(dec_cfg, num_max_positions=n_pos)
        else:
            raise NotImplementedError(f"decoder_type={self.decoder_type!r}")
        self.embed_tokens = nn.Embedding(self.vocab_size, dec_cfg["d_model"])
        self.lm_head = LMHead(dec_cfg, model_cfg.get("init_factor", 1.0), self.tie_word_embeddings)
        if self.tie_word_embeddings:
            self.lm_head.lm_head.weight = self.embed_tokens.weight
        self.eval()

    @staticmethod
    def _unsupported(kind, name):
        raise NotImplementedError(f"{kind} type {name!r} is not on the benchmarked path")

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """x: (B, 1, L) f32 CUDA audio -> encoder hidden states for the decoder
        ((B, T, D) or (B, C, T, D)), in the model precision."""
        feats = self.spectrogram(x)                                   # (B, T, F) f32
        feats = self.pre_encoder(feats)
        enc_hs = self.encoder(inputs_embeds=feats)["last_hidden_state"]
        return self.pre_decoder(enc_hs)

    @torch.no_grad()
    def inference(self, x: torch.Tensor, task_tokens: Optional[torch.Tensor] = None,
                  max_token_length: Optional[int] = None, stop_at_eos: bool = True,
                  early_stop_interval: int = 0, decode_lanes: Optional[int] = None, **unused) -> torch.Tensor:
        """x: (B, 1, L) audio segments -> token ids (B, L_tok) or (B, C, L_tok) (LongTensor, CUDA)."""
        max_len = max_token_length or self.max_token_length
        if self._absorbed():
            feats = self.pre_encoder(self.spectrogram(x))
            return self._generate_absorbed(self.encoder(inputs_embeds=feats)["last_hidden_state"], task_tokens, max_len,
                                           stop_at_eos, early_stop_interval)
        enc_hs = self.encode(x)
        return task_cond_dec_generate(self.decoder, self.decoder_type, self.embed_tokens, self.lm_head, enc_hs,
                                      prefix_ids=task_tokens, max_length=max_len, stop_at_eos=stop_at_eos,
                                      eos_id=self.eos_id, pad_id=self.pad_id, decoder_start_token_id=self.pad_id,
                                      precision=self._prec, early_stop_interval=early_stop_interval,
                                      lanes=self.decode_lanes if decode_lanes is None else decode_lanes)

    def _absorbed(self) -> bool:
        return (self.absorb_cross_attention and self.precision == "bf16" and self.decoder_type == "multi-t5"
                and getattr(self.pre_decoder, "kind", None) == "mc_shared_linear"
                and self.pre_decoder.proj.in_features == 256 and self.decoder.config["num_heads"] <= 8
                and self.feat_length <= 128)   # kernel limits (cross_absorbed.cu): zdim 256, <= 8 heads, T_enc <= 128

    def _generate_absorbed(self, latents, task_tokens, max_len, stop_at_eos=True, early_stop_interval=0):
        return task_cond_dec_generate(self.decoder, self.decoder_type, self.embed_tokens, self.lm_head, latents,
                                      prefix_ids=task_tokens, max_length=max_len, stop_at_eos=stop_at_eos,
                                      eos_id=self.eos_id, pad_id=self.pad_id, decoder_start_token_id=self.pad_id,
                                      precision=self._prec, early_stop_interval=early_stop_interval,
                                      cross_proj=self.pre_decoder.proj)

    @torch.no_grad()
    def score(self, x: torch.Tensor, target_tokens: torch.Tensor, logit_steps=()):
        """Teacher-forced decoder pass (the decoder side of the reference's ``forward(x, target_tokens)``): x (B, 1, L)
        audio, target_tokens (B, Ltok) / (B, C, Ltok).  Returns (argmax like target_tokens, logits
        (len(logit_steps), B[*C], vocab) f32) computed by the SAME kernels / CUDA graph as ``inference`` (absorbed
        cross-attention included when that is the path ``inference`` takes)."""
        from .t5mod_helper import DecoderRuntime
        L = int(target_tokens.shape[-1])
        if self._absorbed():
            enc = self.encoder(inputs_embeds=self.pre_encoder(self.spectrogram(x)))["last_hidden_state"]
            B, T, K, Dl = enc.shape
            zin = self.pre_decoder.proj.in_features
            enc = enc.reshape(B, T, K * Dl // zin, zin)
            proj = self.pre_decoder.proj
        else:
            enc = self.encode(x)
            if enc.dim() == 4:
                enc = enc.reshape(-1, enc.shape[2], enc.shape[3])
            proj = None
        rt = getattr(self.decoder, "_runtime_score", None)
        if rt is None or rt.max_length < L or rt._cross_proj is not proj or rt.precision != self._prec:
            if rt is not None:
                rt.free_native()
            rt = DecoderRuntime(self.decoder, self.embed_tokens, self.lm_head, self._prec, self.vocab_size,
                                max(L, self.max_token_length), self.tie_word_embeddings, self.eos_id, self.pad_id, self.pad_id,
                                cross_proj=proj)
            object.__setattr__(self.decoder, "_runtime_score", rt)
        argmax, logits = rt.score_forced(enc, target_tokens, logit_steps)
        return argmax.long().view(target_tokens.shape), logits

    @torch.no_grad()
    def transcribe_waveform(self, wave: torch.Tensor, bsz: int = 256, **kw) -> torch.Tensor:
        """Whole mono 16 kHz waveform (n_samples,) on the model's device -> tokens (n_seg, L) / (n_seg, C, L).
        The frontend consumes the waveform directly (segmentation + tail zero-padding fused into its loads,
        SURVEY 8f.1); the model then runs in batches of ``bsz`` segments."""
        feats = self.spectrogram.forward_waveform(wave, self.audio_cfg["input_frames"])
        max_len = kw.pop("max_token_length", None) or self.max_token_length
        outs = []
        for i in range(0, feats.shape[0], bsz):
            if self._absorbed():
                lat = self.encoder(inputs_embeds=self.pre_encoder(feats[i:i + bsz]))["last_hidden_state"]
                outs.append(self._generate_absorbed(lat, None, max_len, kw.get("stop_at_eos", True)))
                continue
            enc_hs = self.pre_decoder(self.encoder(inputs_embeds=self.pre_encoder(feats[i:i + bsz]))["last_hidden_state"])
            outs.append(task_cond_dec_generate(self.decoder, self.decoder_type, self.embed_tokens, self.lm_head, enc_hs,
                                               max_length=max_len, stop_at_eos=kw.get("stop_at_eos", True),
                                               eos_id=self.eos_id, pad_id=self.pad_id, decoder_start_token_id=self.pad_id,
                                               precision=self._prec))
        return torch.cat(outs, 0)

    @torch.no_grad()
    def inference_file_sharded(self, bsz: int, audio_segments: torch.Tensor, **kw) -> torch.Tensor:
        """Multi-GPU long-form path (one process per GPU, torch.distributed initialised with NCCL): this
        rank transcribes its contiguous range of segments; ONE all-gather returns the ordered tokens
        (n_seg, L) / (n_seg, C, L) int32 on every rank."""
        from .sharding import transcribe_sharded
        dev = next(self.parameters()).device
        L = kw.get("max_token_length") or self.max_token_length
        nch = getattr(self.decoder, "num_channels", 1)
        shape = (nch, L) if self.decoder_type == "multi-t5" else (L,)
        return transcribe_sharded(lambda x: self.inference(x, None, **kw), audio_segments, bsz, dev, self.pad_id,
                                  token_shape=shape)

    @torch.no_grad()
    def inference_file(self, bsz: int, audio_segments: torch.Tensor, note_token_array=None, task_token_array=None,
                       **kw) -> List[np.ndarray]:
        """audio_segments: (n_seg, 1, L) f32 (CPU or CUDA).  Returns a list with one int array of
        predicted token ids per batch of ``bsz`` segments (reference return shape [RECALL])."""
        n = audio_segments.shape[0]
        dev = next(self.parameters()).device
        out = []
        for i in range(0, n, bsz):
            x = audio_segments[i:i + bsz].to(dev, torch.float32, non_blocking=True)
            out.append(self.inference(x, None, **kw).cpu().numpy())
        return out
