"""Greedy segment-batched generation with the reference's helper signature
(upstream amt/src/model/t5mod_helper.py ``task_cond_dec_generate`` [RECALL]).

Reference loop: Python ``for`` over steps, one decoder call + lm_head + argmax per step and a
device->host sync per step for the EOS test.  Here: ONE C-ABI call (``ymt3_t5dec_generate``)
that pre-computes the cross-attention K/V, captures a decode step into a CUDA graph whose
kernels read the step counter / finished mask from device memory, and replays it
``max_length`` times; tokens, KV cache and masks stay in HBM, no host round trip per step.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .t5mod import T5DecoderYMT3, _NativeOwner, _cfg_struct


def fold_cross_projection(decoder: T5DecoderYMT3, proj: torch.nn.Linear) -> dict:
    """Absorbed cross-attention weights (include/ymt3_b200.h, ``ymt3_t5dec_generate_latent``).

    With ``enc_hs = proj(z)`` every layer's cross K/V are linear in the latent ``z``, so per head h
    ``q_h . K_h,t = ((Wk_h Wp)^T Wq_h x) . z_t + const`` and ``sum_t p_t V_h,t = Wv_h Wp (sum_t p_t z_t) + Wv_h bp``.
    Folded in fp64 on the weights' device, returned as f32 named tensors (state-dict style keys)."""
    H, dk = decoder.config["num_heads"], decoder.config.get("d_kv", 64)
    Wp, bp = proj.weight.detach().double(), proj.bias.detach().double()          # (D, Z), (D)
    out = {}
    for i, blk in enumerate(decoder.block):
        att = blk.layer[1].EncDecAttention
        Wq, Wk, Wv, Wo = (w.weight.detach().double() for w in (att.q, att.k, att.v, att.o))
        KP = (Wk @ Wp).view(H, dk, -1)                                           # (H, dk, Z)
        VP = (Wv @ Wp).view(H, dk, -1)
        q_abs = torch.einsum("hdz,hdm->hzm", KP, Wq.view(H, dk, -1))             # (H, Z, D)
        o_abs = torch.einsum("mhd,hdz->mhz", Wo.view(-1, H, dk), VP)             # (D, H, Z)
        key = f"block.{i}.layer.1.EncDecAttention."
        out[key + "q_absorbed.weight"] = q_abs.reshape(-1, q_abs.shape[-1]).float().contiguous()
        out[key + "o_absorbed.weight"] = o_abs.reshape(o_abs.shape[0], -1).float().contiguous()
        out[key + "o_absorbed.bias"] = (Wo @ (Wv @ bp)).float().contiguous()
    return out


class DecoderRuntime(_NativeOwner):
    """Native decoder handle built from (decoder, embed_tokens, lm_head) parameters."""

    _destroy_name = "ymt3_t5dec_destroy"

    def __init__(self, decoder: T5DecoderYMT3, embed_tokens, lm_head, precision: int, vocab_size: int,
                 max_length: int, tie_word_embeddings: bool, eos_id=1, pad_id=0, start_id=0, cross_proj=None):
        super().__init__()
        # plain attribute refs (not registered as submodules: they are owned by YourMT3)
        object.__setattr__(self, "_decoder", decoder)
        object.__setattr__(self, "_embed", embed_tokens)
        object.__setattr__(self, "_lm", lm_head)
        object.__setattr__(self, "_cross_proj", cross_proj)   # nn.Linear of an `mc_shared_linear` pre_decoder or None
        self._fold_cache = None
        self.precision, self.vocab_size, self.max_length = precision, vocab_size, max_length
        self.tie, self.eos_id, self.pad_id, self.start_id = tie_word_embeddings, eos_id, pad_id, start_id

    def _tensors(self):
        named = dict(self._decoder.named_parameters())
        if getattr(self._decoder, "pos_table", None) is not None:
            named["pos_table"] = self._decoder.pos_table
        if getattr(self._decoder, "has_relative_attention_bias", False):
            from .t5mod import relative_bias_by_distance
            w = named.pop("block.0.layer.0.SelfAttention.relative_attention_bias.weight")
            n = max(self._decoder.num_max_positions, self.max_length)
            named["relative_bias_by_distance"] = relative_bias_by_distance(
                w, n, False, self._decoder.config.get("relative_attention_max_distance", 128))
        named["embed_tokens.weight"] = self._embed.weight
        lm_w = self._lm.lm_head.weight if hasattr(self._lm, "lm_head") else self._lm.weight
        if not self.tie or lm_w.data_ptr() != self._embed.weight.data_ptr():
            named["lm_head.weight"] = lm_w
        if self._cross_proj is not None:
            src = list(named.values()) + [self._cross_proj.weight, self._cross_proj.bias]
            key = tuple((v.data_ptr(), v._version) for v in src)
            if self._fold_cache is None or self._fold_cache[0] != key:
                self._fold_cache = (key, fold_cross_projection(self._decoder, self._cross_proj))
            named.update(self._fold_cache[1])
        return named

    def _create(self, arr, n):
        h = C.c_void_p()
        cfg = _cfg_struct(self._decoder.config, self.precision, vocab=self.vocab_size, max_length=self.max_length,
                          tie=self.tie, eos=self.eos_id, pad=self.pad_id, start=self.start_id)
        _lib.check(_lib.load().ymt3_t5dec_create(C.byref(cfg), arr, n, C.byref(h)), "t5dec_create")
        return h

    def generate(self, enc_hs: torch.Tensor, max_length: int, stop_at_eos: bool = True,
                 early_stop_interval: int = 0, prefix_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
        """enc_hs: (N, T_enc, d_model) in the runtime's precision -> (N, max_length) int32 CUDA."""
        want = _lib.torch_dtype(self.precision)
        if not enc_hs.is_cuda:
            raise RuntimeError("generation runs on CUDA only (no CPU fallback)")
        if enc_hs.dtype != want:
            enc_hs = enc_hs.to(want)
        enc_hs = enc_hs.contiguous()
        N, T, _ = enc_hs.shape
        h = self.native()
        tokens = torch.empty((N, max_length), dtype=torch.int32, device=enc_hs.device)
        P, pfx_ptr = 0, None
        if prefix_ids is not None and prefix_ids.numel() > 0:
            pfx = prefix_ids.to(enc_hs.device, torch.int32).reshape(N, -1).contiguous()
            P, pfx_ptr = pfx.shape[1], pfx.data_ptr()
        with torch.cuda.device(enc_hs.device):
            _lib.check(_lib.load().ymt3_t5dec_generate_prefixed(h, enc_hs.data_ptr(), N, T, pfx_ptr, P, max_length,
                                                                int(stop_at_eos), int(early_stop_interval),
                                                                tokens.data_ptr(), _lib.current_stream_ptr()),
                       "t5dec_generate")
        return tokens

    def generate_latent(self, latents: torch.Tensor, max_length: int, stop_at_eos: bool = True,
                        early_stop_interval: int = 0, prefix_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
        """latents: (B, T_enc, C, zdim) encoder latents regrouped per channel (bf16) -> (B*C, max_length) int32."""
        if self._cross_proj is None or self.precision != _lib.DTYPE_BF16:
            raise RuntimeError("absorbed cross-attention needs a bf16 runtime built with cross_proj")
        if not latents.is_cuda:
            raise RuntimeError("generation runs on CUDA only (no CPU fallback)")
        latents = latents.to(torch.bfloat16).contiguous()
        B, T, Cn, Z = latents.shape
        if Z != self._cross_proj.in_features:
            raise ValueError(f"latent width {Z} != projection input {self._cross_proj.in_features}")
        h = self.native()
        tokens = torch.empty((B * Cn, max_length), dtype=torch.int32, device=latents.device)
        P, pfx_ptr = 0, None
        if prefix_ids is not None and prefix_ids.numel() > 0:
            pfx = prefix_ids.to(latents.device, torch.int32).reshape(B * Cn, -1).contiguous()
            P, pfx_ptr = pfx.shape[1], pfx.data_ptr()
        with torch.cuda.device(latents.device):
            _lib.check(_lib.load().ymt3_t5dec_generate_latent(h, latents.data_ptr(), B, T, Cn, pfx_ptr, P, max_length,
                                                              int(stop_at_eos), int(early_stop_interval),
                                                              tokens.data_ptr(), _lib.current_stream_ptr()),
                       "t5dec_generate_latent")
        return tokens

    def score_forced(self, enc: torch.Tensor, forced_ids: torch.Tensor, logit_steps=()):
        """Teacher-forced scoring (``ymt3_t5dec_score_forced``): decoder inputs [start, forced[:, :-1]].

        enc: (N, T_enc, d_model) hidden states, or (B, T_enc, C, zdim) latents on a runtime built with ``cross_proj``
        (absorbed cross-attention; N = B*C).  forced_ids: (N, L).  Returns (argmax (N, L) int32, logits
        (len(logit_steps), N, vocab) f32): the greedy choice at every step and the logits of the selected steps."""
        if not enc.is_cuda:
            raise RuntimeError("scoring runs on CUDA only (no CPU fallback)")
        latent = enc.dim() == 4
        if latent and (self._cross_proj is None or self.precision != _lib.DTYPE_BF16):
            raise RuntimeError("latent scoring needs a bf16 runtime built with cross_proj")
        enc = enc.to(_lib.torch_dtype(self.precision)).contiguous()
        if latent:
            B, T, Cn, _ = enc.shape
            N = B * Cn
        else:
            (B, T, _), Cn = enc.shape, 0
            N = B
        forced = forced_ids.to(enc.device, torch.int32).reshape(N, -1).contiguous()
        L = forced.shape[1]
        steps = (C.c_int32 * max(1, len(logit_steps)))(*[int(v) for v in logit_steps])
        argmax = torch.empty((N, L), dtype=torch.int32, device=enc.device)
        logits = torch.empty((len(logit_steps), N, self.vocab_size), dtype=torch.float32, device=enc.device)
        with torch.cuda.device(enc.device):
            _lib.check(_lib.load().ymt3_t5dec_score_forced(self.native(), enc.data_ptr(), B, T, Cn, forced.data_ptr(), L,
                                                           argmax.data_ptr(), steps, len(logit_steps),
                                                           logits.data_ptr() if len(logit_steps) else None,
                                                           _lib.current_stream_ptr()), "t5dec_score_forced")
        return argmax, logits

    def last_logits(self, N: int, device) -> torch.Tensor:
        out = torch.empty((N, self.vocab_size), dtype=torch.float32, device=device)
        _lib.check(_lib.load().ymt3_t5dec_last_logits(self.native(), out.data_ptr(), N, _lib.current_stream_ptr()),
                   "t5dec_last_logits")
        return out


def task_cond_dec_generate(decoder, decoder_type: str, embed_tokens, lm_head, encoder_hidden_states: torch.Tensor,
                           shift_right_fn=None, prefix_ids: Optional[torch.Tensor] = None, max_length: int = 1024,
                           stop_at_eos: bool = True, eos_id: int = 1, pad_id: int = 0,
                           decoder_start_token_id: int = 0, precision: int = _lib.DTYPE_F32,
                           early_stop_interval: int = 0, lanes: int = 1, cross_proj=None, **unused) -> torch.Tensor:
    """Returns LongTensor (B, max_length) for 't5' or (B, C, max_length) for 'multi-t5'.

    ``encoder_hidden_states``: (B, T, D) or (B, C, T, D).  ``prefix_ids`` (B, P) / (B, C, P): task tokens,
    teacher-forced after the start token; the returned ids exclude the prefix (max_length generated tokens).

    ``cross_proj`` (bf16 multi-channel path): the ``nn.Linear`` of the `mc_shared_linear` pre_decoder.  Then
    ``encoder_hidden_states`` is the UNPROJECTED encoder output (B, T, K, d_latent) and the decoder runs its
    cross-attention in absorbed form on the latents (``ymt3_t5dec_generate_latent``)."""
    enc = encoder_hidden_states
    multi = decoder_type == "multi-t5"
    if cross_proj is not None:
        if not multi or precision != _lib.DTYPE_BF16:
            raise ValueError("cross_proj (absorbed cross-attention) is the bf16 multi-channel path")
        B, T, K, Dl = enc.shape
        Cn = K * Dl // cross_proj.in_features
        rt = getattr(decoder, "_runtime_latent", None)
        vocab = embed_tokens.weight.shape[0]
        n_prefix = 0 if prefix_ids is None else int(prefix_ids.shape[-1])
        if not (rt is not None and rt._cross_proj is cross_proj and rt.max_length >= max_length + n_prefix
                and rt.vocab_size == vocab and (rt.eos_id, rt.pad_id, rt.start_id) == (eos_id, pad_id, decoder_start_token_id)):
            if rt is not None:
                rt.free_native()
            rt = DecoderRuntime(decoder, embed_tokens, lm_head, precision, vocab, max_length + n_prefix,
                                getattr(lm_head, "tie_word_embeddings", True), eos_id, pad_id, decoder_start_token_id,
                                cross_proj=cross_proj)
            object.__setattr__(decoder, "_runtime_latent", rt)
            object.__setattr__(decoder, "_runtime", rt)
        toks = rt.generate_latent(enc.reshape(B, T, Cn, cross_proj.in_features), max_length, stop_at_eos,
                                  early_stop_interval, prefix_ids)
        return toks.long().view(B, Cn, max_length)
    if multi:
        B, Cn, T, D = enc.shape
        enc = enc.reshape(B * Cn, T, D)
    tie = getattr(lm_head, "tie_word_embeddings", True)
    vocab = embed_tokens.weight.shape[0]
    n_prefix = 0 if prefix_ids is None else int(prefix_ids.shape[-1])
    if prefix_ids is not None:
        prefix_ids = prefix_ids.reshape(enc.shape[0], -1)
    # `lanes` independent decode lanes (disjoint sequence ranges, each with its own native handle, CUDA graph and
    # internal stream) run concurrently on the GPU: the HBM-bound attention kernels of one lane overlap the
    # latency-bound small GEMMs / norms of the other.  Sequences are independent, so results are unchanged.
    lanes = max(1, min(int(lanes), enc.shape[0]))
    rts = getattr(decoder, "_runtimes", None) or []
    ok = (len(rts) >= lanes and all(rt.precision == precision and rt.max_length >= max_length + n_prefix
                                    and rt.vocab_size == vocab
                                    and (rt.eos_id, rt.pad_id, rt.start_id) == (eos_id, pad_id, decoder_start_token_id)
                                    for rt in rts[:lanes]))
    if not ok:
        for rt in rts:
            rt.free_native()
        rts = [DecoderRuntime(decoder, embed_tokens, lm_head, precision, vocab, max_length + n_prefix, tie, eos_id, pad_id,
                              decoder_start_token_id) for _ in range(lanes)]
        object.__setattr__(decoder, "_runtimes", rts)
        object.__setattr__(decoder, "_runtime", rts[0])
    N = enc.shape[0]
    per = -(-N // lanes)
    outs = []
    for li in range(lanes):
        a, b = li * per, min((li + 1) * per, N)
        if a >= b:
            break
        pfx = None if prefix_ids is None else prefix_ids[a:b]
        outs.append(rts[li].generate(enc[a:b], max_length, stop_at_eos, early_stop_interval, pfx))
    toks = (outs[0] if len(outs) == 1 else torch.cat(outs, 0)).long()
    return toks.view(B, Cn, max_length) if multi else toks
