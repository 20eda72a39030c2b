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


class DecoderRuntime(_NativeOwner):
    """Native decoder handle built from (decoder, embed_tokens, lm_head) parameters."""

    _destroy_name = "ymt3_t5dec_destroy"

    def __init__(self, decoder: T5DecoderYMT3, embed_tokens, lm_head, precision: int, vocab_size: int,
                 max_length: int, tie_word_embeddings: bool, eos_id=1, pad_id=0, start_id=0):
        super().__init__()
        # plain attribute refs (not registered as submodules: they are owned by YourMT3)
        object.__setattr__(self, "_decoder", decoder)
        object.__setattr__(self, "_embed", embed_tokens)
        object.__setattr__(self, "_lm", lm_head)
        self.precision, self.vocab_size, self.max_length = precision, vocab_size, max_length
        self.tie, self.eos_id, self.pad_id, self.start_id = tie_word_embeddings, eos_id, pad_id, start_id

    def _tensors(self):
        named = dict(self._decoder.named_parameters())
        if getattr(self._decoder, "pos_table", None) is not None:
            named["pos_table"] = self._decoder.pos_table
        named["embed_tokens.weight"] = self._embed.weight
        lm_w = self._lm.lm_head.weight if hasattr(self._lm, "lm_head") else self._lm.weight
        if not self.tie or lm_w.data_ptr() != self._embed.weight.data_ptr():
            named["lm_head.weight"] = lm_w
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

    def last_logits(self, N: int, device) -> torch.Tensor:
        out = torch.empty((N, self.vocab_size), dtype=torch.float32, device=device)
        _lib.check(_lib.load().ymt3_t5dec_last_logits(self.native(), out.data_ptr(), N, _lib.current_stream_ptr()),
                   "t5dec_last_logits")
        return out


def task_cond_dec_generate(decoder, decoder_type: str, embed_tokens, lm_head, encoder_hidden_states: torch.Tensor,
                           shift_right_fn=None, prefix_ids: Optional[torch.Tensor] = None, max_length: int = 1024,
                           stop_at_eos: bool = True, eos_id: int = 1, pad_id: int = 0,
                           decoder_start_token_id: int = 0, precision: int = _lib.DTYPE_F32,
                           early_stop_interval: int = 0, lanes: int = 1, **unused) -> torch.Tensor:
    """Returns LongTensor (B, max_length) for 't5' or (B, C, max_length) for 'multi-t5'.

    ``encoder_hidden_states``: (B, T, D) or (B, C, T, D).  ``prefix_ids`` (B, P) / (B, C, P): task tokens,
    teacher-forced after the start token; the returned ids exclude the prefix (max_length generated tokens)."""
    enc = encoder_hidden_states
    multi = decoder_type == "multi-t5"
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
