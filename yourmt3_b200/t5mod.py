"""T5 encoder / decoder host modules with the reference's module tree and state-dict keys
(upstream amt/src/model/t5mod.py = modified copy of HF modeling_t5.py [RECALL]):

    block.{i}.layer.0.SelfAttention.{q,k,v,o}.weight      block.{i}.layer.0.layer_norm.weight
    block.{i}.layer.1.EncDecAttention.{q,k,v,o}.weight    (decoder)   ...layer_norm.weight
    block.{i}.layer.{1|2}.DenseReluDense.{wi_0,wi_1,wo}.weight        ...layer_norm.weight
    final_layer_norm.weight

The modules only OWN parameters; ``forward`` runs natively through the C ABI
(``ymt3_t5enc_forward`` / ``ymt3_t5dec_generate``): hand-written sm_100a kernels, no eager
fallback.  Position encoding: fixed absolute sinusoidal table added to ``inputs_embeds``
(upstream ``position_encoding_type='sinusoidal'`` [RECALL]) and / or T5's bucketed relative
attention bias (``position_encoding_type='relative'`` or ``has_relative_attention_bias=True``; HF
modeling_t5.py:189-268): block 0 owns ``relative_attention_bias.weight`` under its HF key and the
kernels add the per-distance table to the scores of every layer.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch
from torch import nn

from . import _lib


def sinusoidal_positions(n_pos: int, d_model: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """(n_pos, d_model) table, [sin | cos] halves [RECALL upstream positional_encoding.py]."""
    half = d_model // 2
    inc = math.log(max_timescale) / (half - 1)
    inv = torch.exp(-inc * torch.arange(half, dtype=torch.float64))
    t = torch.arange(n_pos, dtype=torch.float64)[:, None] * inv[None, :]
    return torch.cat([torch.sin(t), torch.cos(t)], dim=1).to(torch.float32)


def relative_position_bucket(rel: torch.Tensor, bidirectional: bool, num_buckets: int = 32, max_distance: int = 128):
    """Bucket of a relative position rel = key_pos - query_pos; the same tensor ops, in the same order and dtypes,
    as HF ``T5Attention._relative_position_bucket`` (modeling_t5.py:189-234), so the buckets are identical."""
    ret = torch.zeros_like(rel)
    if bidirectional:
        num_buckets //= 2
        ret = ret + (rel > 0).to(torch.long) * num_buckets
        rel = torch.abs(rel)
    else:
        rel = -torch.min(rel, torch.zeros_like(rel))
    max_exact = num_buckets // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact)
                         * (num_buckets - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return ret + torch.where(is_small, rel, large)


def relative_bias_by_distance(weight: torch.Tensor, n_pos: int, bidirectional: bool, max_distance: int = 128):
    """HF ``compute_bias`` (modeling_t5.py:248-268) depends on (j - i) only: fold the (num_buckets, H) embedding into
    a per-distance table for the kernels.  Encoder (bidirectional): (H, 2 n_pos - 1), entry n_pos - 1 + (j - i);
    decoder (causal): (H, n_pos), entry i - j >= 0."""
    dev = weight.device
    if bidirectional:
        rel = torch.arange(-(n_pos - 1), n_pos, dtype=torch.long, device=dev)
    else:
        rel = -torch.arange(0, n_pos, dtype=torch.long, device=dev)
    bucket = relative_position_bucket(rel, bidirectional, num_buckets=weight.shape[0], max_distance=max_distance)
    return weight.detach().float()[bucket].t().contiguous()            # (H, n)


class T5LayerNorm(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))


class T5Attention(nn.Module):
    def __init__(self, d_model, inner):
        super().__init__()
        self.q = nn.Linear(d_model, inner, bias=False)
        self.k = nn.Linear(d_model, inner, bias=False)
        self.v = nn.Linear(d_model, inner, bias=False)
        self.o = nn.Linear(inner, d_model, bias=False)


class T5DenseGatedActDense(nn.Module):
    def __init__(self, d_model, d_ff):
        super().__init__()
        self.wi_0 = nn.Linear(d_model, d_ff, bias=False)
        self.wi_1 = nn.Linear(d_model, d_ff, bias=False)
        self.wo = nn.Linear(d_ff, d_model, bias=False)


class T5LayerSelfAttention(nn.Module):
    def __init__(self, d_model, inner, rel_bias_heads: int = 0, num_buckets: int = 32):
        super().__init__()
        self.SelfAttention = T5Attention(d_model, inner)
        if rel_bias_heads:   # HF: only block 0 owns the table (has_relative_attention_bias), all layers share it
            self.SelfAttention.relative_attention_bias = nn.Embedding(num_buckets, rel_bias_heads)
        self.layer_norm = T5LayerNorm(d_model)


class T5LayerCrossAttention(nn.Module):
    def __init__(self, d_model, inner):
        super().__init__()
        self.EncDecAttention = T5Attention(d_model, inner)
        self.layer_norm = T5LayerNorm(d_model)


class T5LayerFF(nn.Module):
    def __init__(self, d_model, d_ff):
        super().__init__()
        self.DenseReluDense = T5DenseGatedActDense(d_model, d_ff)
        self.layer_norm = T5LayerNorm(d_model)


class T5Block(nn.Module):
    def __init__(self, d_model, inner, d_ff, is_decoder, rel_bias_heads: int = 0, num_buckets: int = 32):
        super().__init__()
        layers = [T5LayerSelfAttention(d_model, inner, rel_bias_heads, num_buckets)]
        if is_decoder:
            layers.append(T5LayerCrossAttention(d_model, inner))
        layers.append(T5LayerFF(d_model, d_ff))
        self.layer = nn.ModuleList(layers)


def _cfg_struct(cfg: Dict, precision: int, *, vocab=0, max_length=0, tie=True, eos=1, pad=0, start=0) -> _lib.T5Cfg:
    return _lib.T5Cfg(precision=precision, d_model=cfg["d_model"], num_heads=cfg["num_heads"], d_kv=cfg.get("d_kv", 64),
                      d_ff=cfg["d_model"] * cfg.get("ff_widening_factor", 2), num_layers=cfg["num_layers"],
                      layer_norm_eps=cfg.get("layer_norm_epsilon", 1e-6), vocab_size=vocab, max_length=max_length,
                      tie_word_embeddings=int(tie), eos_id=eos, pad_id=pad, start_id=start)


class _NativeOwner(nn.Module):
    """Caches a native handle keyed on the identity/version of the tensors it was packed from."""

    _destroy_name = ""

    def __init__(self):
        super().__init__()
        self._handle = None
        self._handle_key = None

    def _tensors(self) -> Dict[str, torch.Tensor]:
        raise NotImplementedError

    def _create(self, arr, n) -> C.c_void_p:
        raise NotImplementedError

    def native(self):
        named = self._tensors()
        key = tuple((k, v.data_ptr(), v._version, v.device.index) for k, v in named.items())
        if self._handle is None or key != self._handle_key:
            self.free_native()
            dev = next(iter(named.values())).device
            if dev.type != "cuda":
                raise RuntimeError("yourmt3_b200 modules run on CUDA only (no CPU fallback); call .cuda() first")
            arr, n, keep = _lib.tensor_table(named)
            with torch.cuda.device(dev):
                torch.cuda.current_stream().synchronize()
                self._handle = self._create(arr, n)
            del keep
            self._handle_key = key
        return self._handle

    def free_native(self):
        if self._handle is not None:
            getattr(_lib.load(), self._destroy_name)(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.free_native()
        except Exception:
            pass


class T5EncoderYMT3(_NativeOwner):
    """``forward(inputs_embeds=(B, T, d_model) f32) -> {"last_hidden_state": (B, T, d_model)}``."""

    _destroy_name = "ymt3_t5enc_destroy"

    def __init__(self, config: Dict, precision: str = "f32", num_max_positions: int = 1024):
        super().__init__()
        self.config = dict(config)
        self.precision = {"f32": _lib.DTYPE_F32, "bf16": _lib.DTYPE_BF16}[precision]
        d, inner = config["d_model"], config["num_heads"] * config.get("d_kv", 64)
        d_ff = d * config.get("ff_widening_factor", 2)
        pe = config.get("position_encoding_type", "sinusoidal")
        # 'relative' (or has_relative_attention_bias=True next to another type): HF T5's bucketed relative attention
        # bias, owned by block 0 under its HF state-dict key and shared by every layer
        self.has_relative_attention_bias = bool(config.get("has_relative_attention_bias", pe == "relative"))
        self.num_max_positions = num_max_positions
        nb = config.get("relative_attention_num_buckets", 32)
        self.block = nn.ModuleList([T5Block(d, inner, d_ff, False,
                                            config["num_heads"] if (i == 0 and self.has_relative_attention_bias) else 0, nb)
                                    for i in range(config["num_layers"])])
        self.final_layer_norm = T5LayerNorm(d)
        if pe == "sinusoidal":
            self.register_buffer("pos_table", sinusoidal_positions(num_max_positions, d), persistent=False)
        elif pe in (None, "none", "relative"):
            self.pos_table = None
        else:
            raise NotImplementedError(f"position_encoding_type={pe!r}")

    def _tensors(self):
        named = dict(self.named_parameters())
        if self.pos_table is not None:
            named["pos_table"] = self.pos_table
        if self.has_relative_attention_bias:
            w = named.pop("block.0.layer.0.SelfAttention.relative_attention_bias.weight")
            named["relative_bias_by_distance"] = relative_bias_by_distance(
                w, self.num_max_positions, True, self.config.get("relative_attention_max_distance", 128))
        return named

    def _create(self, arr, n):
        h = C.c_void_p()
        cfg = _cfg_struct(self.config, self.precision)
        _lib.check(_lib.load().ymt3_t5enc_create(C.byref(cfg), arr, n, C.byref(h)), "t5enc_create")
        return h

    def forward(self, inputs_embeds: torch.Tensor, **unused):
        x = inputs_embeds
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("T5EncoderYMT3 expects a float32 CUDA tensor (no CPU fallback)")
        x = x.contiguous()
        B, S, D = x.shape
        h = self.native()
        out = torch.empty((B, S, D), dtype=_lib.torch_dtype(self.precision), device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().ymt3_t5enc_forward(h, x.data_ptr(), B, S, out.data_ptr(), _lib.current_stream_ptr()),
                       "t5enc_forward")
        return {"last_hidden_state": out}


class T5DecoderYMT3(nn.Module):
    """Parameter container of the decoder stack; generation runs through
    :func:`yourmt3_b200.t5mod_helper.task_cond_dec_generate` (native, device-resident loop)."""

    def __init__(self, config: Dict, num_max_positions: int = 1024):
        super().__init__()
        self.config = dict(config)
        d, inner = config["d_model"], config["num_heads"] * config.get("d_kv", 64)
        d_ff = d * config.get("ff_widening_factor", 2)
        pe = config.get("position_encoding_type", "sinusoidal")
        self.has_relative_attention_bias = bool(config.get("has_relative_attention_bias", pe == "relative"))
        self.num_max_positions = num_max_positions
        nb = config.get("relative_attention_num_buckets", 32)
        self.block = nn.ModuleList([T5Block(d, inner, d_ff, True,
                                            config["num_heads"] if (i == 0 and self.has_relative_attention_bias) else 0, nb)
                                    for i in range(config["num_layers"])])
        self.final_layer_norm = T5LayerNorm(d)
        if pe == "sinusoidal":
            self.register_buffer("pos_table", sinusoidal_positions(num_max_positions, d), persistent=False)
        elif pe in (None, "none", "relative"):
            self.pos_table = None
        else:
            raise NotImplementedError(f"position_encoding_type={pe!r}")

    def forward(self, *a, **k):
        raise NotImplementedError("teacher-forced decoder forward is a training path (out of scope); "
                                  "use task_cond_dec_generate / YourMT3.inference")


class MultiChannelT5Decoder(T5DecoderYMT3):
    """Multi-channel decoder: ``num_channels`` independent sequences per segment share the decoder
    weights; channels are folded into the batch ((B, C, T, D) -> (B*C, T, D)) [RECALL upstream]."""

    def __init__(self, config: Dict, num_max_positions: int = 1024):
        super().__init__(config, num_max_positions)
        self.num_channels = config.get("num_channels", 13)
