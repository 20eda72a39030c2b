"""Perceiver-TF encoder host module (upstream amt/src/model/perceiver_mod.py [RECALL]; layer
naming follows HF ``PerceiverLayer``: attention.self.{layernorm1,layernorm2,query,key,value},
attention.output.dense, layernorm, mlp.{dense1,dense2} | moe.{gate,experts.N.{w1,w2,w3}}).

``forward(inputs_embeds=(B, T, F', C)) -> {"last_hidden_state": (B, T, K, D)}`` runs natively
(``ymt3_ptf_forward``).  Also builds the (pre_encoder, encoder, pre_decoder) triple for YourMT3."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch
from torch import nn

from . import _lib
from .ff_layer import MoE, PerceiverMLP
from .t5mod import _NativeOwner


class _Norm(nn.Module):
    def __init__(self, d, kind):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        if kind == "layer_norm":
            self.bias = nn.Parameter(torch.zeros(d))


class _SelfAttention(nn.Module):
    def __init__(self, d, kv_dim, cross, kind):
        super().__init__()
        self.layernorm1 = _Norm(d, kind)
        if cross:
            self.layernorm2 = _Norm(kv_dim, kind)
        self.query = nn.Linear(d, d)
        self.key = nn.Linear(kv_dim, d)
        self.value = nn.Linear(kv_dim, d)


class _SelfOutput(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.dense = nn.Linear(d, d)


class _Attention(nn.Module):
    def __init__(self, d, kv_dim, cross, kind):
        super().__init__()
        self.self = _SelfAttention(d, kv_dim, cross, kind)
        self.output = _SelfOutput(d)


class PerceiverLayer(nn.Module):
    def __init__(self, cfg: Dict, d: int, kv_dim: int, cross: bool):
        super().__init__()
        kind = cfg.get("layer_norm_type", "layer_norm")
        self.attention = _Attention(d, kv_dim, cross, kind)
        self.layernorm = _Norm(d, kind)
        if cfg.get("ff_layer_type", "mlp") == "moe":
            self.moe = MoE(d, cfg["ff_widening_factor"], cfg["moe_num_experts"], cfg["moe_topk"])
        else:
            self.mlp = PerceiverMLP(d, cfg["ff_widening_factor"])


class PerceiverTFBlock(nn.Module):
    def __init__(self, cfg, d, kv_dim):
        super().__init__()
        self.sca = PerceiverLayer(cfg, d, kv_dim, True)
        self.local = nn.ModuleList([PerceiverLayer(cfg, d, d, False)
                                    for _ in range(cfg["num_local_transformers_per_block"])])
        self.temporal = nn.ModuleList([PerceiverLayer(cfg, d, d, False)
                                       for _ in range(cfg["num_temporal_transformers_per_block"])])


class _Latents(nn.Module):
    def __init__(self, k, d):
        super().__init__()
        self.latents = nn.Parameter(torch.randn(k, d))


class PerceiverTFEncoder(_NativeOwner):
    _destroy_name = "ymt3_ptf_destroy"

    def __init__(self, config: Dict, kv_dim: int, max_time: int, precision: str = "f32"):
        super().__init__()
        self.config = cfg = dict(config)
        self.precision = {"f32": _lib.DTYPE_F32, "bf16": _lib.DTYPE_BF16}[precision]
        self.kv_dim, self.max_time = kv_dim, max_time
        K, D = cfg["num_latents"], cfg["d_latent"]
        self.latent_array = _Latents(K, D)
        pe = cfg.get("position_encoding_type", "trainable")
        if pe == "trainable":
            self.latent_pos_emb = nn.Parameter(torch.zeros(K, D))
            self.temporal_pos_emb = nn.Parameter(torch.zeros(max_time, D))
        elif pe not in ("rope", None, "none"):
            raise NotImplementedError(f"position_encoding_type={pe!r}")
        self.block = nn.ModuleList([PerceiverTFBlock(cfg, D, kv_dim) for _ in range(cfg["num_blocks"])])
        self.layernorm = _Norm(D, cfg.get("layer_norm_type", "layer_norm"))

    def _tensors(self):
        return dict(self.named_parameters())

    def _create(self, arr, n):
        cfg = self.config
        D, hs = cfg["d_latent"], cfg.get("num_self_attention_heads", 8)
        pe = cfg.get("position_encoding_type", "trainable")
        rope = 0
        if pe == "rope":
            rope = (D // hs) // 2 if cfg.get("rotary_partial_pe", False) else D // hs
        c = _lib.PtfCfg(
            precision=self.precision, num_latents=cfg["num_latents"], d_latent=D, kv_dim=self.kv_dim,
            num_blocks=cfg["num_blocks"], num_local=cfg["num_local_transformers_per_block"],
            num_temporal=cfg["num_temporal_transformers_per_block"],
            cross_heads=cfg.get("num_cross_attention_heads", 1), self_heads=hs,
            sca_query_residual=int(cfg.get("sca_use_query_residual", False)),
            norm_type=1 if cfg.get("layer_norm_type", "layer_norm") == "rms_norm" else 0,
            ff_type=1 if cfg.get("ff_layer_type", "mlp") == "moe" else 0, ff_widening=cfg["ff_widening_factor"],
            moe_experts=cfg.get("moe_num_experts", 0), moe_topk=cfg.get("moe_topk", 0),
            act=_lib.ACT_CODES[cfg.get("hidden_act", "gelu")],
            pos_type={"trainable": 1, "rope": 2}.get(pe, 0), rope_dim=rope, max_time=self.max_time,
            norm_eps=cfg.get("layer_norm_eps", 1e-5))
        h = C.c_void_p()
        _lib.check(_lib.load().ymt3_ptf_create(C.byref(c), arr, n, C.byref(h)), "ptf_create")
        return h

    def forward(self, inputs_embeds: torch.Tensor, **unused):
        x = inputs_embeds
        want = _lib.torch_dtype(self.precision)
        if not x.is_cuda:
            raise RuntimeError("PerceiverTFEncoder runs on CUDA only (no CPU fallback)")
        if x.dtype != want:
            x = x.to(want)
        x = x.contiguous()
        B, T, Fp, Cc = x.shape
        if Cc != self.kv_dim:
            raise ValueError(f"expected {self.kv_dim} channels, got {Cc}")
        h = self.native()
        out = torch.empty((B, T, self.config["num_latents"], self.config["d_latent"]), dtype=want, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().ymt3_ptf_forward(h, x.data_ptr(), B, T, Fp, out.data_ptr(), _lib.current_stream_ptr()),
                       "ptf_forward")
        return {"last_hidden_state": out}


def build_perceiver_tf_stages(owner, audio_cfg, model_cfg, pre_enc, pre_dec, precision):
    """(pre_encoder, encoder, pre_decoder) for encoder_type == 'perceiver-tf'."""
    from .conv_block import PreEncoderBlockRes3B
    from .projection_layer import get_projection_layer
    cfg = model_cfg["encoder"]["perceiver-tf"]
    T, F = owner.feat_length, owner.feat_dim
    if pre_enc != "conv":
        raise NotImplementedError(f"pre_encoder type {pre_enc!r} (only 'conv' = res3b is on the benchmarked path)")
    C_out = model_cfg.get("conv_out_channels", 128)
    pre_encoder = PreEncoderBlockRes3B(T, F, channels=(64, C_out, C_out), precision=precision)
    encoder = PerceiverTFEncoder(cfg, kv_dim=C_out, max_time=T, precision=precision)
    dec_cfg = model_cfg["decoder"][model_cfg["decoder_type"]]
    pre_decoder = get_projection_layer(pre_dec, cfg["num_latents"], cfg["d_latent"], dec_cfg["d_model"],
                                       dec_cfg.get("num_channels", 1), precision)
    return pre_encoder, encoder, pre_decoder
