"""Plain-PyTorch (CPU, fp32) restatement of the MT3-style T5 encoder / decoder and the
greedy segment-batched decode loop (oracle; test infra only).

Follows the installed HF transformers 5.5.0 T5 blocks that upstream ``amt/src/model/t5mod.py``
copies (SP = site-packages/transformers/models/t5/modeling_t5.py):

* T5LayerNorm (RMSNorm, fp32 variance, eps 1e-6, no bias)            SP:46-68
* T5DenseGatedActDense  wo(gelu_new(wi_0 x) * wi_1 x)                SP:106-131, activations.py:59-66
* T5Attention: q,k,v,o without bias, NO 1/sqrt(d) scaling, fp32 softmax  SP:153-344 (scores :308, softmax :331)
* T5LayerSelfAttention / T5LayerCrossAttention / T5LayerFF pre-norm residual  SP:347-408, 134-150
* T5Stack: blocks then final_layer_norm                              SP:617-792
* KV cache semantics (self K/V appended per step, cross K/V computed once)  SP:269-305
* logits * d_model**-0.5 when embeddings are tied                    SP:1105-1110

* relative attention bias (bucketed, learned per head in block 0, shared by all blocks)  SP:200-268, 625, 755-760

Position handling is SWITCHABLE: when the state dict holds
``block.0.layer.0.SelfAttention.relative_attention_bias.weight`` the HF relative bias is added to
every self-attention score (bidirectional buckets in the encoder, unidirectional in the decoder);
otherwise no bias is used (upstream's absolute-position variant).  Both are pinned against HF.

Upstream specifics restated from memory of mimbres/YourMT3 ([RECALL], unverifiable here):
absolute sinusoidal position encoding added to ``inputs_embeds`` (no relative attention
bias by default), decoder start token 0 (= pad), EOS = 1, rows that emitted EOS are padded with 0,
the loop stops when every row is finished or ``max_length`` is reached
(``t5mod_helper.task_cond_dec_generate``).

State-dict keys are HF's: ``block.{i}.layer.{j}.SelfAttention.{q,k,v,o}.weight`` etc.
Pinned against ``transformers.models.t5.modeling_t5.T5Stack`` in tests/test_oracle_t5.py.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

Tensor = torch.Tensor


def sinusoidal_positions(n_pos: int, d_model: int, max_timescale: float = 10000.0) -> Tensor:
    """Fixed absolute position table (n_pos, d_model): [sin | cos] halves (whisper-style) [RECALL]."""
    half = d_model // 2
    inc = math.log(max_timescale) / (half - 1)
    inv = torch.exp(-inc * torch.arange(half, dtype=torch.float64))
    t = torch.arange(n_pos, dtype=torch.float64)[:, None] * inv[None, :]
    return torch.cat([torch.sin(t), torch.cos(t)], dim=1).to(torch.float32)


def rms_norm(x: Tensor, w: Tensor, eps: float = 1e-6) -> Tensor:
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def gelu_new(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def _heads(x: Tensor, n_heads: int) -> Tensor:  # (B, S, H*dk) -> (B, H, S, dk)
    B, S, _ = x.shape
    return x.view(B, S, n_heads, -1).transpose(1, 2)


def attention(q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor] = None, scale: float = 1.0) -> Tensor:
    """q (B,H,Sq,dk), k/v (B,H,Sk,dk) -> (B,Sq,H*dk). fp32 softmax; T5 uses scale=1."""
    scores = torch.matmul(q, k.transpose(-1, -2)) * scale
    if mask is not None:
        scores = scores + mask
    p = torch.softmax(scores.float(), dim=-1).type_as(scores)
    o = torch.matmul(p, v)
    B, H, Sq, dk = o.shape
    return o.transpose(1, 2).reshape(B, Sq, H * dk)


def relative_position_bucket(rel: Tensor, bidirectional: bool, num_buckets: int = 32, max_distance: int = 128) -> Tensor:
    """HF ``T5Attention._relative_position_bucket`` (SP:200-246): rel = key_pos - query_pos (int64) -> bucket id."""
    ret = torch.zeros_like(rel)
    if bidirectional:
        num_buckets //= 2
        ret = ret + (rel > 0).to(torch.long) * num_buckets
        rel = rel.abs()
    else:
        rel = -torch.min(rel, torch.zeros_like(rel))
    max_exact = num_buckets // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact)
                         * (num_buckets - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return ret + torch.where(is_small, rel, large)


def relative_bias(sd, prefix: str, q_pos: Tensor, k_len: int, bidirectional: bool, max_distance: int = 128) -> Optional[Tensor]:
    """HF ``compute_bias`` (SP:248-268) for query positions q_pos (Sq,) over keys 0..k_len-1 -> (1, H, Sq, k_len),
    or None when the state dict carries no relative attention bias (absolute-position variant)."""
    w = sd.get(prefix + "block.0.layer.0.SelfAttention.relative_attention_bias.weight")
    if w is None:
        return None
    rel = torch.arange(k_len, dtype=torch.long, device=w.device)[None, :] - q_pos.to(torch.long)[:, None]
    bucket = relative_position_bucket(rel, bidirectional, num_buckets=w.shape[0], max_distance=max_distance)
    return w[bucket].permute(2, 0, 1)[None]           # (Sq, Sk, H) -> (1, H, Sq, Sk)


def t5_self_attention_layer(sd, pre, x, n_heads, eps, mask=None):
    h = rms_norm(x, sd[pre + "layer_norm.weight"], eps)
    a = pre + "SelfAttention."
    q = _heads(h @ sd[a + "q.weight"].T, n_heads)
    k = _heads(h @ sd[a + "k.weight"].T, n_heads)
    v = _heads(h @ sd[a + "v.weight"].T, n_heads)
    return x + attention(q, k, v, mask) @ sd[a + "o.weight"].T, (k, v)


def t5_ff_layer(sd, pre, x, eps):
    h = rms_norm(x, sd[pre + "layer_norm.weight"], eps)
    d = pre + "DenseReluDense."
    g = gelu_new(h @ sd[d + "wi_0.weight"].T) * (h @ sd[d + "wi_1.weight"].T)
    return x + g @ sd[d + "wo.weight"].T


def t5_encoder(sd: Dict[str, Tensor], x: Tensor, *, n_layers: int, n_heads: int, eps: float = 1e-6,
               prefix: str = "", pos: Optional[Tensor] = None) -> Tensor:
    """inputs_embeds (B, T, d_model) -> last_hidden_state (B, T, d_model)."""
    if pos is not None:
        x = x + pos[: x.shape[1]]
    T = x.shape[1]
    bias = relative_bias(sd, prefix, torch.arange(T), T, bidirectional=True)
    for i in range(n_layers):
        b = f"{prefix}block.{i}.layer."
        x, _ = t5_self_attention_layer(sd, b + "0.", x, n_heads, eps, mask=bias)
        x = t5_ff_layer(sd, b + "1.", x, eps)
    return rms_norm(x, sd[prefix + "final_layer_norm.weight"], eps)


def t5_decoder_full(sd, x, enc_hs, *, n_layers, n_heads, eps=1e-6, prefix="", pos=None):
    """Teacher-forced full-sequence decoder: inputs_embeds (B, S, d) with causal mask."""
    S = x.shape[1]
    if pos is not None:
        x = x + pos[:S]
    causal = torch.full((S, S), float("-inf"), device=x.device).triu(1)
    bias = relative_bias(sd, prefix, torch.arange(S), S, bidirectional=False)
    if bias is not None:
        causal = causal + bias
    for i in range(n_layers):
        b = f"{prefix}block.{i}.layer."
        x, _ = t5_self_attention_layer(sd, b + "0.", x, n_heads, eps, mask=causal)
        h = rms_norm(x, sd[b + "1.layer_norm.weight"], eps)
        a = b + "1.EncDecAttention."
        q = _heads(h @ sd[a + "q.weight"].T, n_heads)
        k = _heads(enc_hs @ sd[a + "k.weight"].T, n_heads)
        v = _heads(enc_hs @ sd[a + "v.weight"].T, n_heads)
        x = x + attention(q, k, v) @ sd[a + "o.weight"].T
        x = t5_ff_layer(sd, b + "2.", x, eps)
    return rms_norm(x, sd[prefix + "final_layer_norm.weight"], eps)


class T5DecoderState:
    """Device-agnostic incremental decoder (KV cache) used by greedy_generate."""

    def __init__(self, sd, enc_hs, *, n_layers, n_heads, eps=1e-6, prefix="", pos=None):
        self.sd, self.n_layers, self.n_heads, self.eps, self.prefix, self.pos = sd, n_layers, n_heads, eps, prefix, pos
        self.self_k = [None] * n_layers
        self.self_v = [None] * n_layers
        self.cross = []
        for i in range(n_layers):
            a = f"{prefix}block.{i}.layer.1.EncDecAttention."
            self.cross.append((_heads(enc_hs @ sd[a + "k.weight"].T, n_heads),
                               _heads(enc_hs @ sd[a + "v.weight"].T, n_heads)))
        self.len = 0

    def step(self, x: Tensor) -> Tensor:
        """x: (N, 1, d_model) embedded token at position self.len -> (N, 1, d_model) hidden."""
        sd, H, eps = self.sd, self.n_heads, self.eps
        if self.pos is not None:
            x = x + self.pos[self.len: self.len + 1]
        # relative bias of the single query at position len over keys 0..len (SP:296-303 with a cache)
        bias = relative_bias(sd, self.prefix, torch.tensor([self.len]), self.len + 1, bidirectional=False)
        for i in range(self.n_layers):
            b = f"{self.prefix}block.{i}.layer."
            h = rms_norm(x, sd[b + "0.layer_norm.weight"], eps)
            a = b + "0.SelfAttention."
            q = _heads(h @ sd[a + "q.weight"].T, H)
            k = _heads(h @ sd[a + "k.weight"].T, H)
            v = _heads(h @ sd[a + "v.weight"].T, H)
            if self.self_k[i] is not None:
                k = torch.cat([self.self_k[i], k], dim=2)
                v = torch.cat([self.self_v[i], v], dim=2)
            self.self_k[i], self.self_v[i] = k, v
            x = x + attention(q, k, v, bias) @ sd[a + "o.weight"].T
            h = rms_norm(x, sd[b + "1.layer_norm.weight"], eps)
            a = b + "1.EncDecAttention."
            q = _heads(h @ sd[a + "q.weight"].T, H)
            x = x + attention(q, *self.cross[i]) @ sd[a + "o.weight"].T
            x = t5_ff_layer(sd, b + "2.", x, eps)
        self.len += 1
        return rms_norm(x, sd[self.prefix + "final_layer_norm.weight"], eps)


def greedy_generate(sd, enc_hs: Tensor, *, embed: Tensor, lm_head: Tensor, n_layers: int, n_heads: int,
                    max_length: int, eps: float = 1e-6, prefix: str = "decoder.", pos: Optional[Tensor] = None,
                    tie_word_embeddings: bool = True, eos_id: int = 1, pad_id: int = 0, start_id: int = 0,
                    stop_at_eos: bool = True, return_margins: bool = False, prefix_ids: Optional[Tensor] = None):
    """task_cond_dec_generate semantics (no task prefix): returns (N, max_length) int64 tokens.

    enc_hs: (N, T_enc, d_model) -- for the multi-channel decoder N = B*C (channels folded
    into the batch, shared decoder weights).  Rows that emitted EOS keep emitting pad;
    decoding stops early when all rows are finished, the tail stays pad.
    """
    N, _, d_model = enc_hs.shape
    st = T5DecoderState(sd, enc_hs, n_layers=n_layers, n_heads=n_heads, eps=eps, prefix=prefix, pos=pos)
    dev = enc_hs.device
    out = torch.full((N, max_length), pad_id, dtype=torch.long, device=dev)
    margins = torch.full((N, max_length), float("inf"), device=dev)
    cur = torch.full((N,), start_id, dtype=torch.long, device=dev)
    unfinished = torch.ones(N, dtype=torch.bool, device=dev)
    if prefix_ids is not None and prefix_ids.numel() > 0:
        # task prefix: decoder inputs are [start, prefix..., generated...]; the prefix is teacher-forced
        pfx = prefix_ids.reshape(N, -1).to(dev)
        for s_ in range(pfx.shape[1]):
            st.step(embed[cur][:, None, :])
            cur = pfx[:, s_]
    for t in range(max_length):
        hs = st.step(embed[cur][:, None, :])[:, 0]
        if tie_word_embeddings:
            hs = hs * (d_model ** -0.5)
        logits = hs @ lm_head.T
        nxt = logits.argmax(-1)
        if return_margins:
            top2 = logits.topk(2, dim=-1).values
            margins[:, t] = torch.where(unfinished, top2[:, 0] - top2[:, 1], margins[:, t])
        nxt = torch.where(unfinished, nxt, torch.full_like(nxt, pad_id))
        out[:, t] = nxt
        if stop_at_eos:
            unfinished = unfinished & (nxt != eos_id)
            if not bool(unfinished.any()):
                break
        cur = nxt
    return (out, margins) if return_margins else out
