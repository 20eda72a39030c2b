"""Deterministic, NON-DEGENERATE random initialisation for benchmarks and parity tests.

HF-default init makes greedy decoding emit a constant token, which would make
token-identity checks vacuous.  This init (N(0, 0.05) on matrices, 1 + 0.1 N(0,1) on norm
scales, N(0, 0.2) embeddings) gives varied tokens with a median top-1/top-2 logit margin of
~1e-1, so identical-token comparisons are meaningful and robust to fp32 summation order."""
import torch
from torch import nn


@torch.no_grad()
def init_nondegenerate_(module: nn.Module, seed: int = 0, std: float = 0.05, embed_std: float = 0.2) -> nn.Module:
    g = torch.Generator().manual_seed(seed)
    seen = set()
    for name, p in sorted(module.named_parameters(), key=lambda kv: kv[0]):
        if p.data_ptr() in seen:
            continue
        seen.add(p.data_ptr())
        if "embed_tokens" in name or "latent_array" in name:
            v = torch.randn(p.shape, generator=g) * embed_std
        elif p.dim() >= 2:
            fan_in = p[0].numel()
            # very wide layers (e.g. the K*D -> d_model 'linear' pre-decoder) are scaled down so that they do not
            # swamp the residual stream and collapse the decode to a constant token
            scale = std * min(1.0, (1024.0 / fan_in) ** 0.5)
            v = torch.randn(p.shape, generator=g) * scale
        elif name.endswith("bias"):
            v = torch.randn(p.shape, generator=g) * 0.02
        else:  # norm scales
            v = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
        p.copy_(v.to(p.device, p.dtype))
    for name, b in sorted(module.named_buffers(), key=lambda kv: kv[0]):
        if name.endswith("running_mean"):       # BatchNorm statistics: non-trivial so that folding is exercised
            b.copy_((torch.randn(b.shape, generator=g) * 0.1).to(b.device, b.dtype))
        elif name.endswith("running_var"):
            b.copy_((0.5 + torch.rand(b.shape, generator=g)).to(b.device, b.dtype))
    return module
