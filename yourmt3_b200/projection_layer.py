"""Encoder -> decoder projections (upstream amt/src/model/projection_layer.py [RECALL]):

* ``linear``:            (B, T, K, D) -> 'b t (k d)' -> Linear(K*D, d_dec)             -> (B, T, d_dec)
* ``mc_shared_linear``:  (B, T, K, D) -> 'b t (c k2) d -> b c t (k2 d)' -> one Linear(k2*D, d_dec)
                         shared by all C channels                                    -> (B, C, T, d_dec)

Native: one GEMM with fused bias on the latent buffer viewed as rows of k2*D (no gather needed, the
latents of a channel are contiguous) and, for the multi-channel case, one layout permute."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib


class _Projection(nn.Module):
    def __init__(self, kind, num_latents, d_latent, d_dec, num_channels, precision):
        super().__init__()
        self.kind, self.K, self.D, self.d_dec, self.C = kind, num_latents, d_latent, d_dec, num_channels
        self.precision = {"f32": _lib.DTYPE_F32, "bf16": _lib.DTYPE_BF16}[precision]
        if kind == "mc_shared_linear":
            if num_latents % num_channels:
                raise ValueError("num_latents must be a multiple of num_channels")
            in_f = (num_latents // num_channels) * d_latent
        elif kind == "linear":
            in_f = num_latents * d_latent
        else:
            raise NotImplementedError(f"pre_decoder type {kind!r}")
        self.proj = nn.Linear(in_f, d_dec)
        self._w_cache = None

    def _weight(self):
        w = self.proj.weight
        key = (w.data_ptr(), w._version, self.precision)
        if self._w_cache is None or self._w_cache[0] != key:
            self._w_cache = (key, w.detach().to(_lib.torch_dtype(self.precision)).contiguous(),
                             self.proj.bias.detach().float().contiguous())
        return self._w_cache[1], self._w_cache[2]

    def forward(self, h: torch.Tensor) -> torch.Tensor:
        want = _lib.torch_dtype(self.precision)
        if not h.is_cuda:
            raise RuntimeError("projection runs on CUDA only (no CPU fallback)")
        if h.dtype != want:
            h = h.to(want)
        h = h.contiguous()
        B, T, K, D = h.shape
        W, b = self._weight()
        in_f = W.shape[1]
        rows = B * T * (K * D // in_f)
        y = torch.empty((rows, self.d_dec), dtype=want, device=h.device)
        lib, s = _lib.load(), _lib.current_stream_ptr()
        with torch.cuda.device(h.device):
            _lib.check(lib.ymt3_op_linear(self.precision, h.data_ptr(), in_f, W.data_ptr(), in_f, b.data_ptr(),
                                          y.data_ptr(), self.d_dec, None, 0, rows, self.d_dec, in_f, 0, 0, 1.0,
                                          self.precision, s), "projection")
            if self.kind == "linear":
                return y.view(B, T, self.d_dec)
            out = torch.empty((B, self.C, T, self.d_dec), dtype=want, device=h.device)
            _lib.check(lib.ymt3_op_permute_btcd_bctd(self.precision, y.data_ptr(), out.data_ptr(), B, T, self.C,
                                                     self.d_dec, s), "permute")
        return out


def get_projection_layer(kind, num_latents, d_latent, d_dec, num_channels, precision):
    if kind is None:
        return nn.Identity()
    return _Projection(kind, num_latents, d_latent, d_dec, num_channels, precision)
