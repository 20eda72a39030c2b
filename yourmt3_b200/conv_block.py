"""Residual-conv pre-encoder of the Perceiver-TF models (upstream amt/src/model/conv_block.py
``PreEncoderBlockRes3B`` [RECALL]): ``(B, T, F) -> (B, T, F/8, C)``.

Each of the 3 blocks: ``h = conv2(relu(bn2(conv1(relu(bn1(x))))))``, ``h += shortcut(x)`` (1x1 conv when the
channel count changes), ``AvgPool2d((1, 2))`` over frequency.  BatchNorm runs in eval mode.
The module owns the parameters (torch layers, same state-dict keys as nn.Conv2d / nn.BatchNorm2d);
``forward`` is native: first conv on CUDA cores, the others as implicit GEMMs on tcgen05 (bf16) or
im2col + fp32 FFMA GEMM (f32)."""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib
from .t5mod import _NativeOwner


class Res2DAVPBlock(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1, bias=False)
        if in_channels != out_channels:
            self.shortcut = nn.Conv2d(in_channels, out_channels, 1)


class PreEncoderBlockRes3B(_NativeOwner):
    _destroy_name = "ymt3_res3b_destroy"

    def __init__(self, t_feat_len: int, f_feat_len: int, channels=(64, 128, 128), precision: str = "f32"):
        super().__init__()
        self.t_feat_len, self.f_feat_len, self.channels = t_feat_len, f_feat_len, tuple(channels)
        self.precision = {"f32": _lib.DTYPE_F32, "bf16": _lib.DTYPE_BF16}[precision]
        cin, blocks = 1, []
        for c in self.channels:
            blocks.append(Res2DAVPBlock(cin, c))
            cin = c
        self.blocks = nn.ModuleList(blocks)
        self.output_shape = (t_feat_len, f_feat_len // 8, self.channels[-1])

    def _tensors(self):
        named = dict(self.named_parameters())
        named.update({k: v for k, v in self.named_buffers() if "num_batches_tracked" not in k})
        return named

    def _create(self, arr, n):
        h = C.c_void_p()
        cfg = _lib.Res3bCfg(precision=self.precision, in_freq=self.f_feat_len,
                            channels=(C.c_int32 * 3)(*self.channels), bn_eps=self.blocks[0].bn1.eps)
        _lib.check(_lib.load().ymt3_res3b_create(C.byref(cfg), arr, n, C.byref(h)), "res3b_create")
        return h

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: (B, T, F) float32 CUDA -> (B, T, F/8, C) in the module precision."""
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("PreEncoderBlockRes3B expects a float32 CUDA tensor (no CPU fallback)")
        x = x.contiguous()
        B, T, F = x.shape
        if F != self.f_feat_len:
            raise ValueError(f"expected {self.f_feat_len} frequency bins, got {F}")
        h = self.native()
        out = torch.empty((B, T, F // 8, self.channels[-1]), dtype=_lib.torch_dtype(self.precision), device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().ymt3_res3b_forward(h, x.data_ptr(), B, T, out.data_ptr(), _lib.current_stream_ptr()),
                       "res3b_forward")
        return out
