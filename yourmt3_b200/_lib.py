"""ctypes loader for the C-ABI library (include/ymt3_b200.h).

There is deliberately no fallback: if ``libymt3_b200.so`` is missing or a call
fails, a RuntimeError is raised (north-star: no CPU fallback, no backend dispatch).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# YMT3_B200_LIB: a differently-built copy of the SAME library (debug builds, e.g. the timeline build of tools/trace_gemm.py)
LIB_PATH = os.environ.get("YMT3_B200_LIB") or os.path.join(_HERE, "csrc", "libymt3_b200.so")

_lib = None
_lock = threading.Lock()


class AudioCfg(C.Structure):
    """Mirror of ``ymt3_audio_cfg_t``."""

    _fields_ = [
        ("n_fft", C.c_int32),
        ("hop_length", C.c_int32),
        ("codec", C.c_int32),
        ("n_mels", C.c_int32),
        ("spec_bin0", C.c_int32),
        ("spec_bins", C.c_int32),
        ("power_mode", C.c_int32),
        ("log_eps", C.c_float),
    ]


class Tensor(C.Structure):
    """Mirror of ``ymt3_tensor_t`` (named device tensor = state-dict entry)."""

    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 4)]


class ChainPhase(C.Structure):
    """ymt3_chain_phase_t (include/ymt3_b200.h)"""
    _fields_ = [("A", C.c_void_p), ("lda", C.c_int64), ("W", C.c_void_p), ("ldw", C.c_int64), ("bias", C.c_void_p),
                ("ss_in", C.c_void_p), ("chunks", C.c_int64), ("eps", C.c_float), ("C", C.c_void_p), ("ldc", C.c_int64),
                ("residual", C.c_void_p), ("ldr", C.c_int64), ("ss_out", C.c_void_p), ("N", C.c_int64), ("K", C.c_int64),
                ("act", C.c_int32), ("gated", C.c_int32), ("out_scale", C.c_float)]


class T5Cfg(C.Structure):
    """Mirror of ``ymt3_t5_cfg_t``."""

    _fields_ = [("precision", C.c_int32), ("d_model", C.c_int32), ("num_heads", C.c_int32), ("d_kv", C.c_int32),
                ("d_ff", C.c_int32), ("num_layers", C.c_int32), ("layer_norm_eps", C.c_float),
                ("vocab_size", C.c_int32), ("max_length", C.c_int32), ("tie_word_embeddings", C.c_int32),
                ("eos_id", C.c_int32), ("pad_id", C.c_int32), ("start_id", C.c_int32)]


class Res3bCfg(C.Structure):
    """Mirror of ``ymt3_res3b_cfg_t``."""

    _fields_ = [("precision", C.c_int32), ("in_freq", C.c_int32), ("channels", C.c_int32 * 3), ("bn_eps", C.c_float)]


class PtfCfg(C.Structure):
    """Mirror of ``ymt3_ptf_cfg_t``."""

    _fields_ = [(n, C.c_int32) for n in (
        "precision", "num_latents", "d_latent", "kv_dim", "num_blocks", "num_local", "num_temporal", "cross_heads",
        "self_heads", "sca_query_residual", "norm_type", "ff_type", "ff_widening", "moe_experts", "moe_topk", "act",
        "pos_type", "rope_dim", "max_time")] + [("norm_eps", C.c_float)]


ACT_CODES = {None: 0, "none": 0, "gelu_new": 1, "relu": 2, "silu": 3, "gelu": 4}
DTYPE_F32, DTYPE_BF16 = 0, 1
CODEC_MELSPEC, CODEC_SPEC = 0, 1
_P = C.c_void_p
_I64 = C.c_int64
_I = C.c_int

# name -> (restype, argtypes); must list EVERY symbol declared in include/ymt3_b200.h
SIGNATURES = {
    "ymt3_last_error": (C.c_char_p, []),
    "ymt3_abi_version": (_I, []),
    "ymt3_device_info": (_I, [C.c_char_p, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "ymt3_frontend_create": (_I, [C.POINTER(AudioCfg), _P, _P, C.POINTER(_P)]),
    "ymt3_frontend_destroy": (_I, [_P]),
    "ymt3_frontend_num_frames": (_I64, [_P, _I64]),
    "ymt3_frontend_num_features": (_I64, [_P]),
    "ymt3_logmel_f32": (_I, [_P, _P, _I64, _I64, _P, _P]),
    "ymt3_num_segments": (_I64, [_I64, _I64]),
    "ymt3_logmel_waveform_f32": (_I, [_P, _P, _I64, _I64, _P, _P]),
    "ymt3_logmel_host_f32": (_I, [_P, _P, _I64, _I64, _P, _P]),
    "ymt3_t5enc_create": (_I, [C.POINTER(T5Cfg), C.POINTER(Tensor), _I, C.POINTER(_P)]),
    "ymt3_t5enc_destroy": (_I, [_P]),
    "ymt3_t5enc_forward": (_I, [_P, _P, _I64, _I64, _P, _P]),
    "ymt3_t5dec_create": (_I, [C.POINTER(T5Cfg), C.POINTER(Tensor), _I, C.POINTER(_P)]),
    "ymt3_t5dec_destroy": (_I, [_P]),
    "ymt3_t5dec_generate": (_I, [_P, _P, _I64, _I64, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "ymt3_t5dec_generate_prefixed": (_I, [_P, _P, _I64, _I64, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "ymt3_t5dec_generate_latent": (_I, [_P, _P, _I64, _I64, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P,
                                        _P]),
    "ymt3_t5dec_score_forced": (_I, [_P, _P, _I64, _I64, C.c_int32, _P, C.c_int32, _P, _P, C.c_int32, _P, _P]),
    "ymt3_t5dec_last_logits": (_I, [_P, _P, _I64, _P]),
    "ymt3_res3b_create": (_I, [C.POINTER(Res3bCfg), C.POINTER(Tensor), _I, C.POINTER(_P)]),
    "ymt3_res3b_destroy": (_I, [_P]),
    "ymt3_res3b_forward": (_I, [_P, _P, _I64, _I64, _P, _P]),
    "ymt3_ptf_create": (_I, [C.POINTER(PtfCfg), C.POINTER(Tensor), _I, C.POINTER(_P)]),
    "ymt3_ptf_destroy": (_I, [_P]),
    "ymt3_ptf_forward": (_I, [_P, _P, _I64, _I64, _I64, _P, _P]),
    "ymt3_op_permute_btcd_bctd": (_I, [C.c_int32, _P, _P, _I64, _I64, _I64, _I64, _P]),
    "ymt3_op_convert": (_I, [_P, C.c_int32, _P, C.c_int32, _I64, _P]),
    "ymt3_op_linear": (_I, [C.c_int32, _P, _I64, _P, _I64, _P, _P, _I64, _P, _I64, _I64, _I64, _I64, C.c_int32,
                            C.c_int32, C.c_float, C.c_int32, _P]),
    "ymt3_op_rmsnorm": (_I, [C.c_int32, _P, _P, _P, _I64, _I64, C.c_float, _P]),
    "ymt3_op_layernorm": (_I, [C.c_int32, _P, _P, _P, _P, _I64, _I64, C.c_float, _P]),
    "ymt3_op_attention": (_I, [C.c_int32, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, C.c_float, C.c_int32, _P]),
    "ymt3_op_linear_normfused": (_I, [_P, _I64, _P, _I64, _P, _P, _I64, C.c_float, _P, _I64, _P, _I64, _P, _I64, _I64, _I64,
                                      C.c_int32, C.c_int32, C.c_float, C.c_int32, _P]),
    "ymt3_op_linear_chain_counters": (_I64, [_I64]),
    "ymt3_debug_chain_trace": (_I, [_P]),
    "ymt3_op_linear_chain": (_I, [_P, C.c_int32, _I64, _P, C.c_int32, _P]),
    "ymt3_op_linear_argmax": (_I, [C.c_int32, _P, _I64, _P, _I64, _P, _P, _I64, _I64, _I64, _I64, _I64, C.c_float, _P, _P]),
    "ymt3_op_decode_attention": (_I, [C.c_int32, _P, _P, _P, _P, _P, _P, _I64, _P, _I64, _I64, _I64, _P]),
    "ymt3_op_cross_attn_absorbed": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64, _P]),
    "ymt3_op_moe_workspace_bytes": (_I64, [_I64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "ymt3_op_moe_ff": (_I, [C.c_int32, _P, _P, _P, _I64, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                            _P, _P]),
}


def load() -> C.CDLL:
    """Load the library once; raise loudly if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"yourmt3_b200: CUDA extension not built ({LIB_PATH} missing). "
                    "Run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root. "
                    "There is no CPU fallback.")
            lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the .so is stale
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ymt3_last_error()
        raise RuntimeError(f"yourmt3_b200 {what} failed (status {rc}): {msg.decode() if msg else '?'}")


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def device_info() -> dict:
    lib = load()
    name = C.create_string_buffer(256)
    sms, maj, mnr = _I(), _I(), _I()
    check(lib.ymt3_device_info(name, 256, C.byref(sms), C.byref(maj), C.byref(mnr)), "device_info")
    return {"name": name.value.decode(), "num_sms": sms.value, "cc": (maj.value, mnr.value)}


def tensor_table(named):
    """dict name -> CUDA float32 tensor  ==>  (ctypes array of ymt3_tensor_t, n, keepalive list)."""
    import torch
    keep, arr = [], (Tensor * len(named))()
    for i, (name, t) in enumerate(named.items()):
        if not (isinstance(t, torch.Tensor) and t.is_cuda):
            raise RuntimeError(f"tensor {name!r} must live on the CUDA device (no CPU fallback)")
        t = t.detach()
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(torch.float32).contiguous()
        if t.dim() > 4:
            raise ValueError(f"tensor {name!r}: rank > 4")
        b = name.encode()
        keep += [t, b]
        arr[i].name, arr[i].data, arr[i].dtype, arr[i].ndim = b, t.data_ptr(), DTYPE_F32, t.dim()
        for j, d in enumerate(t.shape):
            arr[i].shape[j] = d
    return arr, len(named), keep


def torch_dtype(precision: int):
    import torch
    return torch.float32 if precision == DTYPE_F32 else torch.bfloat16
