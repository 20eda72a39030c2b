"""ctypes loader for the C-ABI library (include/ymt3_b200.h).

There is deliberately no fallback: if ``libymt3_b200.so`` is missing or a call
fails, a RuntimeError is raised (north-star: no CPU fallback, no backend dispatch).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libymt3_b200.so")

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
    "ymt3_logmel_host_f32": (_I, [_P, _P, _I64, _I64, _P, _P]),
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
