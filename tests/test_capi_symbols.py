"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import os
import re

from yourmt3_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ymt3_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"YMT3_API\s+[\w\s\*]+?\b(ymt3_\w+)\s*\(", src)))


def test_exports_match_header(native_lib):
    syms = declared_symbols()
    assert len(syms) >= 9
    for s in syms:
        assert hasattr(native_lib, s), f"{s} declared in the header but not exported"
    # the ctypes signature table must cover the header exactly
    assert sorted(_lib.SIGNATURES) == syms


def test_abi_version_and_error_string(native_lib):
    assert native_lib.ymt3_abi_version() == 1
    # argument validation happens before any CUDA call -> safe without a GPU
    rc = native_lib.ymt3_frontend_create(None, None, None, None)
    assert rc == 1
    assert b"null" in native_lib.ymt3_last_error()
    cfg = _lib.AudioCfg(n_fft=1024, hop_length=128, codec=0, n_mels=512, power_mode=1, log_eps=1e-5)
    h = ctypes.c_void_p()
    buf = (ctypes.c_float * 2048)()
    rc = native_lib.ymt3_frontend_create(ctypes.byref(cfg), buf, buf, ctypes.byref(h))
    assert rc == 1 and b"n_fft=2048" in native_lib.ymt3_last_error()


def test_op_argument_validation_without_gpu(native_lib):
    """every per-op entry point checks shapes / alignment / null pointers BEFORE its first CUDA call, returns a
    non-zero status and leaves a message (the same contract the GPU tests rely on for error behaviour)."""
    buf = (ctypes.c_float * 4096)()
    p = ctypes.addressof(buf)
    # fused arg-max vocab projection: V must be in (0, N], keys must be given
    assert native_lib.ymt3_op_linear_argmax(1, p, 64, p, 64, None, p, 64, 8, 64, 64, 65, 1.0, p, None) != 0
    assert b"bad shape" in native_lib.ymt3_last_error()
    assert native_lib.ymt3_op_linear_argmax(1, p, 64, p, 64, None, p, 64, 8, 64, 64, 60, 1.0, None, None) != 0
    assert b"null keys" in native_lib.ymt3_last_error()
    # tensor-core linear: K must be a multiple of 8
    assert native_lib.ymt3_op_linear(1, p, 60, p, 60, None, p, 64, None, 0, 8, 64, 60, 0, 0, 1.0, 1, None) != 0
    assert b"multiples of 8" in native_lib.ymt3_last_error()
    # decode attention: cross mode needs 0 < fixed_len <= capacity; self mode needs the device step
    assert native_lib.ymt3_op_decode_attention(1, p, None, None, p, p, None, 300, p, 2, 6, 256, None) != 0
    assert b"bad length" in native_lib.ymt3_last_error()
    assert native_lib.ymt3_op_decode_attention(1, p, p, p, p, p, None, 0, p, 2, 6, 256, None) != 0
    assert b"bad length" in native_lib.ymt3_last_error()
    # absorbed cross-attention: padded length must be a multiple of 16 and <= 128
    assert native_lib.ymt3_op_cross_attn_absorbed(p, p, p, 4, 6, 110, 111, None) != 0
    assert b"encoder length" in native_lib.ymt3_last_error()
    assert native_lib.ymt3_op_cross_attn_absorbed(p, p, p, 4, 9, 110, 112, None) != 0
    assert b"heads" in native_lib.ymt3_last_error()
    # chained linears: 1..4 phases, counters given; every phase needs N % 32 == 0 and out_scale 1; gated phases are
    # gelu_new without bias / residual
    P = _lib.ChainPhase
    ok = P(p, 64, p, 64, None, None, 0, 0.0, p, 64, None, 0, None, 64, 64, 0, 0, 1.0)
    arr = (P * 1)(ok)
    assert native_lib.ymt3_op_linear_chain(arr, 0, 8, p, 0, None) != 0
    assert b"bad arguments" in native_lib.ymt3_last_error()
    assert native_lib.ymt3_op_linear_chain(arr, 1, 8, None, 0, None) != 0
    assert b"bad arguments" in native_lib.ymt3_last_error()
    arr = (P * 1)(P(p, 64, p, 64, None, None, 0, 0.0, p, 72, None, 0, None, 72, 64, 0, 0, 1.0))
    assert native_lib.ymt3_op_linear_chain(arr, 1, 8, p, 0, None) != 0
    assert b"N % 32" in native_lib.ymt3_last_error()
    arr = (P * 1)(P(p, 64, p, 64, None, None, 0, 0.0, p, 32, p, 32, None, 64, 64, 1, 1, 1.0))
    assert native_lib.ymt3_op_linear_chain(arr, 1, 8, p, 0, None) != 0
    assert b"gated phases" in native_lib.ymt3_last_error()
    assert native_lib.ymt3_op_linear_chain_counters(300) == 4 * 3


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import pytest
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_forward_on_cpu_tensor_raises():
    import pytest
    import torch
    from yourmt3_b200 import Melspectrogram
    with pytest.raises(RuntimeError, match="CUDA only"):
        Melspectrogram()(torch.zeros(1, 1, 4096))
