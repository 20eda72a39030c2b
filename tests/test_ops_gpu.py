"""GPU unit-parity of individual kernels through the per-op C ABI, against plain PyTorch fp32
on CPU (the oracle for floating-point kernels). fp32 path tolerance: 2e-5 relative to the
output scale (fp32 accumulate, different summation order than MKL)."""
import math

import pytest
import torch

from oracle import t5 as OT
from yourmt3_b200 import _lib

pytestmark = pytest.mark.gpu
ACTS = {0: lambda x: x, 1: OT.gelu_new, 2: torch.relu, 3: torch.nn.functional.silu, 4: torch.nn.functional.gelu}


def _close(got, ref, tol=2e-5):
    scale = max(1.0, float(ref.abs().max()))
    err = float((got.double() - ref.double()).abs().max()) / scale
    assert err < tol, f"max err {err:.3e} (scale {scale:.2f})"


def linear_native(lib, dev, A, W, bias=None, act=0, gated=0, residual=None, out_scale=1.0):
    M, K = A.shape
    N = W.shape[0]
    Ad, Wd = A.to(dev), W.to(dev)
    No = N // 2 if gated else N
    C_ = torch.full((M, No), float("nan"), device=dev)
    bd = None if bias is None else bias.to(dev)
    Rd = None if residual is None else residual.to(dev)
    rc = lib.ymt3_op_linear(0, Ad.data_ptr(), K, Wd.data_ptr(), K, None if bd is None else bd.data_ptr(), C_.data_ptr(),
                            No, None if Rd is None else Rd.data_ptr(), No, M, N, K, act, gated, out_scale, 0,
                            torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "op_linear")
    return C_.cpu()


@pytest.mark.parametrize("M,N,K", [(1, 4, 4), (7, 12, 20), (64, 64, 16), (100, 1152, 512), (832, 512, 384),
                                   (3000, 2048, 512), (16384, 512, 1024), (333, 596, 512)])
def test_linear_f32_shapes(cuda_device, native_lib, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.05
    _close(linear_native(native_lib, cuda_device, A, W), A @ W.T)


@pytest.mark.parametrize("act", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("gated", [0, 1])
def test_linear_f32_epilogues(cuda_device, native_lib, act, gated):
    g = torch.Generator().manual_seed(10 * act + gated)
    M, N, K = 200, 256, 128
    A, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.1
    bias = torch.randn(N, generator=g)
    R = torch.randn(M, N // 2 if gated else N, generator=g)
    z = A @ W.T + bias
    ref = ACTS[act](z[:, 0::2]) * z[:, 1::2] if gated else ACTS[act](z)
    ref = R + 0.5 * ref
    _close(linear_native(native_lib, cuda_device, A, W, bias, act, gated, R, 0.5), ref)


def test_linear_inplace_residual(cuda_device, native_lib):
    g = torch.Generator().manual_seed(3)
    M, N, K = 130, 512, 384
    A, W, X = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.05, torch.randn(M, N, generator=g)
    Xd, Ad, Wd = X.to(cuda_device), A.to(cuda_device), W.to(cuda_device)
    rc = native_lib.ymt3_op_linear(0, Ad.data_ptr(), K, Wd.data_ptr(), K, None, Xd.data_ptr(), N, Xd.data_ptr(), N, M, N,
                                   K, 0, 0, 1.0, 0, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)
    _close(Xd.cpu(), X + A @ W.T)


def test_linear_rejects_bad_alignment(cuda_device, native_lib):
    A = torch.zeros(4, 6, device=cuda_device)
    rc = native_lib.ymt3_op_linear(0, A.data_ptr(), 6, A.data_ptr(), 6, None, A.data_ptr(), 4, None, 0, 4, 4, 6, 0, 0, 1.0,
                                   0, None)
    assert rc == 1 and b"multiples of 4" in native_lib.ymt3_last_error()


@pytest.mark.parametrize("rows,dim", [(1, 512), (1000, 512), (77, 128), (5, 1024)])
def test_norms(cuda_device, native_lib, rows, dim):
    g = torch.Generator().manual_seed(rows + dim)
    x, w, b = torch.randn(rows, dim, generator=g) * 3 + 0.5, torch.randn(dim, generator=g), torch.randn(dim, generator=g)
    xd, wd, bd = x.to(cuda_device), w.to(cuda_device), b.to(cuda_device)
    y = torch.empty_like(xd)
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(native_lib.ymt3_op_rmsnorm(0, xd.data_ptr(), wd.data_ptr(), y.data_ptr(), rows, dim, 1e-6, s))
    _close(y.cpu(), OT.rms_norm(x, w, 1e-6), 1e-5)
    _lib.check(native_lib.ymt3_op_layernorm(0, xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), rows, dim, 1e-5, s))
    _close(y.cpu(), torch.nn.functional.layer_norm(x, (dim,), w, b, 1e-5), 1e-5)


@pytest.mark.parametrize("B,H,Sq,Sk,dk,causal", [
    (2, 6, 256, 256, 64, 0),     # T5 encoder
    (3, 6, 7, 7, 64, 1),         # decoder prefill, causal
    (2, 6, 5, 19, 64, 0),        # cross attention
    (5, 1, 26, 128, 128, 0),     # Perceiver-TF spectral cross-attention
    (4, 8, 26, 26, 16, 0),       # latent self-attention
    (3, 8, 110, 110, 16, 0),     # temporal self-attention
    (2, 4, 33, 70, 32, 1),       # causal with Sk > Sq
    (1, 2, 1, 300, 64, 0),
])
def test_attention(cuda_device, native_lib, B, H, Sq, Sk, dk, causal):
    g = torch.Generator().manual_seed(B * 1000 + Sq + Sk + dk)
    q, k, v = (torch.randn(B, S, H, dk, generator=g) for S in (Sq, Sk, Sk))
    scale = 1.0 if dk == 64 else 1.0 / math.sqrt(dk)
    mask = None
    if causal:
        i, j = torch.arange(Sq)[:, None], torch.arange(Sk)[None, :]
        mask = torch.where(j <= i + (Sk - Sq), 0.0, float("-inf"))
    ref = OT.attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), mask, scale).view(B, Sq, H, dk)
    qd, kd, vd = q.to(cuda_device), k.to(cuda_device), v.to(cuda_device)
    o = torch.full((B, Sq, H, dk), float("nan"), device=cuda_device)
    _lib.check(native_lib.ymt3_op_attention(0, qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), o.data_ptr(), B, H, Sq, Sk, dk,
                                            scale, causal, torch.cuda.current_stream().cuda_stream))
    _close(o.cpu(), ref, 1e-5)


# ----------------------------------------------------------------------------------------------
# bf16 path: tcgen05 GEMM (TMEM accumulators, TMA operands). Reference = fp32 matmul of the
# bf16-rounded operands; tolerance = bf16 output rounding (2^-8) + fp32 accumulation noise.
# ----------------------------------------------------------------------------------------------
def linear_bf16_native(lib, dev, A, W, bias=None, act=0, gated=0, residual=None, out_scale=1.0, out_f32=False):
    M, K = A.shape
    N = W.shape[0]
    Ad, Wd = A.to(dev, torch.bfloat16), W.to(dev, torch.bfloat16)
    No = N // 2 if gated else N
    odt = torch.float32 if out_f32 else torch.bfloat16
    C_ = torch.full((M, No), float("nan"), device=dev, dtype=odt)
    bd = None if bias is None else bias.to(dev)
    Rd = None if residual is None else residual.to(dev, odt)
    rc = lib.ymt3_op_linear(1, Ad.data_ptr(), K, Wd.data_ptr(), K, None if bd is None else bd.data_ptr(), C_.data_ptr(),
                            No, None if Rd is None else Rd.data_ptr(), No, M, N, K, act, gated, out_scale,
                            0 if out_f32 else 1, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "op_linear bf16")
    torch.cuda.synchronize()
    return C_.float().cpu()


def _bf(x):
    return x.to(torch.bfloat16).float()


@pytest.mark.parametrize("M,N,K", [
    (128, 128, 64), (128, 32, 64), (1, 64, 128), (200, 1152, 512), (832, 512, 384), (5000, 2048, 512),
    (40000, 512, 1024), (333, 600, 512), (16384, 128, 256), (77, 96, 72), (300, 1536, 128),
    (20000, 600, 512),        # 128 x 256 tiles with a ragged last N tile (88 valid columns, TMA-store box clipped)
    (6656, 1152, 512),        # the decode-step qkv shape at the default batch (256-wide tiles, 4.5 N tiles)
    (2500, 640, 2112),        # K >= 2048, 20 M tiles: the CTA-PAIR kernel (cta_group::2), ragged in M (odd super-tile
                              # count: the last pair has one CTA entirely beyond M), N (2.5 tiles) and K (33 k-blocks)
    (4096, 1024, 2048),       # CTA pair, exact tiles
])
def test_linear_bf16_tcgen05_shapes(cuda_device, native_lib, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.05
    ref = _bf(A) @ _bf(W).T
    got = linear_bf16_native(native_lib, cuda_device, A, W)
    _close(got, ref, 6e-3)
    got32 = linear_bf16_native(native_lib, cuda_device, A, W, out_f32=True)
    _close(got32, ref, 2e-5 if K < 2048 else 4e-5)   # fp32 accumulate in TMEM, fp32 out: only summation order differs


@pytest.mark.parametrize("act,gated", [(0, 0), (1, 1), (3, 1), (4, 0), (2, 0)])
@pytest.mark.parametrize("out_f32", [False, True])
def test_linear_bf16_epilogues(cuda_device, native_lib, act, gated, out_f32):
    g = torch.Generator().manual_seed(17 * act + gated)
    M, N, K = 300, 256, 192
    A, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.1
    bias = torch.randn(N, generator=g)
    R = torch.randn(M, N // 2 if gated else N, generator=g)
    z = _bf(A) @ _bf(W).T + bias
    ref = ACTS[act](z[:, 0::2]) * z[:, 1::2] if gated else ACTS[act](z)
    ref = (R if out_f32 else _bf(R)) + 0.5 * ref
    got = linear_bf16_native(native_lib, cuda_device, A, W, bias, act, gated, R, 0.5, out_f32)
    _close(got, ref, 3e-5 if out_f32 else 6e-3)


@pytest.mark.parametrize("act", [1, 3])
def test_linear_bf16_gated_wide_tiles(cuda_device, native_lib, act):
    """gated epilogues on the 128 x 256 tile path (large M, N = 1024): the MoE expert / decoder FFN shapes."""
    g = torch.Generator().manual_seed(act)
    M, N, K = 20000, 1024, 128
    A, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.1
    z = _bf(A) @ _bf(W).T
    ref = ACTS[act](z[:, 0::2]) * z[:, 1::2]
    got = linear_bf16_native(native_lib, cuda_device, A, W, None, act, 1, None, 1.0, False)
    _close(got, ref, 6e-3)


def test_norm_attention_bf16_io(cuda_device, native_lib):
    g = torch.Generator().manual_seed(5)
    x, w = torch.randn(100, 512, generator=g), torch.randn(512, generator=g)
    xd, wd = x.to(cuda_device, torch.bfloat16), w.to(cuda_device)
    y = torch.empty_like(xd)
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(native_lib.ymt3_op_rmsnorm(1, xd.data_ptr(), wd.data_ptr(), y.data_ptr(), 100, 512, 1e-6, s))
    _close(y.float().cpu(), OT.rms_norm(_bf(x), w, 1e-6), 6e-3)
    B, H, Sq, Sk, dk = 2, 6, 50, 70, 64
    q, k, v = (torch.randn(B, S, H, dk, generator=g) * 0.5 for S in (Sq, Sk, Sk))
    ref = OT.attention(_bf(q).transpose(1, 2), _bf(k).transpose(1, 2), _bf(v).transpose(1, 2)).view(B, Sq, H, dk)
    qd, kd, vd = (t.to(cuda_device, torch.bfloat16) for t in (q, k, v))
    o = torch.empty_like(qd)
    _lib.check(native_lib.ymt3_op_attention(1, qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), o.data_ptr(), B, H, Sq, Sk, dk,
                                            1.0, 0, s))
    _close(o.float().cpu(), ref, 6e-3)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("N,H,step,Lcap", [(5, 6, 0, 16), (7, 6, 37, 64), (300, 6, 200, 256),
                                            # few sequences, long cache: 8 / 4 warps per (sequence, head), incl. a
                                            # ragged last block (step + 1 not a multiple of 16) and idle warps
                                            (3, 6, 500, 1024), (64, 6, 1023, 1024), (150, 6, 3, 64), (1, 1, 17, 32)])
def test_decode_attention_self_and_cross(cuda_device, native_lib, dtype, N, H, step, Lcap):
    """single-query attention over the KV cache (decode.cu) vs plain torch fp32: self mode appends the new K/V row at
    *step and attends [0, step]; cross mode attends [0, fixed_len)."""
    td, code = (torch.float32, 0) if dtype == "f32" else (torch.bfloat16, 1)
    g = torch.Generator().manual_seed(N + step)
    q = (torch.randn(N, H, 64, generator=g) * 0.5).to(cuda_device, td)
    kn = torch.randn(N, H, 64, generator=g).to(cuda_device, td)
    vn = torch.randn(N, H, 64, generator=g).to(cuda_device, td)
    Kc = torch.randn(N, H, Lcap, 64, generator=g).to(cuda_device, td)
    Vc = torch.randn(N, H, Lcap, 64, generator=g).to(cuda_device, td)
    K0, V0 = Kc.clone(), Vc.clone()
    st = torch.tensor([step], dtype=torch.int32, device=cuda_device)
    out = torch.empty(N, H, 64, dtype=td, device=cuda_device)
    s = _lib.current_stream_ptr()
    _lib.check(native_lib.ymt3_op_decode_attention(code, q.data_ptr(), kn.data_ptr(), vn.data_ptr(), Kc.data_ptr(),
                                                   Vc.data_ptr(), st.data_ptr(), 0, out.data_ptr(), N, H, Lcap, s))
    assert torch.equal(Kc[:, :, step], kn) and torch.equal(Vc[:, :, step], vn)           # appended in place
    assert torch.equal(Kc[:, :, step + 1:], K0[:, :, step + 1:])                         # nothing else touched
    Kr, Vr = Kc[:, :, :step + 1].float(), Vc[:, :, :step + 1].float()
    p = torch.softmax(torch.einsum("nhd,nhld->nhl", q.float(), Kr), -1)
    ref = torch.einsum("nhl,nhld->nhd", p, Vr)
    tol = 2e-5 if dtype == "f32" else 2e-2
    assert float((out.float() - ref).abs().max()) < tol
    L = step + 1
    _lib.check(native_lib.ymt3_op_decode_attention(code, q.data_ptr(), None, None, K0.data_ptr(), V0.data_ptr(), None,
                                                   L, out.data_ptr(), N, H, Lcap, s))
    p = torch.softmax(torch.einsum("nhd,nhld->nhl", q.float(), K0[:, :, :L].float()), -1)
    ref = torch.einsum("nhl,nhld->nhd", p, V0[:, :, :L].float())
    assert float((out.float() - ref).abs().max()) < tol


@pytest.mark.parametrize("B,H,Sq,Sk", [(5, 1, 26, 128), (3, 1, 24, 128), (2, 2, 32, 100), (4, 1, 7, 128), (3, 1, 16, 40),
                                       (600, 1, 26, 128)])
def test_attention_bf16_dk128_tensor_core(cuda_device, native_lib, B, H, Sq, Sk):
    """wide-head tensor-core cross-attention (attn_wide_tc_kernel: bf16, dk 128, Sq <= 32, Sk <= 128 = the Perceiver-TF
    spectral cross-attention) vs torch fp32 on the same bf16 inputs, and vs the SIMT kernel it replaces."""
    import os
    g = torch.Generator().manual_seed(B + Sq * 7 + Sk)
    q, k, v = (torch.randn(B, S, H, 128, generator=g).to(torch.bfloat16) for S in (Sq, Sk, Sk))
    scale = 1.0 / math.sqrt(128)
    ref = OT.attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2), None,
                       scale).view(B, Sq, H, 128)
    qd, kd, vd = q.to(cuda_device), k.to(cuda_device), v.to(cuda_device)
    outs = []
    for no_tc in (False, True):
        if no_tc:
            os.environ["YMT3_NO_TC_ATTN"] = "1"
        try:
            o = torch.full((B, Sq, H, 128), float("nan"), device=cuda_device, dtype=torch.bfloat16)
            _lib.check(native_lib.ymt3_op_attention(1, qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), o.data_ptr(), B, H, Sq,
                                                    Sk, 128, scale, 0, torch.cuda.current_stream().cuda_stream))
            outs.append(o.float().cpu())
        finally:
            os.environ.pop("YMT3_NO_TC_ATTN", None)
    # bf16 probabilities + bf16 output rounding on O(0.3) values
    assert float((outs[0] - ref).abs().max()) < 2e-2
    assert float((outs[0] - ref).abs().mean()) < 2e-3
    assert float((outs[1] - ref).abs().max()) < 1e-2          # the SIMT kernel (only the output is rounded)


@pytest.mark.parametrize("M,N,K,gated", [(300, 1152, 512, 0), (3328, 2048, 512, 1), (77, 600, 512, 0)])
def test_linear_fused_rmsnorm_consumer(cuda_device, native_lib, M, N, K, gated):
    """GEMM on the un-normalised x with the norm weight folded into W and the row scale taken from sum-of-squares
    partials == rmsnorm kernel followed by the plain GEMM (reference order), to bf16 accuracy."""
    g = torch.Generator().manual_seed(M + N)
    x = (torch.randn(M, K, generator=g) * 1.7).to(cuda_device, torch.bfloat16)
    w_ln = (1.0 + 0.2 * torch.randn(K, generator=g)).to(cuda_device)
    W = (torch.randn(N, K, generator=g) * 0.05).to(cuda_device)
    eps = 1e-6
    ss = (x.float() ** 2).view(M, K // 32, 32).sum(-1).contiguous()
    Wn = (W * w_ln[None, :]).to(torch.bfloat16).contiguous()
    No = N // 2 if gated else N
    out = torch.empty(M, No, dtype=torch.bfloat16, device=cuda_device)
    act = 1 if gated else 0
    _lib.check(native_lib.ymt3_op_linear_normfused(x.data_ptr(), K, Wn.data_ptr(), K, None, ss.data_ptr(), K // 32, eps,
                                                   out.data_ptr(), No, None, 0, None, M, N, K, act, gated, 1.0, 1,
                                                   _lib.current_stream_ptr()))
    xf = x.float()
    h = xf * torch.rsqrt((xf ** 2).mean(-1, keepdim=True) + eps) * w_ln
    y = h @ W.T
    if gated:   # rows interleaved: 2j = activated branch, 2j+1 = linear branch
        y = OT.gelu_new(y[:, 0::2]) * y[:, 1::2]
    err = float((out.float() - y).abs().max()) / max(1.0, float(y.abs().max()))
    assert err < 2e-2, err


def test_linear_fused_rmsnorm_producer(cuda_device, native_lib):
    """the residual GEMM emits, per row and 32-column chunk, the sum of squares of exactly the bf16 values it stored."""
    M, N, K = 333, 512, 384
    g = torch.Generator().manual_seed(9)
    a = torch.randn(M, K, generator=g).to(cuda_device, torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * 0.05).to(cuda_device, torch.bfloat16)
    x = torch.randn(M, N, generator=g).to(cuda_device, torch.bfloat16)
    ss = torch.full((M, N // 32), float("nan"), device=cuda_device)
    _lib.check(native_lib.ymt3_op_linear_normfused(a.data_ptr(), K, W.data_ptr(), K, None, None, 0, 0.0, x.data_ptr(), N,
                                                   x.data_ptr(), N, ss.data_ptr(), M, N, K, 0, 0, 1.0, 1,
                                                   _lib.current_stream_ptr()))
    want = (x.float() ** 2).view(M, N // 32, 32).sum(-1)
    assert float((ss - want).abs().max()) <= 1e-4 * float(want.abs().max())


@pytest.mark.parametrize("B,H,Sq,Sk", [(3, 8, 26, 26), (3, 8, 110, 110), (2, 8, 128, 128), (5, 2, 16, 16), (4, 8, 5, 37),
                                       (700, 8, 110, 110)])
def test_attention_bf16_dk16_tensor_core(cuda_device, native_lib, B, H, Sq, Sk):
    """tiny-sequence tensor-core attention (attn_small_tc_kernel: bf16, dk 16) vs torch fp32 on the same bf16 inputs,
    and vs the fp32-math SIMT kernel it replaces (YMT3_NO_TC_ATTN=1)."""
    import os
    g = torch.Generator().manual_seed(B + Sq * 7 + Sk)
    q, k, v = (torch.randn(B, S, H, 16, generator=g).to(torch.bfloat16) for S in (Sq, Sk, Sk))
    scale = 0.25
    ref = OT.attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2), None,
                       scale).view(B, Sq, H, 16)
    qd, kd, vd = q.to(cuda_device), k.to(cuda_device), v.to(cuda_device)
    outs = []
    for no_tc in (False, True):
        if no_tc:
            os.environ["YMT3_NO_TC_ATTN"] = "1"
        try:
            o = torch.full((B, Sq, H, 16), float("nan"), device=cuda_device, dtype=torch.bfloat16)
            _lib.check(native_lib.ymt3_op_attention(1, qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), o.data_ptr(), B, H, Sq,
                                                    Sk, 16, scale, 0, torch.cuda.current_stream().cuda_stream))
            outs.append(o.float().cpu())
        finally:
            os.environ.pop("YMT3_NO_TC_ATTN", None)
    # bf16 probabilities + bf16 q*scale rounding + bf16 output: a few 1e-3 absolute on O(1) values
    assert float((outs[0] - ref).abs().max()) < 3e-2
    assert float((outs[0] - ref).abs().mean()) < 3e-3
    assert float((outs[1] - ref).abs().max()) < 2e-2          # the SIMT kernel (only the output is rounded)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("M,V,K", [(7, 596, 512), (300, 596, 512), (3328, 1391, 512), (129, 40, 64)])
def test_linear_argmax_fused_epilogue(cuda_device, native_lib, dtype, M, V, K):
    """greedy selection fused into the vocab-projection epilogue: the stored fp32 logits are unchanged and the packed key
    of every row decodes to torch.argmax of those logits over the first V columns (first maximum among equals - rows
    with duplicated weight rows force exact ties, also across tiles); columns >= V never win."""
    td, code = (torch.float32, 0) if dtype == "f32" else (torch.bfloat16, 1)
    g = torch.Generator().manual_seed(M + V)
    Vp = (V + 7) // 8 * 8
    W = torch.randn(Vp, K, generator=g) * 0.05
    W[V:] = 10.0 * W[:Vp - V].abs()                 # padding columns with LARGE logits must be ignored
    W[5] = W[3]; W[V - 1] = W[3]; W[min(V - 2, 200)] = W[3]      # exact ties, some in another N tile
    x = torch.randn(M, K, generator=g)
    x[1] = W[3] * 50.0                              # row 1: the tied columns are the maximum
    if M > 2:
        x[2] = 0.0                                  # row 2: all logits equal (0) -> column 0
    xd, Wd = x.to(cuda_device, td), W.to(cuda_device, td)
    logits = torch.empty(M, Vp, dtype=torch.float32, device=cuda_device)
    keys = torch.zeros(M, dtype=torch.int64, device=cuda_device)
    _lib.check(native_lib.ymt3_op_linear_argmax(code, xd.data_ptr(), K, Wd.data_ptr(), K, None, logits.data_ptr(), Vp, M, Vp, K,
                                                V, 0.125, keys.data_ptr(), _lib.current_stream_ptr()))
    plain = torch.empty_like(logits)
    _lib.check(native_lib.ymt3_op_linear(code, xd.data_ptr(), K, Wd.data_ptr(), K, None, plain.data_ptr(), Vp, None, 0, M, Vp, K,
                                         0, 0, 0.125, 0, _lib.current_stream_ptr()))
    assert torch.equal(logits, plain)
    col = 0xFFFFFFFF - (keys & 0xFFFFFFFF)
    ref = torch.argmax(logits[:, :V], dim=-1)
    assert torch.equal(col, ref), (col[:8], ref[:8])
    assert int(col[1]) == 3 and (M <= 2 or int(col[2]) == 0)


def test_decode_attention_step_beyond_capacity_writes_nothing(cuda_device, native_lib):
    """contract of ymt3_op_decode_attention: *step < Lcap.  A counter at / beyond the capacity must not append out
    of bounds: the launch leaves the caches and `out` untouched (guard in decode_attn_kernel)."""
    lib, dev = native_lib, cuda_device
    N, H, Lcap = 37, 6, 16
    g = torch.Generator().manual_seed(4)
    q, kn, vn = (torch.randn(N, H * 64, generator=g).to(dev) for _ in range(3))
    # caches followed by a guard band in the same allocation: an out-of-bounds append would land in the band
    buf = torch.randn(2, N * H * Lcap * 64 + 4096, generator=g).to(dev)
    before = buf.clone()
    out = torch.full((N, H * 64), 7.0, device=dev)
    for bad in (Lcap, Lcap + 5):
        step = torch.tensor([bad], dtype=torch.int32, device=dev)
        _lib.check(lib.ymt3_op_decode_attention(_lib.DTYPE_F32, q.data_ptr(), kn.data_ptr(), vn.data_ptr(), buf[0].data_ptr(),
                                                buf[1].data_ptr(), step.data_ptr(), 0, out.data_ptr(), N, H, Lcap,
                                                _lib.current_stream_ptr()), "decode_attention")
        torch.cuda.synchronize()
        assert torch.equal(buf, before) and bool((out == 7.0).all())
    step = torch.tensor([Lcap - 1], dtype=torch.int32, device=dev)       # last legal step still works
    _lib.check(lib.ymt3_op_decode_attention(_lib.DTYPE_F32, q.data_ptr(), kn.data_ptr(), vn.data_ptr(), buf[0].data_ptr(),
                                            buf[1].data_ptr(), step.data_ptr(), 0, out.data_ptr(), N, H, Lcap,
                                            _lib.current_stream_ptr()), "decode_attention")
    torch.cuda.synchronize()
    assert not bool((out == 7.0).any()) and torch.equal(buf[:, N * H * Lcap * 64:], before[:, N * H * Lcap * 64:])


@pytest.mark.parametrize("M", [100, 1000])
def test_linear_chain_matches_separate_launches(cuda_device, native_lib, M):
    """ymt3_op_linear_chain (one persistent launch, per-row-tile dependency counters) == the same phases issued one by
    one through ymt3_op_linear_normfused, bit for bit: [A0 W0^T + bias + x -> x (sum of squares out)] ->
    [norm(x) W1^T, gated gelu_new] -> [g W2^T + x -> x (sum of squares out)] -> [norm(x) W3^T]; three launches on the
    same counters (launch ordinal 0, 1, 2), M = 1000: eight row tiles, the last ragged."""
    D, K0, F, N3 = 512, 384, 1024, 1152
    g = torch.Generator().manual_seed(M)
    rnd = lambda *sh, sc=1.0: (torch.randn(*sh, generator=g) * sc).to(cuda_device, torch.bfloat16)  # noqa: E731
    a0, x0 = rnd(M, K0), rnd(M, D)
    W0, W1, W2, W3 = rnd(D, K0, sc=.05), rnd(2 * F, D, sc=.05), rnd(D, F, sc=.03), rnd(N3, D, sc=.05)
    b0 = (torch.randn(D, generator=g) * 0.1).to(cuda_device)
    s_ = _lib.current_stream_ptr()

    def bufs():
        return dict(x=x0.clone(), gbuf=torch.zeros(M, F, device=cuda_device, dtype=torch.bfloat16),
                    out=torch.zeros(M, N3, device=cuda_device, dtype=torch.bfloat16),
                    s1=torch.zeros(M, D // 32, device=cuda_device), s2=torch.zeros(M, D // 32, device=cuda_device))

    def phases(b):
        P = _lib.ChainPhase
        return [P(a0.data_ptr(), K0, W0.data_ptr(), K0, b0.data_ptr(), None, 0, 0.0, b["x"].data_ptr(), D, b["x"].data_ptr(), D,
                  b["s1"].data_ptr(), D, K0, 0, 0, 1.0),
                P(b["x"].data_ptr(), D, W1.data_ptr(), D, None, b["s1"].data_ptr(), D // 32, 1e-6, b["gbuf"].data_ptr(), F, None, 0,
                  None, 2 * F, D, 1, 1, 1.0),
                P(b["gbuf"].data_ptr(), F, W2.data_ptr(), F, None, None, 0, 0.0, b["x"].data_ptr(), D, b["x"].data_ptr(), D,
                  b["s2"].data_ptr(), D, F, 0, 0, 1.0),
                P(b["x"].data_ptr(), D, W3.data_ptr(), D, None, b["s2"].data_ptr(), D // 32, 1e-6, b["out"].data_ptr(), N3, None, 0,
                  None, N3, D, 0, 0, 1.0)]

    ref, got = bufs(), bufs()
    pr, pg = phases(ref), phases(got)
    arr = (_lib.ChainPhase * 4)(*pg)
    cnt = torch.zeros(int(native_lib.ymt3_op_linear_chain_counters(M)), device=cuda_device, dtype=torch.int32)
    for ordinal in range(3):
        for q in pr:
            _lib.check(native_lib.ymt3_op_linear_normfused(q.A, q.lda, q.W, q.ldw, q.bias, q.ss_in, q.chunks, q.eps, q.C, q.ldc,
                                                           q.residual, q.ldr, q.ss_out, M, q.N, q.K, q.act, q.gated, 1.0, 1, s_))
        _lib.check(native_lib.ymt3_op_linear_chain(arr, 4, M, cnt.data_ptr(), ordinal, s_))
        torch.cuda.synchronize()
        for k in ref:
            assert torch.equal(ref[k], got[k]), (ordinal, k)
    assert float(ref["out"].float().abs().mean()) > 1e-3


_LEAN_SCRIPT = r"""
import sys, torch
sys.path.insert(0, sys.argv[1])
from yourmt3_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda")
g = torch.Generator().manual_seed(3)
outs = {}
s_ = None
for name, (M, N, K, gated, act, bias, res, ss, norm) in {
        "plain": (300, 1152, 512, 0, 0, 0, 0, 0, 0), "bias": (1000, 384, 128, 0, 0, 1, 0, 0, 0),
        "res_ss": (333, 512, 384, 0, 0, 0, 1, 1, 0), "bias_res_ss": (700, 512, 1536, 0, 0, 1, 1, 1, 0),
        "bias_res": (129, 128, 128, 0, 0, 1, 1, 0, 0), "norm_plain": (9464, 1536, 512, 0, 0, 0, 0, 0, 1),
        "gated_gelu": (300, 2048, 512, 1, 1, 0, 0, 0, 1), "gated_silu_bias": (5000, 1024, 128, 1, 3, 1, 0, 0, 0)}.items():
    A = torch.randn(M, K, generator=g).to(dev, torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * 0.05).to(dev, torch.bfloat16)
    No = N // 2 if gated else N
    C = torch.randn(M, No, generator=g).to(dev, torch.bfloat16)          # also the residual (in place)
    b = (torch.randn(N, generator=g) * 0.1).to(dev) if bias else None
    sso = torch.zeros(M, No // 32, device=dev) if ss else None
    ssi = (A.float() ** 2).view(M, K // 32, 32).sum(-1).contiguous() if norm else None
    _lib.check(lib.ymt3_op_linear_normfused(A.data_ptr(), K, W.data_ptr(), K, b.data_ptr() if bias else None,
                                            ssi.data_ptr() if norm else None, K // 32 if norm else 0, 1e-6, C.data_ptr(), No,
                                            C.data_ptr() if res else None, No, sso.data_ptr() if ss else None, M, N, K, act, gated,
                                            1.0, 1, s_))
    torch.cuda.synchronize()
    outs[name] = C.cpu()
    if ss:
        outs[name + "_ss"] = sso.cpu()
torch.save(outs, sys.argv[2])
"""


def test_lean_epilogue_bit_identical_to_general(cuda_device, native_lib, tmp_path):
    """The lean GEMM epilogue instances (compile-time bias / residual / sum-of-squares / gated flags, packed f32x2 math,
    pipelined 16-column halves) produce the same BITS as the general epilogue path (YMT3_GEMM_NO_LEAN=1, read once per
    process -> two subprocesses): plain, bias, residual + sum of squares, gated gelu_new / SiLU, fused-norm consumer;
    ragged M, partially filled last N tile (1152 = 4.5 x 256)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for mode in ("lean", "general"):
        env = dict(os.environ)
        env.pop("YMT3_GEMM_NO_LEAN", None)
        if mode == "general":
            env["YMT3_GEMM_NO_LEAN"] = "1"
        out = tmp_path / f"{mode}.pt"
        r = subprocess.run([sys.executable, "-c", _LEAN_SCRIPT, root, str(out)], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        res[mode] = torch.load(out)
    assert set(res["lean"]) == set(res["general"])
    for k in res["lean"]:
        a, b = res["lean"][k], res["general"][k]
        assert float(a.float().abs().mean()) > 1e-3, k
        assert torch.equal(a, b), (k, float((a.float() - b.float()).abs().max()))
