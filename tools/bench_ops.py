"""Micro-benchmarks of individual kernels through the C ABI (CUDA events, median of N).
usage: python tools/bench_ops.py [gemm|attn|decode_attn|all] [--once]   (--once: single pass for ncu)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from yourmt3_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
once = "--once" in sys.argv
what = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "all"


def timeit(fn, reps=20):
    if once:
        fn()
        torch.cuda.synchronize()
        return float("nan")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def gemm(M, N, K, gated=0, act=0, residual=False, name=""):
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    No = N // 2 if gated else N
    C = torch.empty(M, No, device=dev, dtype=torch.bfloat16)
    R = torch.randn(M, No, device=dev).bfloat16() if residual else None
    s = torch.cuda.current_stream().cuda_stream

    def run():
        _lib.check(lib.ymt3_op_linear(1, A.data_ptr(), K, W.data_ptr(), K, None, C.data_ptr(), No,
                                      R.data_ptr() if residual else None, No, M, N, K, act, gated, 1.0, 1, s))
    us = timeit(run)
    flops = 2.0 * M * N * K
    byts = (M * K + N * K + M * No * (2 if residual else 1)) * 2
    print(f"gemm {name:28s} M={M:8d} N={N:5d} K={K:5d} gated={gated} res={int(residual)}: {us:9.1f} us  "
          f"{flops / us / 1e6:8.1f} TFLOP/s  {byts / us / 1e3:8.1f} GB/s", flush=True)


if what in ("gemm", "all"):
    gemm(183040, 384, 128, name="ptf qkv")
    gemm(183040, 128, 128, residual=True, name="ptf out-proj+res")
    gemm(366080, 1024, 128, gated=1, act=3, name="moe gemm1 (ungrouped equiv)")
    gemm(366080, 128, 512, name="moe gemm2 (ungrouped equiv)")
    gemm(901120, 256, 128, name="ptf sca kv")
    gemm(91520, 768, 512, name="decoder cross kv")
    gemm(832, 1152, 512, name="decode qkv")
    gemm(832, 512, 384, residual=True, name="decode o-proj+res")
    gemm(832, 2048, 512, gated=1, act=1, name="decode ffn wi")
    gemm(832, 512, 1024, residual=True, name="decode ffn wo+res")
    gemm(16384, 1152, 512, name="t5 enc qkv B=64")
    gemm(8192, 8192, 8192, name="square 8k")

if what == "decode728":
    # the GEMMs of one YPTF.MoE+Multi decode step at the default bench batch (728 segments x 13 channels)
    M = 9464
    gemm(M, 1152, 512, name="dec self qkv")
    gemm(M, 512, 384, residual=True, name="dec self o+res")
    gemm(M, 1536, 512, name="dec cross q absorbed")
    gemm(M, 512, 1536, residual=True, name="dec cross o absorbed+res")
    gemm(M, 2048, 512, gated=1, act=1, name="dec ffn wi")
    gemm(M, 512, 1024, residual=True, name="dec ffn wo+res")
    gemm(M, 600, 512, name="dec lm head")

if what == "chain728":
    # the two GEMM chains of one YPTF.MoE+Multi decoder layer at the default bench batch, chained vs one by one
    import ctypes
    M = int(os.environ.get("CHAIN_M", 9464))
    D, INNER, HZ, F = 512, 384, 1536, 1024
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *sh, sc=1.0: (torch.randn(*sh, device=dev, generator=g) * sc).bfloat16()
    x0, attn, cz = rnd(M, D), rnd(M, INNER), rnd(M, HZ)
    Wo, Wxq, Wxo, Wwi, Wwo, Wqkv = rnd(D, INNER, sc=.05), rnd(HZ, D, sc=.05), rnd(D, HZ, sc=.03), rnd(2 * F, D, sc=.05), rnd(D, F, sc=.03), rnd(3 * INNER, D, sc=.05)
    s_ = torch.cuda.current_stream().cuda_stream
    ncnt = int(lib.ymt3_op_linear_chain_counters(M))

    def mk():
        return dict(x=x0.clone(), qz=torch.zeros(M, HZ, device=dev, dtype=torch.bfloat16), gbuf=torch.zeros(M, F, device=dev, dtype=torch.bfloat16),
                    qkv=torch.zeros(M, 3 * INNER, device=dev, dtype=torch.bfloat16), sA=torch.zeros(M, 16, device=dev), sB=torch.zeros(M, 16, device=dev),
                    sC=torch.zeros(M, 16, device=dev))

    def phases(b, which):
        P = _lib.ChainPhase
        def ph(A, lda, W, C_, ldc, N, K, res=False, ss_in=None, ss_out=None, act=0, gated=0):
            return P(A.data_ptr(), lda, W.data_ptr(), K, None, ss_in.data_ptr() if ss_in is not None else None, 16, 1e-6,
                     C_.data_ptr(), ldc, C_.data_ptr() if res else None, ldc, ss_out.data_ptr() if ss_out is not None else None, N, K, act, gated, 1.0)
        if which == "A":
            return [ph(attn, INNER, Wo, b["x"], D, D, INNER, res=True, ss_out=b["sB"]),
                    ph(b["x"], D, Wxq, b["qz"], HZ, HZ, D, ss_in=b["sB"])]
        return [ph(cz, HZ, Wxo, b["x"], D, D, HZ, res=True, ss_out=b["sC"]),
                ph(b["x"], D, Wwi, b["gbuf"], F, 2 * F, D, ss_in=b["sC"], act=1, gated=1),
                ph(b["gbuf"], F, Wwo, b["x"], D, D, F, res=True, ss_out=b["sA"]),
                ph(b["x"], D, Wqkv, b["qkv"], 3 * INNER, 3 * INNER, D, ss_in=b["sA"])]

    def run_separate(ps):
        for q in ps:
            _lib.check(lib.ymt3_op_linear_normfused(q.A, q.lda, q.W, q.ldw, None, q.ss_in, q.chunks if q.ss_in else 0, q.eps, q.C, q.ldc,
                                                    q.residual, q.ldr, q.ss_out, M, q.N, q.K, q.act, q.gated, 1.0, 1, s_))

    for which in ("A", "B"):
        b1, b2 = mk(), mk()
        p1, p2 = phases(b1, which), phases(b2, which)
        arr = (_lib.ChainPhase * len(p2))(*p2)
        cnt = torch.zeros(ncnt, device=dev, dtype=torch.int32)
        ordinal = [0]

        def run_chain():
            _lib.check(lib.ymt3_op_linear_chain(arr, len(p2), M, cnt.data_ptr(), ordinal[0], s_))
            ordinal[0] += 1
        run_separate(p1)
        run_chain()
        torch.cuda.synchronize()
        same = all(torch.equal(b1[k], b2[k]) for k in b1)
        t_sep = timeit(lambda: run_separate(p1))
        t_ch = timeit(run_chain)
        print(f"chain {which} M={M}: separate {t_sep:7.1f} us   chained {t_ch:7.1f} us   identical={same}", flush=True)

if what in ("decode", "all"):
    # the GEMMs of one YPTF.MoE+Multi decode step at B=256 (N = 3328 sequences)
    M = 3328
    gemm(M, 1152, 512, name="dec self qkv")
    gemm(M, 512, 384, residual=True, name="dec self o+res")
    gemm(M, 384, 512, name="dec cross q (K/V mode)")
    gemm(M, 1536, 512, name="dec cross q absorbed")
    gemm(M, 512, 1536, residual=True, name="dec cross o absorbed+res")
    gemm(M, 2048, 512, gated=1, act=1, name="dec ffn wi")
    gemm(M, 512, 1024, residual=True, name="dec ffn wo+res")
    gemm(M, 600, 512, name="dec lm head")
    for (N, H, T) in [(3328, 6, 110), (832, 6, 110)]:
        Tp = (T + 15) // 16 * 16
        q = torch.randn(N, H * 256, device=dev).bfloat16()
        z = torch.zeros(N, Tp, 256, device=dev, dtype=torch.bfloat16)
        z[:, :T] = torch.randn(N, T, 256, device=dev).bfloat16()
        o = torch.empty_like(q)
        s_ = torch.cuda.current_stream().cuda_stream
        us = timeit(lambda: _lib.check(lib.ymt3_op_cross_attn_absorbed(q.data_ptr(), z.data_ptr(), o.data_ptr(), N, H, T, Tp, s_)))
        byts = z.numel() * 2 + 2 * q.numel() * 2
        print(f"cross_attn_absorbed N={N} H={H} T={T}: {us:8.1f} us  {byts / us / 1e3:8.1f} GB/s (z + q + out bytes)", flush=True)

if what == "xattn":
    # absorbed cross-attention alone at the B = 256 / 728 shapes of YPTF.MoE+Multi
    for (N, H, T) in [(3328, 6, 110), (9464, 6, 110)]:
        Tp = (T + 15) // 16 * 16
        q = torch.randn(N, H * 256, device=dev).bfloat16()
        z = torch.zeros(N, Tp, 256, device=dev, dtype=torch.bfloat16)
        z[:, :T] = torch.randn(N, T, 256, device=dev).bfloat16()
        o = torch.empty_like(q)
        s_ = torch.cuda.current_stream().cuda_stream
        us = timeit(lambda: _lib.check(lib.ymt3_op_cross_attn_absorbed(q.data_ptr(), z.data_ptr(), o.data_ptr(), N, H, T, Tp, s_)))
        byts = z.numel() * 2 + 2 * q.numel() * 2
        print(f"cross_attn_absorbed N={N} H={H} T={T}: {us:8.1f} us  {byts / us / 1e3:8.1f} GB/s (z + q + out bytes)", flush=True)

if what in ("decode_attn", "all"):
    # self-attention over a bf16 KV cache (decode.cu) at the YPTF.MoE+Multi B=256 shape, cache length L
    N, H, Lcap = 3328, 6, 256
    q = torch.randn(N, H * 64, device=dev).bfloat16()
    kn, vn, o = torch.randn_like(q), torch.randn_like(q), torch.empty_like(q)
    Kc = torch.randn(N, H, Lcap, 64, device=dev).bfloat16()
    Vc = torch.randn(N, H, Lcap, 64, device=dev).bfloat16()
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    s_ = torch.cuda.current_stream().cuda_stream
    for L in ([128] if once else [1, 4, 16, 32, 64, 128, 192, 256]):
        st.fill_(L - 1)
        us = timeit(lambda: _lib.check(lib.ymt3_op_decode_attention(1, q.data_ptr(), kn.data_ptr(), vn.data_ptr(), Kc.data_ptr(),
                                                                    Vc.data_ptr(), st.data_ptr(), 0, o.data_ptr(), N, H, Lcap, s_)))
        byts = N * H * 64 * 2 * (2 * L + 6)
        print(f"decode_attn N={N} H={H} len={L}: {us:8.1f} us  {byts / us / 1e3:8.1f} GB/s (algorithmic)", flush=True)
