"""How much time do small kernels cost INSIDE a CUDA graph (launch gaps, smem carve-out switches)?
Captures chains of decode-step-like kernels through the C ABI and times graph replays."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from yourmt3_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 3328
x = torch.randn(M, 512, device=dev).bfloat16()
h = torch.empty_like(x)
w_ln = torch.ones(512, device=dev)
Wqkv = (torch.randn(1152, 512, device=dev) * 0.05).bfloat16()
Wo = (torch.randn(512, 384, device=dev) * 0.05).bfloat16()
qkv = torch.empty(M, 1152, device=dev, dtype=torch.bfloat16)


def norm(s):
    _lib.check(lib.ymt3_op_rmsnorm(1, x.data_ptr(), w_ln.data_ptr(), h.data_ptr(), M, 512, 1e-6, s))


def gemm_qkv(s):
    _lib.check(lib.ymt3_op_linear(1, h.data_ptr(), 512, Wqkv.data_ptr(), 512, None, qkv.data_ptr(), 1152, None, 0, M, 1152,
                                  512, 0, 0, 1.0, 1, s))


def gemm_o(s):
    _lib.check(lib.ymt3_op_linear(1, qkv.data_ptr(), 1152, Wo.data_ptr(), 384, None, x.data_ptr(), 512, x.data_ptr(), 512, M,
                                  512, 384, 0, 0, 1.0, 1, s))


def time_graph(name, seq, reps=50):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for f in seq:
            f(st.cuda_stream)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(20):
                for f in seq:
                    f(st.cuda_stream)
        for _ in range(3):
            g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(reps):
            g.replay()
        b.record(st)
        b.synchronize()
    per = a.elapsed_time(b) * 1e3 / (reps * 20 * len(seq))
    print(f"{name:40s} {per:7.2f} us per kernel (in graph, M={M})", flush=True)


time_graph("rmsnorm only", [norm])
time_graph("gemm qkv only", [gemm_qkv])
time_graph("gemm o-proj(+res) only", [gemm_o])
time_graph("norm -> qkv -> o (alternating)", [norm, gemm_qkv, gemm_o])
time_graph("qkv -> o (alternating gemms)", [gemm_qkv, gemm_o])
