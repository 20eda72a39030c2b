"""MoE feed-forward through ymt3_op_moe_ff: fused expert kernel vs the two grouped GEMMs (YMT3_NO_MOE_FUSED) at several
token counts (CUDA events, median of 20).  usage: python tools/bench_moe.py [N ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from yourmt3_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
D, I, E, topk = 128, 512, 8, 2
g = torch.Generator().manual_seed(0)
gate = (torch.randn(E, D, generator=g) * 0.5).to(dev)
w13 = (torch.randn(E, 2 * I, D, generator=g) * 0.05).to(dev, torch.bfloat16)
w2 = (torch.randn(E, D, I, generator=g) * 0.05).to(dev, torch.bfloat16)
for N in [int(a) for a in sys.argv[1:]] or [2860, 22880, 183040, 2082080]:
    x = torch.randn(N, D, generator=g).to(dev, torch.bfloat16)
    out = torch.empty_like(x)
    ws = torch.empty(max(1, lib.ymt3_op_moe_workspace_bytes(N, D, I, E, topk, 1)), dtype=torch.uint8, device=dev)
    s_ = torch.cuda.current_stream().cuda_stream
    res = {}
    for mode in ("fused", "grouped"):
        if mode == "grouped":
            os.environ["YMT3_NO_MOE_FUSED"] = "1"
        else:
            os.environ.pop("YMT3_NO_MOE_FUSED", None)

        def run():
            _lib.check(lib.ymt3_op_moe_ff(1, x.data_ptr(), x.data_ptr(), out.data_ptr(), N, gate.data_ptr(), w13.data_ptr(),
                                          w2.data_ptr(), D, I, E, topk, 3, ws.data_ptr(), s_))
        for _ in range(3):
            run()
        ts = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        res[mode] = ts[len(ts) // 2]
    os.environ.pop("YMT3_NO_MOE_FUSED", None)
    print(f"moe_ff N={N:8d} tokens (route + sort + experts + combine): fused {res['fused']:9.1f} us   grouped GEMMs {res['grouped']:9.1f} us", flush=True)
