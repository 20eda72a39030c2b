"""Per-CTA timeline of the chained decode-step GEMMs (gemm_chain_kernel, debug stamps through ymt3_debug_chain_trace).
usage: python tools/trace_chain.py [A|B] [M]      prints, per tile of a CTA's list, the median over CTAs of the event times
(us after the first CTA passed its prologue):
  t_begin  producer reaches the tile      t_dep   dependency flag seen        t_data  first ring stage landed (MMA starts)
  t_issued last MMA of the tile issued    t_acc   accumulator complete        t_epi   tile's epilogue done (stores issued)
  t_pub    published (stores complete + fence + counter), only on a CTA's last N tile of a phase"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from yourmt3_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
which = sys.argv[1] if len(sys.argv) > 1 else "B"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 9464
D, INNER, HZ, F = 512, 384, 1536, 1024
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *sh, sc=1.0: (torch.randn(*sh, device=dev, generator=g) * sc).bfloat16()  # noqa: E731
attn, cz = rnd(M, INNER), rnd(M, HZ)
Wo, Wxq, Wxo, Wwi, Wwo, Wqkv = (rnd(D, INNER, sc=.05), rnd(HZ, D, sc=.05), rnd(D, HZ, sc=.03), rnd(2 * F, D, sc=.05),
                                rnd(D, F, sc=.03), rnd(3 * INNER, D, sc=.05))
x = rnd(M, D)
qz, gbuf, qkv = (torch.empty(M, n, device=dev, dtype=torch.bfloat16) for n in (HZ, F, 3 * INNER))
sA, sB, sC = (torch.zeros(M, 16, device=dev) for _ in range(3))
P = _lib.ChainPhase


def ph(A, lda, W, C_, ldc, N, K, res=False, ss_in=None, ss_out=None, act=0, gated=0):
    return P(A.data_ptr(), lda, W.data_ptr(), K, None, ss_in.data_ptr() if ss_in is not None else None, 16, 1e-6,
             C_.data_ptr(), ldc, C_.data_ptr() if res else None, ldc, ss_out.data_ptr() if ss_out is not None else None,
             N, K, act, gated, 1.0)


if which == "A":
    ps = [ph(attn, INNER, Wo, x, D, D, INNER, res=True, ss_out=sB), ph(x, D, Wxq, qz, HZ, HZ, D, ss_in=sB)]
    names = ["o+res", "xq"]
else:
    ps = [ph(cz, HZ, Wxo, x, D, D, HZ, res=True, ss_out=sC), ph(x, D, Wwi, gbuf, F, 2 * F, D, ss_in=sC, act=1, gated=1),
          ph(gbuf, F, Wwo, x, D, D, F, res=True, ss_out=sA), ph(x, D, Wqkv, qkv, 3 * INNER, 3 * INNER, D, ss_in=sA)]
    names = ["xo+res", "wi", "wo+res", "qkv"]
arr = (P * len(ps))(*ps)
cnt = torch.zeros(int(lib.ymt3_op_linear_chain_counters(M)), device=dev, dtype=torch.int32)
s_ = torch.cuda.current_stream().cuda_stream
n_cta = torch.cuda.get_device_properties(0).multi_processor_count
for i in range(5):
    _lib.check(lib.ymt3_op_linear_chain(arr, len(ps), M, cnt.data_ptr(), i, s_))
trace = torch.zeros(n_cta * 16 * 32, device=dev, dtype=torch.int64)
_lib.check(lib.ymt3_debug_chain_trace(trace.data_ptr()))
_lib.check(lib.ymt3_op_linear_chain(arr, len(ps), M, cnt.data_ptr(), 5, s_))
torch.cuda.synchronize()
_lib.check(lib.ymt3_debug_chain_trace(None))
t = trace.cpu().numpy().reshape(n_cta, 16, 32).astype(np.float64)
t[t == 0] = np.nan
t0 = np.nanmin(t[:, 0, 7])
t = (t - t0) / 1e3
print(f"chain {which}, M = {M}: phases {names}; CTA prologue done at median {np.nanmedian(t[:, 0, 7]):.2f} us (max {np.nanmax(t[:, 0, 7]):.2f})")
print("tile   t_begin   t_dep  t_data t_issued   t_acc   t_epi   t_pub   (median over CTAs, us; lane 0 / lane 1 of the pairs)")
for lane in (0, 1):
    sel = t[lane::2] if os.environ.get("YMT3_GEMM_CHAIN_SPLIT", "2") == "2" else t
    for j in range(16):
        if np.all(np.isnan(sel[:, j, :7])):
            break
        med = [np.nanmedian(sel[:, j, e]) if not np.all(np.isnan(sel[:, j, e])) else float("nan") for e in range(7)]
        print(f"L{lane} {j:2d}  " + " ".join(f"{v:7.2f}" for v in med))
raw = trace.cpu().numpy().reshape(n_cta, 16, 32).astype(np.float64)
raw[raw == 0] = np.nan
print("epilogue of warp 2, SM cycles per 32-column chunk: hand-over to TMA | accumulator load | math | box wait | residual + pack + "
      "store || chunk total   (median over CTAs)")
for j in range(16):
    if np.all(np.isnan(t[:, j, :7])):
        break
    c = raw[:, j, 8:32].reshape(n_cta, 4, 6)
    d = np.diff(c, axis=2)
    tot = c[:, :, 5] - c[:, :, 0]
    gap = c[:, 1:, 0] - c[:, :-1, 5]
    print(f"   {j:2d}  " + "   ".join(" ".join(f"{np.nanmedian(d[:, k, e]):5.0f}" for e in range(5)) + f" ||{np.nanmedian(tot[:, k]):5.0f}"
                                     for k in range(4)) + f"   tile {np.nanmedian(c[:, 3, 5] - c[:, 0, 0]):6.0f}")
print(f"last event at {np.nanmax(t[:, :, :7]):.2f} us")
