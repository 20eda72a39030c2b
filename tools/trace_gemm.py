"""Per-CTA timeline of ONE tcgen05 GEMM launch (gemm_bf16_tc_kernel) - needs the library built with -DYMT3_GEMM_TRACE
(YMT3_EXTRA_NVCC_FLAGS=-DYMT3_GEMM_TRACE python -m yourmt3_b200.build; rebuild without it afterwards).
usage: python tools/trace_gemm.py M N K [res] [gated]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from yourmt3_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
M, N, K = (int(v) for v in sys.argv[1:4])
res, gated = "res" in sys.argv, "gated" in sys.argv
A = torch.randn(M, K, device=dev).bfloat16()
W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
No = N // 2 if gated else N
C = torch.randn(M, No, device=dev).bfloat16()
s_ = torch.cuda.current_stream().cuda_stream
n_cta = torch.cuda.get_device_properties(0).multi_processor_count


def run():
    _lib.check(lib.ymt3_op_linear(1, A.data_ptr(), K, W.data_ptr(), K, None, C.data_ptr(), No, C.data_ptr() if res else None, No,
                                  M, N, K, 1 if gated else 0, int(gated), 1.0, 1, s_))


for _ in range(5):
    run()
trace = torch.zeros(n_cta * 16 * 32, device=dev, dtype=torch.int64)
_lib.check(lib.ymt3_debug_chain_trace(trace.data_ptr()))
run()
torch.cuda.synchronize()
_lib.check(lib.ymt3_debug_chain_trace(None))
raw = trace.cpu().numpy().reshape(n_cta, 16, 32).astype(np.float64)
raw[raw == 0] = np.nan
t = raw.copy()
t0 = np.nanmin(t[:, 0, 7])
t = (t - t0) / 1e3
print(f"gemm M={M} N={N} K={K} res={res} gated={gated}: prologue done at median {np.nanmedian(t[:, 0, 7]):.2f} us (max {np.nanmax(t[:, 0, 7]):.2f}); "
      f"all stores complete at median {np.nanmedian(t[:, 0, 6]):.2f} (max {np.nanmax(t[:, 0, 6]):.2f})")
print("tile  t_data t_issued   t_acc   t_epi   (us, median over CTAs)")
for j in range(16):
    if np.all(np.isnan(t[:, j, 2:6])):
        break
    print(f"{j:3d}  " + " ".join(f"{np.nanmedian(t[:, j, e]):7.2f}" for e in (2, 3, 4, 5)))
print("epilogue of warp 2, SM cycles (general path: per 32-column chunk: hand-over to TMA | accumulator load | math | box wait | "
      "residual + pack + store; lean path: first four 16-column halves: box hand-over | TMEM wait | next load + prefetch + release | "
      "box wait | math + staging) || total   (median over CTAs)")
for j in range(16):
    if np.all(np.isnan(t[:, j, 2:6])):
        break
    c = raw[:, j, 8:32].reshape(n_cta, 4, 6)
    d = np.diff(c, axis=2)
    tot = c[:, :, 5] - c[:, :, 0]
    print(f"   {j:2d}  " + "   ".join(" ".join(f"{np.nanmedian(d[:, k, e]):5.0f}" for e in range(5)) + f" ||{np.nanmedian(tot[:, k]):5.0f}"
                                     for k in range(4) if not np.all(np.isnan(tot[:, k]))))
