"""HBM bandwidth by access mix on this GPU (context for write-heavy kernels): fill (write only), sum (read only),
copy (read + write), torch kernels, CUDA events, best of 10 over 2 GiB buffers."""
import torch

n = 1 << 30   # bf16 elements = 2 GiB
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty_like(a)
a.fill_(1.0)


def best(fn, byts):
    ts = []
    for _ in range(12):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    t = min(ts[2:])
    return byts / t / 1e6


print(f"fill (write only): {best(lambda: a.fill_(0.5), 2 * n):8.1f} GB/s")
print(f"sum  (read only) : {best(lambda: a.view(torch.int32).sum(), 2 * n):8.1f} GB/s")
print(f"copy (read+write): {best(lambda: b.copy_(a), 4 * n):8.1f} GB/s")
