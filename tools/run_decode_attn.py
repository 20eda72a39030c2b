"""ONE shape of the decode self-attention kernel through the C ABI (for ncu captures at the bench shape).
usage: python tools/run_decode_attn.py N H LEN [LCAP] [reps]   (bf16 cache; default bench shape 9464 6 128 256)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from yourmt3_b200 import _lib  # noqa: E402

N, H, ln = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (9464, 6, 128)))
Lcap = int(sys.argv[4]) if len(sys.argv) > 4 else 256
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
lib, dev = _lib.load(), torch.device("cuda")
q = torch.randn(N, H * 64, device=dev).bfloat16()
kn, vn, out = torch.randn_like(q), torch.randn_like(q), torch.empty_like(q)
Kc = torch.randn(N, H, Lcap, 64, device=dev).bfloat16()
Vc = torch.randn(N, H, Lcap, 64, device=dev).bfloat16()
step = torch.full((1,), ln - 1, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(reps):
    flush.zero_()                                            # L2 flush between launches
    _lib.check(lib.ymt3_op_decode_attention(_lib.DTYPE_BF16, q.data_ptr(), kn.data_ptr(), vn.data_ptr(), Kc.data_ptr(),
                                            Vc.data_ptr(), step.data_ptr(), 0, out.data_ptr(), N, H, Lcap,
                                            _lib.current_stream_ptr()), "decode_attention")
torch.cuda.synchronize()
print("ok", N, H, ln, "algorithmic bytes per launch", N * H * 64 * 2 * (2 * ln + 6))
