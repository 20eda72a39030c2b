"""ONE launch shape of the log-mel kernel (for ncu).  usage: python tools/run_logmel.py [spec|mel] [B] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import yourmt3_b200 as ymt3  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "spec"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 728
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = ymt3.get_audio_cfg(codec="spec", hop_length=300) if kind == "spec" else ymt3.get_audio_cfg()
layer, (T, F) = ymt3.get_spectrogram_layer_from_audio_cfg(cfg)
x = torch.randn(B, 1, 32767, device="cuda") * 0.1
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    flush.zero_()
    y = layer(x)
torch.cuda.synchronize()
print("ok", kind, B, tuple(y.shape))
