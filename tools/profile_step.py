"""Small driver for ncu: run a few decode steps (and optionally the encoder) of a preset so the
per-kernel launch list can be captured.  usage: python tools/profile_step.py PRESET BATCH STEPS [precision] [runs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import yourmt3_b200 as ymt3  # noqa: E402

preset, batch, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
precision = sys.argv[4] if len(sys.argv) > 4 else "bf16"
lanes = int(os.environ.get("YMT3_LANES", "1"))
audio = {"codec": "spec", "hop_length": 300} if preset.startswith("yptf") else {}
m = ymt3.YourMT3(audio_cfg=ymt3.get_audio_cfg(**audio), model_cfg=ymt3.get_model_cfg(preset), precision=precision)
ymt3.init_nondegenerate_(m, 0)
m = m.cuda()
m.decode_lanes = lanes
x = torch.randn(batch, 1, 32767, device="cuda") * 0.1
runs = int(sys.argv[5]) if len(sys.argv) > 5 else 2
for _ in range(runs):
    t = m.inference(x, max_token_length=steps, stop_at_eos=False)
torch.cuda.synchronize()
print("ok", tuple(t.shape))
