"""Diagnostic: teacher-forced logit error of the bf16 path vs the fp32 oracle, per step (GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests import test_fulldepth_gpu as T  # noqa: E402

dev = torch.device("cuda")
for preset, L, rand in (("mt3_t5_small", 256, False), ("mt3_t5_small", 1024, True), ("yptf_moe_multi", 256, False), ("yptf", 256, False)):
    audio = T._audio((101, 104) if not rand else (101,))
    m32 = T._model(preset, "f32", dev)
    if rand:
        ref = torch.randint(3, m32.vocab_size, (1, L), generator=torch.Generator().manual_seed(5))
        margins = None
    else:
        ref, margins, _ = T._oracle_tokens(m32, audio, L)
    ref_logits = T._oracle_teacher_forced_logits(m32, audio, ref)
    rng = float(ref_logits.max() - ref_logits.min())
    x = torch.from_numpy(audio).unsqueeze(1).to(dev)
    shape = ref.shape if m32.decoder_type != "multi-t5" else (audio.shape[0], 13, L)
    ac = T._oracle_teacher_forced_logits(m32, audio, ref, device="cuda", autocast=True)
    e_ac = (ac - ref_logits).abs().amax(-1) / rng
    print(f"{preset} L={L} torch autocast bf16 (eager oracle modules on cuda): err/range max {float(e_ac.max()):.5f} median "
          f"{float(e_ac.median()):.5f} p99 {float(e_ac.flatten().quantile(0.99)):.5f}; agree "
          f"{float((ac.argmax(-1) == ref_logits.argmax(-1)).float().mean()):.4f}")
    for prec in ("f32", "bf16"):
        m = m32 if prec == "f32" else T._model(preset, "bf16", dev)
        am, lg = m.score(x, ref.view(shape).to(dev), logit_steps=list(range(L)))
        lg = lg.permute(1, 0, 2).cpu()
        err = (lg - ref_logits).abs().amax(-1) / rng           # (N, L)
        agree = (am.reshape(-1, L).cpu() == ref_logits.argmax(-1)).float()
        steps = [s for s in (0, 1, 3, 7, 15, 31, 63, 127, 255, 511, 1023) if s < L]
        print(f"{preset} L={L} {prec}: range {rng:.2f}; err/range max {float(err.max()):.5f} median {float(err.median()):.5f} "
              f"p99 {float(err.flatten().quantile(0.99)):.5f}; agree {float(agree.mean()):.4f}")
        print("   per step max:", {s: round(float(err[:, s].max()), 5) for s in steps})
        print("   per step mean agree (windows of 32):", [round(float(agree[:, i:i + 32].mean()), 3) for i in range(0, L, max(32, L // 8))])
        if margins is not None:
            print(f"   oracle margin/range: median {float(margins.median()) / rng:.4f} p10 {float(margins.flatten().quantile(0.1)) / rng:.5f}")
        # row-level view: which rows carry the max
        print("   per row max:", [round(float(v), 4) for v in err.amax(1)[:13]])
