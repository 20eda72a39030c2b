"""same-process timing of the res3b conv pre-encoder and the Perceiver-TF encoder (bf16) - run once per
YMT3_GEMM_CLUSTER setting (the switch is read once per process)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import yourmt3_b200 as ymt3
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = ymt3.init_nondegenerate_(ymt3.YourMT3(audio_cfg=ymt3.get_audio_cfg(codec="spec", hop_length=300),
                                          model_cfg=ymt3.get_model_cfg("yptf_moe_multi"), precision="bf16"), 0).to(dev)
x = torch.randn(B, 1, 32767, device=dev) * 0.1
feats = m.spectrogram(x)
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
pre = m.pre_encoder(feats)
print(f"cluster={os.environ.get('YMT3_GEMM_CLUSTER', '1')} B={B}: pre-encoder {t(lambda: m.pre_encoder(feats)):.2f} ms, "
      f"perceiver-tf encoder {t(lambda: m.encoder(inputs_embeds=pre)):.2f} ms", flush=True)
