"""The "bar to beat" of SURVEY.md 2c: the reference's own PyTorch path (torchaudio / eager torch
modules) run on the SAME B200, next to this repo's kernels.  Prints a markdown table.

  python tools/ref_gpu_bar.py frontend      # log-mel batch sweep 1..4096 vs torchaudio on cuda
  python tools/ref_gpu_bar.py t5            # T5-small greedy decode (eager torch oracle on cuda) vs native
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import yourmt3_b200 as ymt3  # noqa: E402

dev = torch.device("cuda")


def cuda_time(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def frontend():
    import torchaudio
    ms = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048, hop_length=128, f_min=50.0, f_max=8000.0,
                                              n_mels=512, power=1.0).to(dev)
    layer, _ = ymt3.get_spectrogram_layer_from_audio_cfg(ymt3.get_audio_cfg())
    print("| B | torchaudio cuda ms | native ms | speed-up | native audio-s/s | native GB/s (algorithmic) |")
    print("|---|---|---|---|---|---|")
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
        x = torch.randn(B, 32767, device=dev) * 0.1
        ref = cuda_time(lambda: torch.log(torch.clamp(ms(x), min=1e-5)).transpose(1, 2).contiguous(), reps=5 if B > 512 else 10)
        nat = cuda_time(lambda: layer(x))
        gbs = B * 655356 / (nat * 1e-3) / 1e9
        print(f"| {B} | {ref:.3f} | {nat:.3f} | {ref / nat:.1f}x | {B * 2.047937 / (nat * 1e-3):,.0f} | {gbs:.0f} |", flush=True)
        del x
        torch.cuda.empty_cache()


def t5():
    from oracle import pipeline as OP
    B, steps = 64, 64
    m = ymt3.YourMT3(precision="bf16")
    ymt3.init_nondegenerate_(m, 0)
    m = m.to(dev)
    sd = {k: v.detach().float() for k, v in m.state_dict().items()}
    x = torch.randn(B, 1, 32767, device=dev) * 0.1
    n_pos = m.decoder.pos_table.shape[0]

    def ref():
        with torch.no_grad():
            feats = m.spectrogram(x)                       # frontend excluded from the comparison (same kernel)
            enc = OP.OT.t5_encoder({k[8:]: v for k, v in sd.items() if k.startswith("encoder.")}, feats, n_layers=8,
                                   n_heads=6, pos=OP.OT.sinusoidal_positions(n_pos, 512).to(dev))
            return OP.OT.greedy_generate({k[8:]: v for k, v in sd.items() if k.startswith("decoder.")}, enc,
                                         embed=sd["embed_tokens.weight"], lm_head=sd["embed_tokens.weight"], n_layers=8,
                                         n_heads=6, max_length=steps, prefix="", pos=OP.OT.sinusoidal_positions(n_pos, 512).to(dev),
                                         stop_at_eos=False)
    t_ref = cuda_time(ref, reps=3, warm=1)
    t_nat = cuda_time(lambda: m.inference(x, max_token_length=steps, stop_at_eos=False), reps=5, warm=2)
    print(f"T5-small, B={B}, {steps} decode steps: eager torch fp32 on cuda {t_ref:.1f} ms, native bf16 {t_nat:.1f} ms "
          f"-> {t_ref / t_nat:.1f}x")
    m32 = ymt3.init_nondegenerate_(ymt3.YourMT3(precision="f32"), 0).to(dev)
    t_n32 = cuda_time(lambda: m32.inference(x, max_token_length=steps, stop_at_eos=False), reps=3, warm=1)
    print(f"                                  native fp32 (exact path) {t_n32:.1f} ms -> {t_ref / t_n32:.1f}x")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "frontend"
    {"frontend": frontend, "t5": t5}[which]()
