"""Phase timing of one inference batch with CUDA events: frontend+encoder vs the decode loop, and the in-graph cost of
each kernel family of the decode step by leaving it out (YMT3_DEBUG_SKIP, profiling aid in csrc/t5.cu).
usage: python tools/time_phases.py [PRESET] [BATCH]      (spawns itself once per skip mask)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(preset, batch):
    import torch
    import yourmt3_b200 as ymt3
    audio = {"codec": "spec", "hop_length": 300} if preset.startswith("yptf") else {}
    m = ymt3.init_nondegenerate_(ymt3.YourMT3(audio_cfg=ymt3.get_audio_cfg(**audio), model_cfg=ymt3.get_model_cfg(preset),
                                              precision="bf16"), 0).cuda()
    x = torch.randn(batch, 1, 32767, device="cuda") * 0.1
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for rep in range(3):
        ev[0].record()
        if m._absorbed():
            lat = m.encoder(inputs_embeds=m.pre_encoder(m.spectrogram(x)))["last_hidden_state"]
            ev[1].record()
            m._generate_absorbed(lat, None, m.max_token_length, stop_at_eos=False)
        else:
            enc = m.encode(x)
            ev[1].record()
            from yourmt3_b200.t5mod_helper import task_cond_dec_generate
            task_cond_dec_generate(m.decoder, m.decoder_type, m.embed_tokens, m.lm_head, enc, max_length=m.max_token_length,
                                   stop_at_eos=False, precision=m._prec)
        ev[2].record()
        torch.cuda.synchronize()
    print(f"{ev[0].elapsed_time(ev[1]):9.2f} {ev[1].elapsed_time(ev[2]):9.2f}", flush=True)


if len(sys.argv) > 3 and sys.argv[3] == "--worker":
    worker(sys.argv[1], int(sys.argv[2]))
else:
    preset = sys.argv[1] if len(sys.argv) > 1 else "yptf_moe_multi"
    batch = sys.argv[2] if len(sys.argv) > 2 else "256"
    names = {0: "full step", 1: "without self-attention", 2: "without cross-attention kernel", 4: "without layer GEMMs",
             8: "without norms", 15: "embed + lm head + greedy only"}
    base = None
    print(f"# {preset} B={batch}: encode ms / decode-loop ms (third repetition), decode steps = model event_length")
    for mask, name in names.items():
        env = dict(os.environ)
        if mask:
            env["YMT3_DEBUG_SKIP"] = str(mask)
        out = subprocess.run([sys.executable, __file__, preset, batch, "--worker"], env=env, capture_output=True, text=True)
        if out.returncode:
            print(name, "FAILED", out.stderr[-400:])
            continue
        enc, dec = (float(v) for v in out.stdout.split()[-2:])
        base = dec if base is None else base
        print(f"{name:34s} encode {enc:8.2f} ms   decode {dec:8.2f} ms   (family cost {base - dec:8.2f} ms)", flush=True)
