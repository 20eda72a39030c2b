"""Model part of ``__graft_entry__.smoke()``: one tiny YPTF.MoE+Multi transcription on cuda:0 through every
kernel family (log-spectrogram, conv pre-encoder, Perceiver-TF with MoE + RoPE, projection, multi-channel greedy
decode) checked against the CPU oracle (fp32 tokens identical), then the same model on the bf16 / tcgen05 path."""
import numpy as np
import torch

import yourmt3_b200 as ymt3


def _small(precision, seed=5):
    cfg = ymt3.get_model_cfg("yptf_moe_multi")
    cfg["encoder"]["perceiver-tf"]["num_blocks"] = 1
    cfg["decoder"]["multi-t5"]["num_layers"] = 2
    cfg["event_length"] = 8
    m = ymt3.YourMT3(audio_cfg=ymt3.get_audio_cfg(codec="spec", hop_length=300), model_cfg=cfg, precision=precision)
    return ymt3.init_nondegenerate_(m, seed)


def run(dev):
    from oracle import pipeline as OP
    g = torch.Generator().manual_seed(21)
    audio = (torch.randn(2, 32767, generator=g) * 0.1).numpy().astype(np.float32)
    m = _small("f32").to(dev)
    got = m.inference(torch.from_numpy(audio).unsqueeze(1).to(dev), stop_at_eos=False).cpu().numpy().reshape(26, 8)
    ref, margins = OP.transcribe(m.state_dict(), audio, m.audio_cfg, m.model_cfg, n_pos=m.decoder.pos_table.shape[0],
                                 max_length=8, stop_at_eos=False, return_margins=True)
    ref = ref.numpy()
    bad = [(n, int(np.argmax(got[n] != ref[n]))) for n in range(26) if (got[n] != ref[n]).any()]
    bad = [(n, t) for n, t in bad if float(margins[n, t]) >= 1e-4]
    assert not bad, f"fp32 tokens differ from the oracle at (row, step) {bad}"
    print(f"smoke yptf_moe_multi fp32: {got.shape} tokens identical to the CPU oracle, {len(np.unique(got))} distinct ids")
    m16 = _small("bf16").to(dev)
    t16 = m16.inference(torch.from_numpy(audio).unsqueeze(1).to(dev), stop_at_eos=False).cpu().numpy().reshape(26, 8)
    agree = float((t16[:, 0] == got[:, 0]).mean())
    print(f"smoke yptf_moe_multi bf16 (tcgen05 GEMMs / implicit-GEMM convs, fused RMSNorm, absorbed cross-attention="
          f"{m16._absorbed()}): first-token agreement with fp32 {agree:.2f}")
    assert agree >= 0.6
    # T5-small shape, 1 layer each, bf16
    cfg = ymt3.get_model_cfg("mt3_t5_small")
    cfg["encoder"]["t5"]["num_layers"] = cfg["decoder"]["t5"]["num_layers"] = 1
    cfg["event_length"] = 8
    t5 = ymt3.init_nondegenerate_(ymt3.YourMT3(model_cfg=cfg, precision="bf16"), 0).to(dev)
    tok = t5.inference(torch.from_numpy(audio).unsqueeze(1).to(dev), stop_at_eos=False)
    assert tok.shape == (2, 8)
    print("smoke t5_small bf16:", tok[0].tolist())
