"""tests/golden/t5_hf_small.npz was produced by the installed transformers T5Stack (tools/make_golden_t5.py):
encoder hidden states and a 24-step greedy decode with KV cache and tied LM head.  CPU tier: the oracle reproduces it.
GPU tier: the native fp32 path reproduces it THROUGH THE C ABI without the oracle in the loop (tokens identical)."""
import os

import numpy as np
import pytest
import torch
from torch import nn

import yourmt3_b200 as ymt3
from oracle import t5 as OT
from tests.util import GOLDEN_DIR, fill_by_name_
from yourmt3_b200.t5mod import sinusoidal_positions

G = np.load(os.path.join(GOLDEN_DIR, "t5_hf_small.npz"))
D, H, DKV, DFF, NL, V, L, NPOS = (int(v) for v in G["dims"])
STD, EMBED_STD = (float(v) for v in G["stds"])
CFG = {"d_model": D, "num_heads": H, "d_kv": DKV, "num_layers": NL, "ff_widening_factor": DFF // D,
       "position_encoding_type": "sinusoidal", "layer_norm_epsilon": 1e-6, "vocab_size": V}


def _modules():
    enc = ymt3.T5EncoderYMT3(CFG, precision="f32", num_max_positions=NPOS)
    dec = ymt3.T5DecoderYMT3(CFG, num_max_positions=NPOS)
    fill_by_name_(list(enc.named_parameters()), std=STD)
    fill_by_name_(list(dec.named_parameters()), std=STD)
    emb = nn.Embedding(V, D)
    fill_by_name_([("embed_tokens.weight", emb.weight)], std=EMBED_STD)
    head = ymt3.LMHead(CFG, 1.0, True)
    head.lm_head.weight = emb.weight
    return enc, dec, emb, head


def test_oracle_reproduces_hf_golden():
    enc, dec, emb, _ = _modules()
    pos = sinusoidal_positions(NPOS, D)
    with torch.no_grad():
        x = torch.from_numpy(G["x"])
        got = OT.t5_encoder(enc.state_dict(), x + pos[None, :x.shape[1]], n_layers=NL, n_heads=H)
        assert torch.allclose(got, torch.from_numpy(G["enc_out"]), atol=3e-5, rtol=1e-5)
        sd = {"decoder." + k: v for k, v in dec.state_dict().items()}
        toks, margins = OT.greedy_generate(sd, torch.from_numpy(G["enc_hs"]), embed=emb.weight.detach(),
                                           lm_head=emb.weight.detach(), n_layers=NL, n_heads=H, max_length=L, pos=pos,
                                           stop_at_eos=False, return_margins=True)
    assert np.array_equal(toks.numpy(), G["tokens"])
    assert np.allclose(margins.numpy(), G["margins"], atol=2e-4)


@pytest.mark.gpu
def test_native_fp32_reproduces_hf_golden(cuda_device, native_lib):
    enc, dec, emb, head = (m.to(cuda_device) for m in _modules())
    x = torch.from_numpy(G["x"]).to(cuda_device)
    got = enc(inputs_embeds=x)["last_hidden_state"].cpu()
    ref = torch.from_numpy(G["enc_out"])
    assert float((got - ref).abs().max()) / float(ref.abs().max()) < 2e-5
    toks = ymt3.task_cond_dec_generate(dec, "t5", emb, head, torch.from_numpy(G["enc_hs"]).to(cuda_device), max_length=L,
                                       stop_at_eos=False).cpu().numpy()
    gold, margins = G["tokens"], G["margins"]
    for n in range(gold.shape[0]):
        if not (toks[n] == gold[n]).all():
            t = int(np.argmax(toks[n] != gold[n]))
            assert margins[n, t] < 1e-4, f"row {n} step {t}: token {toks[n, t]} != HF {gold[n, t]} (margin {margins[n, t]:.2e})"
    first = dec._runtime  # noqa: F841  (runtime exists after generation)
