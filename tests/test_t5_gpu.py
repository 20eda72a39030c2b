"""GPU parity of the T5 path (fp32 FFMA kernels) against the CPU oracle pinned to HF T5:
encoder hidden states, decoded TOKENS (bar: identical), EOS semantics, full inference()."""
import numpy as np
import pytest
import torch

import yourmt3_b200 as ymt3
from oracle import pipeline as OP
from tests.util import synth_multitrack, synth_noise

pytestmark = pytest.mark.gpu


def small_cfg(n_layers=2, event_length=48):
    cfg = ymt3.get_model_cfg("mt3_t5_small")
    cfg["encoder"]["t5"]["num_layers"] = n_layers
    cfg["decoder"]["t5"]["num_layers"] = n_layers
    cfg["event_length"] = event_length
    return cfg


@pytest.fixture(scope="module")
def model_small(cuda_device, native_lib):
    m = ymt3.YourMT3(model_cfg=small_cfg())
    ymt3.init_nondegenerate_(m, seed=0)
    return m.to(cuda_device)


def test_encoder_hidden_states(cuda_device, model_small):
    m = model_small
    x = torch.randn(3, 256, 512, generator=torch.Generator().manual_seed(1))
    got = m.encoder(inputs_embeds=x.to(cuda_device))["last_hidden_state"].cpu()
    ref = OP.t5_encode(m.state_dict(), x, m.model_cfg, m.encoder.pos_table.shape[0])
    err = float((got - ref).abs().max()) / float(ref.abs().max())
    assert err < 2e-5, err


def assert_tokens_identical(got, ref, margins, what=""):
    """Bar: identical tokens. A divergence is only tolerated as an fp32 summation-order flip at a
    near-tie of the ORACLE's own logits (top1-top2 margin < 1e-4); anything else fails."""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape
    for n in range(ref.shape[0]):
        if (got[n] == ref[n]).all():
            continue
        t = int(np.argmax(got[n] != ref[n]))
        assert float(margins[n, t]) < 1e-4, f"{what} row {n}: token mismatch at step {t} with margin {float(margins[n, t]):.3e}"


def test_generate_tokens_identical(cuda_device, model_small):
    m = model_small
    enc_hs = torch.randn(5, 40, 512, generator=torch.Generator().manual_seed(2))
    ref, margins = OP.t5_generate(m.state_dict(), enc_hs, m.model_cfg, m.decoder.pos_table.shape[0], 48,
                                  stop_at_eos=False, return_margins=True)
    got = ymt3.task_cond_dec_generate(m.decoder, "t5", m.embed_tokens, m.lm_head, enc_hs.to(cuda_device), max_length=48,
                                      stop_at_eos=False)
    assert got.dtype == torch.long and got.shape == (5, 48)
    assert len(np.unique(ref.numpy())) > 10, "degenerate oracle decode"
    assert_tokens_identical(got.cpu().numpy(), ref.numpy(), margins.numpy(), "generate")
    assert float(margins.median()) > 1e-3


def test_generate_eos_semantics(cuda_device, model_small):
    """Rows that emit EOS are padded afterwards; early stop leaves pad; matches the oracle."""
    m = model_small
    enc_hs = torch.randn(4, 21, 512, generator=torch.Generator().manual_seed(7))
    free = OP.t5_generate(m.state_dict(), enc_hs, m.model_cfg, m.decoder.pos_table.shape[0], 32, stop_at_eos=False)
    eos = int(free[0, 6])                      # a token that really occurs -> EOS fires mid-sequence
    ref, margins = OP.t5_generate(m.state_dict(), enc_hs, m.model_cfg, m.decoder.pos_table.shape[0], 32, stop_at_eos=True,
                                  eos_id=eos, return_margins=True)
    for interval in (0, 4):
        got = ymt3.task_cond_dec_generate(m.decoder, "t5", m.embed_tokens, m.lm_head, enc_hs.to(cuda_device),
                                          max_length=32, stop_at_eos=True, eos_id=eos, early_stop_interval=interval)
        assert_tokens_identical(got.cpu().numpy(), ref.numpy(), margins.numpy(), f"eos interval={interval}")
    row = ref[0].tolist()
    assert eos in row and all(t == 0 for t in row[row.index(eos) + 1:])


def test_task_prefix_tokens(cuda_device, model_small):
    """reference `prefix_ids` / `task_tokens`: teacher-forced prefix, ids returned without the prefix."""
    m = model_small
    enc_hs = torch.randn(4, 21, 512, generator=torch.Generator().manual_seed(9))
    prefix = torch.tensor([[5, 17, 300], [9, 9, 9], [1, 2, 3], [595, 0, 44]])
    ref, margins = OP.t5_generate(m.state_dict(), enc_hs, m.model_cfg, m.decoder.pos_table.shape[0], 20, stop_at_eos=False,
                                  return_margins=True, prefix_ids=prefix)
    got = ymt3.task_cond_dec_generate(m.decoder, "t5", m.embed_tokens, m.lm_head, enc_hs.to(cuda_device), max_length=20,
                                      stop_at_eos=False, prefix_ids=prefix.to(cuda_device))
    assert got.shape == (4, 20)
    assert_tokens_identical(got.cpu().numpy(), ref.numpy(), margins.numpy(), "prefix")
    free = ymt3.task_cond_dec_generate(m.decoder, "t5", m.embed_tokens, m.lm_head, enc_hs.to(cuda_device), max_length=20,
                                       stop_at_eos=False)
    assert not torch.equal(free, got)          # the prefix really conditions the decode
    # property: forcing the free run's first 3 tokens reproduces its continuation
    cont = ymt3.task_cond_dec_generate(m.decoder, "t5", m.embed_tokens, m.lm_head, enc_hs.to(cuda_device), max_length=17,
                                       stop_at_eos=False, prefix_ids=free[:, :3])
    assert torch.equal(cont, free[:, 3:])


def test_decode_lanes_give_identical_tokens(cuda_device, model_small):
    m = model_small
    enc_hs = torch.randn(7, 21, 512, generator=torch.Generator().manual_seed(10)).to(cuda_device)
    a = ymt3.task_cond_dec_generate(m.decoder, "t5", m.embed_tokens, m.lm_head, enc_hs, max_length=16, stop_at_eos=False)
    b = ymt3.task_cond_dec_generate(m.decoder, "t5", m.embed_tokens, m.lm_head, enc_hs, max_length=16, stop_at_eos=False,
                                    lanes=3)
    assert torch.equal(a, b)


def test_full_inference_matches_oracle(cuda_device, model_small):
    """audio -> tokens through YourMT3.inference vs the end-to-end CPU oracle (configs[0] shape, 2 layers)."""
    m = model_small
    audio = np.concatenate([synth_noise(1, seed=11), synth_multitrack(1, seed=12)], 0)
    ref, margins = OP.transcribe_t5(m.state_dict(), audio, m.audio_cfg, m.model_cfg, m.encoder.pos_table.shape[0], 48,
                                    stop_at_eos=False, return_margins=True)
    got = m.inference(torch.from_numpy(audio).unsqueeze(1).to(cuda_device), stop_at_eos=False)
    assert got.shape == (2, 48)
    assert_tokens_identical(got.cpu().numpy(), ref.numpy(), margins.numpy(), "inference")
    outs = m.inference_file(1, torch.from_numpy(audio).unsqueeze(1), stop_at_eos=False)
    assert len(outs) == 2 and (np.concatenate(outs, 0) == got.cpu().numpy()).all()


@pytest.mark.parametrize("seed", [1, 2])
def test_t5_small_full_depth_tokens(cuda_device, native_lib, seed):
    """BASELINE configs[0]: full T5-small shape (8+8 layers), one 2.048 s segment, greedy decode."""
    cfg = ymt3.get_model_cfg("mt3_t5_small")
    cfg["event_length"] = 96
    m = ymt3.init_nondegenerate_(ymt3.YourMT3(model_cfg=cfg), seed=seed).to(cuda_device)
    audio = synth_multitrack(1, seed=100 + seed)
    ref, margins = OP.transcribe_t5(m.state_dict(), audio, m.audio_cfg, m.model_cfg, m.encoder.pos_table.shape[0], 96,
                                    stop_at_eos=False, return_margins=True)
    got = m.inference(torch.from_numpy(audio).unsqueeze(1).to(cuda_device), stop_at_eos=False)
    assert len(np.unique(ref.numpy())) > 5
    assert_tokens_identical(got.cpu().numpy(), ref.numpy(), margins.numpy(), "t5-small")


# ----------------------------------------------------------------------------------------------
# bf16 path (tcgen05 GEMMs, bf16 activations/KV cache, fp32 softmax/norm/accumulate/logits).
# Stated tolerance: first-step logits within 3% of the logit range of the fp32 oracle, and
# >= 90% of the greedily decoded tokens equal to the fp32 oracle's over the first 24 steps
# (after a flip the sequences legitimately diverge, so agreement is measured up to the first
# mismatch per row and averaged).
# ----------------------------------------------------------------------------------------------
def test_bf16_logits_and_tokens(cuda_device, native_lib):
    m = ymt3.YourMT3(model_cfg=small_cfg(), precision="bf16")
    ymt3.init_nondegenerate_(m, seed=0)
    m = m.to(cuda_device)
    enc_hs = torch.randn(6, 40, 512, generator=torch.Generator().manual_seed(2))
    sd = m.state_dict()
    ref, margins = OP.t5_generate(sd, enc_hs, m.model_cfg, m.decoder.pos_table.shape[0], 24, stop_at_eos=False,
                                  return_margins=True)
    # step-0 logits of the oracle
    from oracle import t5 as OT
    dsd = {k[len("decoder."):]: v.cpu().float() for k, v in sd.items() if k.startswith("decoder.")}
    st = OT.T5DecoderState(dsd, enc_hs, n_layers=2, n_heads=6, pos=OT.sinusoidal_positions(8, 512))
    E = sd["embed_tokens.weight"].cpu().float()
    hs = st.step(E[torch.zeros(6, dtype=torch.long)][:, None, :])[:, 0] * (512 ** -0.5)
    ref_logits = hs @ E.T
    got1 = ymt3.task_cond_dec_generate(m.decoder, "t5", m.embed_tokens, m.lm_head, enc_hs.to(cuda_device), max_length=1,
                                       stop_at_eos=False, precision=1)
    logits = m.decoder._runtime.last_logits(6, cuda_device).cpu()
    rng = float(ref_logits.max() - ref_logits.min())
    err = float((logits - ref_logits).abs().max())
    assert err < 0.03 * rng, f"bf16 logit error {err:.4f} vs range {rng:.3f}"
    got = ymt3.task_cond_dec_generate(m.decoder, "t5", m.embed_tokens, m.lm_head, enc_hs.to(cuda_device), max_length=24,
                                      stop_at_eos=False, precision=1).cpu().numpy()
    assert (got[:, 0] == got1.cpu().numpy()[:, 0]).all()
    agree = []
    for n in range(6):
        neq = np.nonzero(got[n] != ref[n].numpy())[0]
        agree.append((neq[0] if len(neq) else 24) / 24.0)
    assert np.mean(agree) >= 0.9, agree


def test_bf16_note_onset_f1_vs_fp32(cuda_device, native_lib):
    """North-star bf16 criterion restated for random weights: decode the SAME synthetic multitrack audio with the
    fp32 (exact) and the bf16 (tcgen05) paths, detokenise both to note events and score bf16 against fp32 with the
    mir_eval-style onset F1 (50 ms).

    With RANDOM weights the greedy margins (median ~0.1) are of the order of the bf16 logit error (<= 3 % of a
    ~10-wide logit range), so sequences diverge after the first flipped token and a high free-running F1 cannot
    be demanded (measured: 0.29 over 16 segments x 32 tokens).  What is asserted is what random weights can
    support: the first decoded token (no error accumulation) agrees for >= 75 % of the rows, the detokeniser
    yields notes for both paths, and the F1 is reported.  A trained checkpoint is required for the north-star's
    "equal onset F1" criterion."""
    from yourmt3_b200 import event_codec as EC
    cfg = small_cfg(n_layers=2, event_length=32)
    audio = torch.from_numpy(synth_multitrack(16, seed=77)).unsqueeze(1).to(cuda_device)
    toks = {}
    for prec in ("f32", "bf16"):
        m = ymt3.init_nondegenerate_(ymt3.YourMT3(model_cfg=cfg, precision=prec), seed=3).to(cuda_device)
        toks[prec] = m.inference(audio, stop_at_eos=False).cpu().numpy()
    ref, est = EC.batch_tokens_to_notes(toks["f32"]), EC.batch_tokens_to_notes(toks["bf16"])
    # random-init decodes have no tie token: treat every pitch token as an onset by prepending one
    if not ref:
        tie = EC.encode_event("tie", 0)
        ref = EC.batch_tokens_to_notes(np.concatenate([np.full((16, 1), tie), toks["f32"]], 1))
        est = EC.batch_tokens_to_notes(np.concatenate([np.full((16, 1), tie), toks["bf16"]], 1))
    p, r, f = EC.onset_f1(ref, est)
    agree = float((toks["f32"] == toks["bf16"]).mean())
    print(f"bf16 vs fp32: token agreement {agree:.3f}, notes {len(ref)}/{len(est)}, onset P/R/F1 = {p:.3f}/{r:.3f}/{f:.3f}")
    assert len(ref) > 0 and len(est) > 0
    assert 0.0 <= f <= 1.0
    assert float((toks["f32"][:, 0] == toks["bf16"][:, 0]).mean()) >= 0.75


@pytest.mark.parametrize("n_rows", [7, 300])
def test_gemm_chain_is_bit_identical(cuda_device, native_lib, monkeypatch, n_rows):
    """The chained decode step (gemm_chain_kernel: [o-proj -> cross-q] and [cross-o -> wi -> wo -> next qkv] as ONE
    persistent launch each, dependencies per 128-row tile) computes tile by tile what the separate launches compute:
    tokens AND last-step logits are bit-identical with the chain on and off (YMT3_GEMM_CHAIN=1, read when the
    decoder runtime is created); 300 rows = 3 row tiles, the last one ragged."""
    enc_hs = torch.randn(n_rows, 40, 512, generator=torch.Generator().manual_seed(5)).to(cuda_device)
    out = {}
    for mode in ("chain", "separate"):
        if mode == "chain":
            monkeypatch.setenv("YMT3_GEMM_CHAIN", "1")
        else:
            monkeypatch.delenv("YMT3_GEMM_CHAIN", raising=False)
        m = ymt3.YourMT3(model_cfg=small_cfg(n_layers=3), precision="bf16")
        ymt3.init_nondegenerate_(m, seed=0)
        m = m.to(cuda_device)
        toks = ymt3.task_cond_dec_generate(m.decoder, "t5", m.embed_tokens, m.lm_head, enc_hs, max_length=40,
                                           stop_at_eos=False, precision=1)
        out[mode] = (toks.cpu(), m.decoder._runtime.last_logits(n_rows, cuda_device).cpu())
        # a second call on the same runtime re-zeroes the dependency counters
        toks2 = ymt3.task_cond_dec_generate(m.decoder, "t5", m.embed_tokens, m.lm_head, enc_hs, max_length=40,
                                            stop_at_eos=False, precision=1)
        assert torch.equal(toks2.cpu(), out[mode][0])
    assert len(np.unique(out["chain"][0].numpy())) > 10
    assert torch.equal(out["chain"][0], out["separate"][0])
    assert torch.equal(out["chain"][1], out["separate"][1])
