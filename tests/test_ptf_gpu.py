"""GPU parity of the YPTF stages (residual-conv pre-encoder, Perceiver-TF encoder with MLP / MoE
feed-forward and RoPE, encoder->decoder projections, multi-channel decode) against the CPU oracle
(oracle/perceiver_tf.py; blocks pinned to HF Perceiver / Mixtral in tests/test_oracle_ptf.py)."""
import numpy as np
import pytest
import torch

import yourmt3_b200 as ymt3
from oracle import perceiver_tf as OPTF
from oracle import pipeline as OP
from tests.test_t5_gpu import assert_tokens_identical
from tests.util import synth_noise

pytestmark = pytest.mark.gpu
SPEC = dict(codec="spec", hop_length=300)


def _err(got, ref):
    return float((got.double() - ref.double()).abs().max()) / max(1.0, float(ref.abs().max()))


def small_model(preset, precision="f32", blocks=1, dec_layers=2, event_length=16, seed=0):
    cfg = ymt3.get_model_cfg(preset)
    cfg["encoder"]["perceiver-tf"]["num_blocks"] = blocks
    cfg["decoder"][cfg["decoder_type"]]["num_layers"] = dec_layers
    cfg["event_length"] = event_length
    m = ymt3.YourMT3(audio_cfg=ymt3.get_audio_cfg(**SPEC), model_cfg=cfg, precision=precision)
    return ymt3.init_nondegenerate_(m, seed)


@pytest.mark.parametrize("precision,tol", [("f32", 3e-5), ("bf16", 3e-2)])
def test_res3b_pre_encoder(cuda_device, native_lib, precision, tol):
    m = small_model("yptf", precision).to(cuda_device)
    x = torch.randn(2, 5, 1024, generator=torch.Generator().manual_seed(1)) * 2.0 - 3.0     # log-spectrogram-like
    got = m.pre_encoder(x.to(cuda_device)).float().cpu()
    with torch.no_grad():
        ref = OPTF.pre_encoder_res3b({k: v.cpu().float() for k, v in m.state_dict().items()}, x)
    assert got.shape == ref.shape == (2, 5, 128, 128)
    assert _err(got, ref) < tol


def _encoder_oracle(m, x, cfg):
    """oracle hidden states + the conditioning of every MoE routing decision it took (oracle.perceiver_tf.ROUTER_TRACE)"""
    OPTF.ROUTER_TRACE = []
    try:
        with torch.no_grad():
            ref = OPTF.perceiver_tf_encoder({k: v.cpu().float() for k, v in m.state_dict().items()}, x, cfg)
        gaps = torch.cat([g for _, g in OPTF.ROUTER_TRACE]) if OPTF.ROUTER_TRACE else None
    finally:
        OPTF.ROUTER_TRACE = None
    return ref, gaps


@pytest.mark.parametrize("preset", ["yptf", "yptf_moe_multi"])
def test_perceiver_tf_encoder_f32(cuda_device, native_lib, preset):
    """fp32 path, 2 blocks.  MoE: a routing decision can only differ from the oracle's where the oracle's own gap
    between the last selected and the first rejected expert is at fp32 round-off level; the instance (seed) is chosen
    so that no decision is closer than 5e-6 (asserted), and then EVERY token must match to fp32 accuracy - no
    blanket allowance for flipped tokens."""
    m = small_model(preset, "f32", blocks=2).to(cuda_device)
    cfg = m.model_cfg["encoder"]["perceiver-tf"]
    x = torch.randn(2, 7, 128, 128, generator=torch.Generator().manual_seed(2))
    got = m.encoder(inputs_embeds=x.to(cuda_device))["last_hidden_state"].cpu()
    ref, gaps = _encoder_oracle(m, x, cfg)
    assert got.shape == ref.shape == (2, 7, cfg["num_latents"], 128)
    d = (got - ref).abs() / max(1.0, float(ref.abs().max()))
    if cfg["ff_layer_type"] == "moe":
        print(f"{preset} f32: {gaps.numel()} routing decisions, min relative gap {float(gaps.min()):.2e}; max err {float(d.max()):.2e}")
        assert float(gaps.min()) > 5e-6, "ill-conditioned routing instance: pick another seed"
    assert float(d.max()) < 5e-6        # measured 7.3e-7 (MoE, min routing gap 3.1e-5) on B200


def test_perceiver_tf_encoder_bf16(cuda_device, native_lib):
    """bf16 path, 1 block (5 MoE layers).  Stated tolerances = measured on B200 (round 2, profiles/r02_ptf_tolerances.txt)
    + margin: median element error 0.0019 of the output range (gate 0.004); a token counts as re-routed when its
    error exceeds 0.1 of the range: measured 6.0 % of the tokens, next to 6.4 % of the oracle's routing decisions lying
    within the bf16 perturbation (relative logit gap < 2e-2) of a tie (gate 10 %; round 1 allowed 25 %); max 0.485 of
    the range (gate 0.6)."""
    m = small_model("yptf_moe_multi", "bf16", blocks=1).to(cuda_device)
    cfg = m.model_cfg["encoder"]["perceiver-tf"]
    x = torch.randn(2, 7, 128, 128, generator=torch.Generator().manual_seed(2))
    got = m.encoder(inputs_embeds=x.to(cuda_device))["last_hidden_state"].float().cpu()
    ref, gaps = _encoder_oracle(m, x, cfg)
    d = (got - ref).abs() / max(1.0, float(ref.abs().max()))
    flipped = float((d.amax(-1) > 0.1).float().mean())
    near = float((gaps < 2e-2).float().mean())
    print(f"yptf_moe_multi bf16 encoder: median err {float(d.median()):.4f}, max {float(d.max()):.3f}, tokens beyond 0.1 of "
          f"the range {flipped:.4f}; routing decisions within the bf16 perturbation of a tie {near:.4f}")
    assert float(d.median()) < 4e-3
    assert flipped < 0.10 and flipped < 2.0 * near + 0.02      # re-routed tokens track the near-tie decisions
    assert float(d.max()) < 0.6


@pytest.mark.parametrize("preset", ["yptf", "yptf_moe_multi"])
def test_pre_decoder_projection(cuda_device, native_lib, preset):
    m = small_model(preset, "f32").to(cuda_device)
    cfg = m.model_cfg["encoder"]["perceiver-tf"]
    h = torch.randn(2, 9, cfg["num_latents"], 128, generator=torch.Generator().manual_seed(3))
    got = m.pre_decoder(h.to(cuda_device)).cpu()
    kind = "mc_shared_linear" if preset == "yptf_moe_multi" else "linear"
    ref = OPTF.pre_decoder({k: v.cpu().float() for k, v in m.state_dict().items()}, h, kind, 13)
    assert got.shape == ref.shape
    assert _err(got, ref) < 3e-5


@pytest.mark.parametrize("preset", ["yptf", "yptf_moe_multi"])
def test_full_inference_tokens_f32(cuda_device, native_lib, preset):
    """audio -> tokens, reduced depth, fp32 path: identical tokens to the end-to-end CPU oracle."""
    m = small_model(preset, "f32", blocks=1, dec_layers=2, event_length=12, seed=5).to(cuda_device)
    audio = synth_noise(1, seed=21)
    ref, margins = OP.transcribe(m.state_dict(), audio, m.audio_cfg, m.model_cfg, n_pos=m.decoder.pos_table.shape[0],
                                 max_length=12, stop_at_eos=False, return_margins=True)
    got = m.inference(torch.from_numpy(audio).unsqueeze(1).to(cuda_device), stop_at_eos=False)
    if preset == "yptf_moe_multi":
        assert got.shape == (1, 13, 12)
        got = got.reshape(13, 12)
    else:
        assert got.shape == (1, 12)
    if preset == "yptf_moe_multi":
        assert len(np.unique(ref.numpy())) > 3
    # (single-channel 'yptf' with random weights decodes a constant token: the check is then only as strong as
    #  the hidden-state parity tests above; the multi-channel model gives varied tokens)
    assert_tokens_identical(got.cpu().numpy(), ref.numpy(), margins.numpy(), preset)


def test_full_inference_bf16_runs_and_agrees(cuda_device, native_lib):
    m32 = small_model("yptf_moe_multi", "f32", blocks=1, dec_layers=2, event_length=8, seed=5).to(cuda_device)
    m16 = small_model("yptf_moe_multi", "bf16", blocks=1, dec_layers=2, event_length=8, seed=5).to(cuda_device)
    x = torch.from_numpy(synth_noise(2, seed=22)).unsqueeze(1).to(cuda_device)
    a, b = m32.inference(x, stop_at_eos=False).cpu(), m16.inference(x, stop_at_eos=False).cpu()
    assert a.shape == b.shape == (2, 13, 8)
    assert float((a[..., 0] == b[..., 0]).float().mean()) >= 0.75      # first-step tokens mostly agree in bf16


def test_transcribe_waveform_equals_segmented_inference(cuda_device, native_lib):
    """whole-waveform entry (fused segmentation) == slice_padded_array + inference, identical tokens."""
    from yourmt3_b200.audio_utils import slice_padded_array
    m = small_model("yptf_moe_multi", "f32", blocks=1, dec_layers=2, event_length=6, seed=5).to(cuda_device)
    wave = torch.from_numpy(synth_noise(1, 32767 * 2 + 9000, seed=33)[0]).to(cuda_device)
    a = m.transcribe_waveform(wave, bsz=2, stop_at_eos=False)
    b = m.inference(slice_padded_array(wave, 32767, 32767), stop_at_eos=False)
    assert a.shape == b.shape == (3, 13, 6)
    assert torch.equal(a, b)


def test_perceiver_tf_bf16_tensor_core_attention_matches_simt(cuda_device, native_lib):
    """RoPE + scale folded into the tensor-core tiny attention: the bf16 Perceiver-TF encoder output with the
    tensor-core kernel vs with the fp32-math SIMT kernel (YMT3_NO_TC_ATTN=1) - same weights, same input."""
    import os
    m = small_model("yptf_moe_multi", "bf16", blocks=1).to(cuda_device)
    x = torch.randn(2, 7, 128, 128, generator=torch.Generator().manual_seed(4)).to(cuda_device)
    a = m.encoder(inputs_embeds=x)["last_hidden_state"].float().cpu()
    os.environ["YMT3_NO_TC_ATTN"] = "1"
    try:
        b = m.encoder(inputs_embeds=x)["last_hidden_state"].float().cpu()
    finally:
        os.environ.pop("YMT3_NO_TC_ATTN", None)
    d = (a - b).abs() / max(1.0, float(b.abs().max()))
    flipped = float((d.amax(-1) > 0.1).float().mean())
    print(f"tensor-core vs SIMT attention (bf16 encoder): median diff {float(d.median()):.4f}, tokens beyond 0.1: {flipped:.4f}")
    # two bf16 evaluations of the same encoder re-route independently: measured median 0.0024, 9.3 % of the tokens
    # beyond 0.1 of the range (about 1.5x the single-path rate of test_perceiver_tf_encoder_bf16); gates = + margin
    assert float(d.median()) < 5e-3
    assert flipped < 0.14
