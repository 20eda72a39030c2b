"""Pin the building blocks of the YPTF oracle (oracle/perceiver_tf.py) against the installed
dependency code they restate: HF PerceiverLayer, HF Mixtral sparse MoE + rotary embedding,
torch Conv2d/BatchNorm2d/AvgPool2d.  The block WIRING (Perceiver-TF, res3b) has no reference
available and stays 'parity unpinned' (DESIGN.md 2). CPU only."""
import pytest
import torch
from torch import nn

from oracle import perceiver_tf as OP

transformers = pytest.importorskip("transformers")


def _rand_(m, g, std=0.1):
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * std + (1.0 if p.dim() == 1 and "norm" in "" else 0.0))
    return m


@pytest.mark.parametrize("cross", [False, True])
def test_perceiver_layer_matches_hf(cross):
    from transformers import PerceiverConfig
    from transformers.models.perceiver.modeling_perceiver import PerceiverLayer
    g = torch.Generator().manual_seed(0)
    D, C, H = 128, 96, (1 if cross else 8)
    cfg = PerceiverConfig(d_latents=D, hidden_act="gelu", attention_probs_dropout_prob=0.0,
                          cross_attention_shape_for_attention="q")
    layer = PerceiverLayer(cfg, is_cross_attention=cross, qk_channels=D, v_channels=D, num_heads=H, q_dim=D,
                           kv_dim=C if cross else D, widening_factor=1, use_query_residual=not cross).eval()
    _rand_(layer, g)
    sd = {"L." + k: v.detach() for k, v in layer.state_dict().items()}
    h = torch.randn(3, 26, D, generator=g)
    inputs = torch.randn(3, 40, C, generator=g) if cross else None
    with torch.no_grad():
        ref = layer(h, inputs=inputs)[0]
        got = OP.perceiver_layer(sd, "L.", h, {"ff_layer_type": "mlp", "hidden_act": "gelu", "layer_norm_eps": 1e-5},
                                 num_heads=H, inputs=inputs, query_residual=not cross)
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5)


def test_moe_matches_hf_mixtral():
    from transformers import MixtralConfig
    from transformers.models.mixtral.modeling_mixtral import MixtralSparseMoeBlock
    g = torch.Generator().manual_seed(1)
    D, I, E, k = 128, 512, 8, 2
    cfg = MixtralConfig(hidden_size=D, intermediate_size=I, num_local_experts=E, num_experts_per_tok=k, hidden_act="silu")
    blk = MixtralSparseMoeBlock(cfg).eval()
    with torch.no_grad():
        blk.gate.weight.copy_(torch.randn(E, D, generator=g) * 0.3)
        blk.experts.gate_up_proj.copy_(torch.randn(E, 2 * I, D, generator=g) * 0.05)
        blk.experts.down_proj.copy_(torch.randn(E, D, I, generator=g) * 0.05)
    sd = {"m.gate.weight": blk.gate.weight.detach()}
    for e in range(E):
        sd[f"m.experts.{e}.w1.weight"] = blk.experts.gate_up_proj[e, :I].detach()     # gate (activated) half
        sd[f"m.experts.{e}.w3.weight"] = blk.experts.gate_up_proj[e, I:].detach()     # up half
        sd[f"m.experts.{e}.w2.weight"] = blk.experts.down_proj[e].detach()
    x = torch.randn(1, 200, D, generator=g)
    with torch.no_grad():
        ref = blk(x)[0] if isinstance(blk(x), tuple) else blk(x)
        got = OP.moe_ff(sd, "m.", x[0], num_experts=E, topk=k, act="silu")
    assert torch.allclose(got, ref.reshape(200, D), atol=2e-5, rtol=1e-5)


def test_rope_matches_hf():
    from transformers.models.mixtral.modeling_mixtral import apply_rotary_pos_emb
    g = torch.Generator().manual_seed(2)
    q, k = torch.randn(2, 8, 26, 16, generator=g), torch.randn(2, 8, 26, 16, generator=g)
    cos, sin = OP.rope_cos_sin(26, 16)
    rq, rk = apply_rotary_pos_emb(q, k, cos[None], sin[None])
    assert torch.allclose(OP.apply_rope(q, 16), rq, atol=1e-6)
    assert torch.allclose(OP.apply_rope(k, 16), rk, atol=1e-6)
    # partial: only the first 8 dims rotate
    part = OP.apply_rope(q, 8)
    assert torch.equal(part[..., 8:], q[..., 8:]) and not torch.allclose(part[..., :8], q[..., :8])


def test_res_block_matches_torch_modules():
    from yourmt3_b200.conv_block import Res2DAVPBlock
    from yourmt3_b200.init_utils import init_nondegenerate_
    blk = init_nondegenerate_(Res2DAVPBlock(4, 8), seed=3).eval()
    x = torch.randn(2, 4, 5, 16, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        h = blk.conv1(torch.relu(blk.bn1(x)))
        h = blk.conv2(torch.relu(blk.bn2(h)))
        ref = nn.functional.avg_pool2d(h + blk.shortcut(x), (1, 2))
        got = OP.res_block({"b." + k: v for k, v in blk.state_dict().items()}, "b.", x)
    assert torch.allclose(got, ref, atol=1e-6)


def test_encode_shapes_both_presets():
    import yourmt3_b200 as ymt3
    for preset, shape in (("yptf", (1, 6, 512)), ("yptf_moe_multi", (1, 13, 6, 512))):
        cfg = ymt3.get_model_cfg(preset)
        cfg["encoder"]["perceiver-tf"]["num_blocks"] = 1
        m = ymt3.YourMT3(audio_cfg=ymt3.get_audio_cfg(codec="spec", hop_length=300), model_cfg=cfg)
        ymt3.init_nondegenerate_(m, 0)
        feats = torch.randn(1, 6, 1024, generator=torch.Generator().manual_seed(0))
        with torch.no_grad():
            enc = OP.encode(m.state_dict(), feats, cfg)
        assert tuple(enc.shape) == shape and torch.isfinite(enc).all()
