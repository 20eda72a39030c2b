"""Pin the plain-torch T5 oracle (oracle/t5.py) against the installed reference dependency:
transformers.models.t5.modeling_t5.T5Stack with the same weights (relative attention bias
zeroed = upstream's absolute-position variant). CPU only."""
import copy

import pytest
import torch

from oracle import t5 as OT

transformers = pytest.importorskip("transformers")
from transformers import T5Config  # noqa: E402
from transformers.models.t5.modeling_t5 import T5Stack  # noqa: E402


def make_stacks(n_layers=2, seed=0, std=0.05, relative_bias=False):
    cfg = T5Config(vocab_size=64, d_model=512, d_kv=64, d_ff=1024, num_layers=n_layers, num_decoder_layers=n_layers,
                   num_heads=6, feed_forward_proj="gated-gelu", dropout_rate=0.0, layer_norm_epsilon=1e-6,
                   is_decoder=False, use_cache=False)
    dcfg = copy.deepcopy(cfg)
    dcfg.is_decoder, dcfg.use_cache = True, True
    g = torch.Generator().manual_seed(seed)
    stacks = []
    for c in (cfg, dcfg):
        m = T5Stack(c).eval()
        with torch.no_grad():
            for n, p in m.named_parameters():
                if "relative_attention_bias" in n:
                    if relative_bias:
                        p.copy_(torch.randn(p.shape, generator=g) * 0.5)
                    else:
                        p.zero_()
                elif p.dim() >= 2:
                    p.copy_(torch.randn(p.shape, generator=g) * std)
                else:
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
        stacks.append(m)
    return stacks


def test_encoder_matches_hf():
    enc, _ = make_stacks()
    x = torch.randn(2, 37, 512, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = enc(inputs_embeds=x).last_hidden_state
        got = OT.t5_encoder(enc.state_dict(), x, n_layers=2, n_heads=6)
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5)


def test_decoder_full_and_incremental_match_hf():
    enc, dec = make_stacks(seed=3)
    g = torch.Generator().manual_seed(2)
    enc_hs = torch.randn(2, 19, 512, generator=g)
    x = torch.randn(2, 7, 512, generator=g)
    sd = dec.state_dict()
    with torch.no_grad():
        ref = dec(inputs_embeds=x, encoder_hidden_states=enc_hs, use_cache=False).last_hidden_state
        got = OT.t5_decoder_full(sd, x, enc_hs, n_layers=2, n_heads=6)
        assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5)
        # incremental with KV cache == full teacher forcing (oracle) == HF incremental
        st = OT.T5DecoderState(sd, enc_hs, n_layers=2, n_heads=6)
        past = None
        for t in range(7):
            o = st.step(x[:, t:t + 1])
            h = dec(inputs_embeds=x[:, t:t + 1], encoder_hidden_states=enc_hs, past_key_values=past, use_cache=True)
            past = h.past_key_values
            assert torch.allclose(o, ref[:, t:t + 1], atol=3e-5, rtol=1e-5)
            assert torch.allclose(o, h.last_hidden_state, atol=3e-5, rtol=1e-5)


def test_relative_attention_bias_matches_hf():
    """the HF relative position bias (bucketed, block 0, shared) is switchable in the oracle - encoder
    (bidirectional buckets, lengths beyond max_distance), decoder full and incremental (unidirectional)."""
    enc, dec = make_stacks(seed=7, relative_bias=True)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 150, 512, generator=g)            # > max_distance 128: the logarithmic buckets saturate
    with torch.no_grad():
        ref = enc(inputs_embeds=x).last_hidden_state
        got = OT.t5_encoder(enc.state_dict(), x, n_layers=2, n_heads=6)
        assert torch.allclose(got, ref, atol=3e-5, rtol=1e-5)
        off = OT.t5_encoder({k: v for k, v in enc.state_dict().items() if "relative_attention_bias" not in k}, x,
                            n_layers=2, n_heads=6)
        assert float((off - ref).abs().max()) > 1e-2     # the bias matters: the switch is not vacuous
        enc_hs = torch.randn(2, 19, 512, generator=g)
        y = torch.randn(2, 40, 512, generator=g)
        sd = dec.state_dict()
        ref = dec(inputs_embeds=y, encoder_hidden_states=enc_hs, use_cache=False).last_hidden_state
        assert torch.allclose(OT.t5_decoder_full(sd, y, enc_hs, n_layers=2, n_heads=6), ref, atol=3e-5, rtol=1e-5)
        st = OT.T5DecoderState(sd, enc_hs, n_layers=2, n_heads=6)
        for t in range(40):
            assert torch.allclose(st.step(y[:, t:t + 1]), ref[:, t:t + 1], atol=3e-4, rtol=1e-4)   # fp32 summation order
    # bucket function against HF's static method on a grid of offsets
    from transformers.models.t5.modeling_t5 import T5Attention
    rel = torch.arange(-300, 301)[None, :]
    for bidir in (True, False):
        assert torch.equal(OT.relative_position_bucket(rel, bidir), T5Attention._relative_position_bucket(rel, bidir))


def test_sinusoidal_table_properties():
    p = OT.sinusoidal_positions(1024, 512)
    assert p.shape == (1024, 512)
    assert torch.allclose(p[0, :256], torch.zeros(256)) and torch.allclose(p[0, 256:], torch.ones(256))
    assert torch.allclose(p[:, 0], torch.sin(torch.arange(1024.0)), atol=1e-4)
    assert torch.allclose((p[:, :256] ** 2 + p[:, 256:] ** 2), torch.ones(1024, 256), atol=1e-5)


def test_greedy_generate_semantics():
    """EOS handling: finished rows emit pad, loop stops when all rows are done; tokens non-degenerate."""
    _, dec = make_stacks(seed=5)
    g = torch.Generator().manual_seed(4)
    sd = {"decoder." + k: v for k, v in dec.state_dict().items()}
    V = 96
    embed = torch.randn(V, 512, generator=g) * 0.2   # non-degenerate init
    enc_hs = torch.randn(3, 11, 512, generator=g)
    pos = OT.sinusoidal_positions(64, 512)
    kw = dict(embed=embed, lm_head=embed, n_layers=2, n_heads=6, max_length=24, pos=pos)
    toks, margins = OT.greedy_generate(sd, enc_hs, stop_at_eos=False, return_margins=True, **kw)
    assert toks.shape == (3, 24)
    assert len(torch.unique(toks)) > 8, "degenerate decode (SURVEY H4)"
    # choose an eos id that actually occurs, then check padding after it
    row0 = toks[0].tolist()
    eos = row0[5]
    first = row0.index(eos)
    toks2 = OT.greedy_generate(sd, enc_hs[:1], eos_id=eos, **kw)
    assert toks2[0, : first + 1].tolist() == row0[: first + 1]
    assert (toks2[0, first + 1:] == 0).all()


def test_prefix_conditioning_property():
    """Forcing the first k tokens of a free-running decode as task prefix reproduces its continuation."""
    _, dec = make_stacks(seed=5)
    g = torch.Generator().manual_seed(4)
    sd = {"decoder." + k: v for k, v in dec.state_dict().items()}
    embed = torch.randn(96, 512, generator=g) * 0.2
    enc_hs = torch.randn(3, 11, 512, generator=g)
    kw = dict(embed=embed, lm_head=embed, n_layers=2, n_heads=6, pos=OT.sinusoidal_positions(64, 512), stop_at_eos=False)
    free = OT.greedy_generate(sd, enc_hs, max_length=16, **kw)
    forced = OT.greedy_generate(sd, enc_hs, max_length=13, prefix_ids=free[:, :3], **kw)
    assert torch.equal(forced, free[:, 3:])
