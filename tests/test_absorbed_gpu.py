"""Absorbed cross-attention on the GPU (cross_absorbed.cu, ymt3_t5dec_generate_latent): kernel vs a plain torch fp32
reference, and the whole bf16 multi-channel decode in absorbed form vs the reference-order bf16 path and the fp32
path (same function of the same weights; stated bf16 tolerances)."""
import pytest
import torch

import yourmt3_b200 as ymt3
from yourmt3_b200 import _lib
from tests.test_ptf_gpu import small_model
from tests.util import synth_noise

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,H,T", [(3, 6, 110), (5, 8, 16), (2, 1, 33), (700, 6, 110), (150, 6, 128), (9, 4, 97)])
def test_kernel_vs_torch(cuda_device, native_lib, N, H, T):
    g = torch.Generator().manual_seed(N * 1000 + T)
    Tp = (T + 15) // 16 * 16
    q = (torch.randn(N, H, 256, generator=g) * 0.3).to(cuda_device, torch.bfloat16)
    z = torch.zeros(N, Tp, 256)
    z[:, :T] = torch.randn(N, T, 256, generator=g)
    z = z.to(cuda_device, torch.bfloat16)
    out = torch.full((N, H, 256), float("nan"), dtype=torch.bfloat16, device=cuda_device)
    _lib.check(native_lib.ymt3_op_cross_attn_absorbed(q.data_ptr(), z.data_ptr(), out.data_ptr(), N, H, T, Tp,
                                                      _lib.current_stream_ptr()), "cross_attn_absorbed")
    qf, zf = q.float(), z.float()[:, :T]
    p = torch.softmax(torch.einsum("nhz,ntz->nht", qf, zf), -1)
    ref = torch.einsum("nht,ntz->nhz", p, zf)
    err = float((out.float() - ref).abs().max())
    # bf16 probabilities (2^-9 relative) and bf16 output rounding on |values| <~ 4
    assert err < 3e-2, err
    assert float((out.float() - ref).abs().mean()) < 2e-3


def test_kernel_rejects_bad_shapes(cuda_device, native_lib):
    x = torch.zeros(16 * 256, dtype=torch.bfloat16, device=cuda_device)
    assert native_lib.ymt3_op_cross_attn_absorbed(x.data_ptr(), x.data_ptr(), x.data_ptr(), 1, 9, 16, 16, None) != 0
    assert native_lib.ymt3_op_cross_attn_absorbed(x.data_ptr(), x.data_ptr(), x.data_ptr(), 1, 6, 16, 15, None) != 0
    assert native_lib.ymt3_op_cross_attn_absorbed(x.data_ptr(), x.data_ptr(), x.data_ptr(), 1, 6, 130, 144, None) != 0


def test_absorbed_decode_matches_reference_order(cuda_device, native_lib):
    """YPTF.MoE+Multi bf16: absorbed vs reference-order cross-attention give the same first-step logits to bf16
    accuracy, and both agree with the fp32 path on the first greedy token of (nearly) every channel."""
    L = 8
    m16 = small_model("yptf_moe_multi", "bf16", blocks=1, dec_layers=2, event_length=L, seed=5).to(cuda_device)
    m32 = small_model("yptf_moe_multi", "f32", blocks=1, dec_layers=2, event_length=L, seed=5).to(cuda_device)
    x = torch.from_numpy(synth_noise(3, seed=22)).unsqueeze(1).to(cuda_device)
    assert m16._absorbed()
    tok_abs = m16.inference(x, stop_at_eos=False, max_token_length=1).cpu()
    log_abs = m16.decoder._runtime.last_logits(3 * 13, cuda_device).cpu()
    m16.absorb_cross_attention = False
    tok_ref = m16.inference(x, stop_at_eos=False, max_token_length=1).cpu()
    log_ref = m16.decoder._runtime.last_logits(3 * 13, cuda_device).cpu()
    tok_32 = m32.inference(x, stop_at_eos=False, max_token_length=1).cpu()
    log_32 = m32.decoder._runtime.last_logits(3 * 13, cuda_device).cpu()
    rng = float(log_32.max() - log_32.min())
    e_abs, e_ref, e_ar = (float((a - b).abs().max()) / rng for a, b in ((log_abs, log_32), (log_ref, log_32), (log_abs, log_ref)))
    m_abs = float((log_abs - log_32).abs().median()) / rng
    msg = f"max err / fp32 logit range: absorbed {e_abs:.4f}, reference-order bf16 {e_ref:.4f}, between them {e_ar:.4f}"
    print(msg)
    # stated tolerance: the two bf16 evaluations of the same function differ by < 2 % of the fp32 logit range; against
    # fp32 (where MoE routing flips in the shared bf16 encoder dominate the max) the absorbed form is no worse than
    # 1.5x the reference-order bf16 path, and its median error is < 0.5 % of the range
    assert e_ar < 0.02, msg
    assert e_abs < max(0.03, 1.5 * e_ref), msg
    assert m_abs < 0.005, msg
    assert float((tok_abs == tok_32).float().mean()) >= 0.75
    assert float((tok_abs == tok_ref).float().mean()) >= 0.75
    # full-length run in absorbed form: shape, determinism, token variety
    m16.absorb_cross_attention = True
    a = m16.inference(x, stop_at_eos=False)
    b = m16.inference(x, stop_at_eos=False)
    assert a.shape == (3, 13, L) and torch.equal(a, b)
    assert len(torch.unique(a)) > 3


def test_absorbed_with_task_prefix(cuda_device, native_lib):
    m16 = small_model("yptf_moe_multi", "bf16", blocks=1, dec_layers=2, event_length=6, seed=5).to(cuda_device)
    x = torch.from_numpy(synth_noise(2, seed=23)).unsqueeze(1).to(cuda_device)
    pfx = torch.randint(3, 500, (2, 13, 2), generator=torch.Generator().manual_seed(1))
    a = m16.inference(x, task_tokens=pfx.to(cuda_device), stop_at_eos=False).cpu()
    m16.absorb_cross_attention = False
    b = m16.inference(x, task_tokens=pfx.to(cuda_device), stop_at_eos=False).cpu()
    assert a.shape == b.shape == (2, 13, 6)
    assert float((a[..., 0] == b[..., 0]).float().mean()) >= 0.75


def test_gemm_chain_bit_identical_absorbed(cuda_device, native_lib, monkeypatch):
    """Absorbed (latent) cross-attention decode with the chained GEMM launches vs the separate launches: same tokens
    and last-step logits, bit for bit (13 channels x 12 segments = 156 rows, two row tiles, the second ragged)."""
    x = torch.from_numpy(synth_noise(12, seed=31)).unsqueeze(1).to(cuda_device)
    out = {}
    for mode in ("chain", "separate"):
        if mode == "chain":
            monkeypatch.setenv("YMT3_GEMM_CHAIN", "1")
        else:
            monkeypatch.delenv("YMT3_GEMM_CHAIN", raising=False)
        m = small_model("yptf_moe_multi", "bf16", blocks=1, dec_layers=3, event_length=24, seed=5).to(cuda_device)
        assert m._absorbed()
        toks = m.inference(x, stop_at_eos=False).cpu()
        out[mode] = (toks, m.decoder._runtime.last_logits(12 * 13, cuda_device).cpu())
    assert len(torch.unique(out["chain"][0])) > 3
    assert torch.equal(out["chain"][0], out["separate"][0])
    assert torch.equal(out["chain"][1], out["separate"][1])
