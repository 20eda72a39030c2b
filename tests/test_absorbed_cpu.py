"""Host logic of the absorbed cross-attention (t5mod_helper.fold_cross_projection): the folded weights must
reproduce the reference cross-attention (HF modeling_t5.py:269-305 on enc_hs = mc_shared_linear(z)) exactly
(fp64 algebra check, CPU only; the kernel itself is checked in tests/test_absorbed_gpu.py)."""
import torch

from yourmt3_b200.t5mod import MultiChannelT5Decoder
from yourmt3_b200.t5mod_helper import fold_cross_projection


def test_fold_matches_reference_cross_attention():
    torch.manual_seed(0)
    cfg = dict(d_model=64, num_heads=3, d_kv=16, num_layers=2, ff_widening_factor=2, num_channels=4)
    dec = MultiChannelT5Decoder(cfg, num_max_positions=8)
    for p in dec.parameters():
        torch.nn.init.normal_(p, 0.0, 0.3)
    Z, D, H, dk, T, N = 24, 64, 3, 16, 11, 5
    proj = torch.nn.Linear(Z, D)
    torch.nn.init.normal_(proj.weight, 0.0, 0.3)
    torch.nn.init.normal_(proj.bias, 0.0, 0.5)
    folded = fold_cross_projection(dec, proj)
    z = torch.randn(N, T, Z, dtype=torch.float64)
    x = torch.randn(N, D, dtype=torch.float64)            # normed decoder hidden state of one step
    for i, blk in enumerate(dec.block):
        att = blk.layer[1].EncDecAttention
        Wq, Wk, Wv, Wo = (w.weight.detach().double() for w in (att.q, att.k, att.v, att.o))
        enc = z @ proj.weight.detach().double().T + proj.bias.detach().double()
        q = (x @ Wq.T).view(N, H, dk)
        k = (enc @ Wk.T).view(N, T, H, dk)
        v = (enc @ Wv.T).view(N, T, H, dk)
        p = torch.softmax(torch.einsum("nhd,nthd->nht", q, k), -1)      # T5: no 1/sqrt(d)
        ref = torch.einsum("nht,nthd->nhd", p, v).reshape(N, H * dk) @ Wo.T
        key = f"block.{i}.layer.1.EncDecAttention."
        qa, oa, ob = (folded[key + s].double() for s in ("q_absorbed.weight", "o_absorbed.weight", "o_absorbed.bias"))
        assert qa.shape == (H * Z, D) and oa.shape == (D, H * Z) and ob.shape == (D,)
        qz = (x @ qa.T).view(N, H, Z)
        pz = torch.softmax(torch.einsum("nhz,ntz->nht", qz, z), -1)
        got = torch.einsum("nht,ntz->nhz", pz, z).reshape(N, H * Z) @ oa.T + ob
        # folded tensors are stored in f32 -> agreement to f32 rounding of the weights
        assert float((got - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
