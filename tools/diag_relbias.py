import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import yourmt3_b200 as ymt3
from oracle import t5 as OT
from tests import test_golden_t5 as TG
dev = torch.device("cuda")
enc, dec, emb, head = TG._modules_rel()
x = torch.from_numpy(TG.GR["x"])
sd = {k: v.clone() for k, v in enc.state_dict().items()}
with torch.no_grad():
    ref_b = OT.t5_encoder(sd, x, n_layers=TG.NL, n_heads=TG.H)
    sd0 = {k: v for k, v in sd.items() if "relative_attention_bias" not in k}
    ref_0 = OT.t5_encoder(sd0, x, n_layers=TG.NL, n_heads=TG.H)
encg = enc.to(dev)
got = encg(inputs_embeds=x.to(dev))["last_hidden_state"].float().cpu()
n = float(ref_b.abs().max())
print("native vs oracle(with bias):", float((got - ref_b).abs().max()) / n)
print("native vs oracle(no bias)  :", float((got - ref_0).abs().max()) / n)
print("oracle bias vs no bias     :", float((ref_b - ref_0).abs().max()) / n)
print("golden vs oracle(with bias):", float((torch.from_numpy(TG.GR["enc_out"]) - ref_b).abs().max()) / n)
t = encg._tensors()
print({k: tuple(v.shape) for k, v in t.items() if "relative" in k or "pos" in k})
# one layer only
for nl in (1,):
    cfg = dict(TG.CFG_R, num_layers=nl)
    e1 = ymt3.T5EncoderYMT3(cfg, precision="f32", num_max_positions=TG.NPOS_R)
    e1.load_state_dict({k: v for k, v in sd.items() if not k.startswith("block.1.")})
    with torch.no_grad():
        r1 = OT.t5_encoder(e1.state_dict(), x, n_layers=nl, n_heads=TG.H)
    g1 = e1.to(dev)(inputs_embeds=x.to(dev))["last_hidden_state"].float().cpu()
    d = (g1 - r1).abs()
    print("1 layer native vs oracle:", float(d.max()) / float(r1.abs().max()), "worst position", np.unravel_index(int(d.argmax()), d.shape))
    print("per-position max err (first 20 / last 5 of seq):", [round(float(v), 4) for v in d.amax((0, 2))[:20]], [round(float(v), 4) for v in d.amax((0, 2))[-5:]])
