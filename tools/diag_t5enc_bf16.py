"""Diagnostic (GPU box): bf16 T5 encoder hidden-state error vs the fp32 oracle, by depth and input kind."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import yourmt3_b200 as ymt3  # noqa: E402
from oracle import pipeline as OP  # noqa: E402
from tests.util import synth_multitrack  # noqa: E402

dev = torch.device("cuda")
audio = torch.from_numpy(synth_multitrack(2, seed=101))
for layers in (1, 2, 4, 8):
    cfg = ymt3.get_model_cfg("mt3_t5_small")
    cfg["encoder"]["t5"]["num_layers"] = layers
    cfg["decoder"]["t5"]["num_layers"] = 1
    res = {}
    for prec in ("f32", "bf16"):
        m = ymt3.init_nondegenerate_(ymt3.YourMT3(model_cfg=cfg, precision=prec), 0).to(dev)
        feats = m.spectrogram(audio.unsqueeze(1).to(dev))
        for kind, x in (("logmel", feats), ("randn", torch.randn(2, 256, 512, generator=torch.Generator().manual_seed(1)).to(dev))):
            got = m.encoder(inputs_embeds=x)["last_hidden_state"].float().cpu()
            ref = OP.t5_encode(m.state_dict(), x.cpu(), m.model_cfg, m.encoder.pos_table.shape[0])
            e = (got - ref).abs()
            res[(prec, kind)] = (float(e.max()) / float(ref.abs().max()), float(e.median()) / float(ref.abs().max()),
                                 float(((got - ref).norm() / ref.norm())))
    print(layers, "layers:", {k: tuple(round(v, 5) for v in vv) for k, vv in res.items()}, "| feats range",
          float(feats.min()), float(feats.max()), flush=True)
