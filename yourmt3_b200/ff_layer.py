"""Feed-forward variants of the Perceiver-TF layers (upstream amt/src/model/ff_layer.py [RECALL]):
dense MLP (HF PerceiverMLP) and Mixtral-style sparse MoE (router -> top-k -> gated experts).
Parameter containers only; the arithmetic runs inside ``ymt3_ptf_forward``."""
from torch import nn


class PerceiverMLP(nn.Module):
    def __init__(self, d, widening):
        super().__init__()
        self.dense1 = nn.Linear(d, widening * d)
        self.dense2 = nn.Linear(widening * d, d)


class MoEExpert(nn.Module):
    """out = w2(act(w1 x) * w3 x)  (Mixtral BlockSparseTop2MLP naming)"""

    def __init__(self, d, hidden):
        super().__init__()
        self.w1 = nn.Linear(d, hidden, bias=False)
        self.w2 = nn.Linear(hidden, d, bias=False)
        self.w3 = nn.Linear(d, hidden, bias=False)


class MoE(nn.Module):
    def __init__(self, d, widening, num_experts, topk):
        super().__init__()
        self.num_experts, self.topk = num_experts, topk
        self.gate = nn.Linear(d, num_experts, bias=False)
        self.experts = nn.ModuleList([MoEExpert(d, widening * d) for _ in range(num_experts)])
