"""LM head (upstream amt/src/model/lm_head.py [RECALL]): bias-free vocab projection; when the
embeddings are tied the hidden state is scaled by d_model**-0.5 (HF modeling_t5.py:1105-1110).
In the native path the projection + greedy arg-max run inside the decode step."""
from torch import nn


class LMHead(nn.Module):
    def __init__(self, decoder_config, init_factor: float = 1.0, tie_word_embeddings: bool = True):
        super().__init__()
        self.d_model = decoder_config["d_model"]
        self.init_factor = init_factor
        self.tie_word_embeddings = tie_word_embeddings
        self.lm_head = nn.Linear(decoder_config["d_model"], decoder_config["vocab_size"], bias=False)
