"""RNNPredictor with the reference's constructor and parameter names
(model/component/predictor.py:11-98; wenet/transducer/predictor.py:60-210 is the same math plus the
cache batching helpers).  The training forward is the library LSTM (cuDNN) exactly as in the
reference (SURVEY.md §8f ranks it "next"); the decode loops never call forward_step per token any
more — the on-device decoders in decode.py consume the parameters directly."""
from typing import List, Optional, Tuple

import torch
from torch import nn


class RNNPredictor(nn.Module):
    def __init__(self, voca_size: int, embed_size: int, output_size: int, embed_dropout: float, hidden_size: int,
                 num_layers: int, bias: bool = True, rnn_type: str = "lstm", dropout: float = 0.1) -> None:
        super().__init__()
        if rnn_type != "lstm":
            raise RuntimeError("ctcvr_b200.RNNPredictor: only rnn_type='lstm' (the one the reference builds)")
        self.n_layers = num_layers
        self.hidden_size = hidden_size
        self._output_size = output_size
        self.embed = nn.Embedding(voca_size, embed_size)
        self.dropout = nn.Dropout(embed_dropout)
        self.rnn = nn.LSTM(input_size=embed_size, hidden_size=hidden_size, num_layers=num_layers, bias=bias,
                           batch_first=True, dropout=dropout if num_layers > 1 else 0.0)
        self.projection = nn.Linear(hidden_size, output_size)

    def output_size(self):
        return self._output_size

    def init_state(self, batch_size: int, device: torch.device, method: str = "zero") -> List[torch.Tensor]:
        assert batch_size > 0
        return [torch.zeros(self.n_layers, batch_size, self.hidden_size, device=device),
                torch.zeros(self.n_layers, batch_size, self.hidden_size, device=device)]

    def forward(self, input: torch.Tensor, cache: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        embed = self.dropout(self.embed(input))
        if cache is None:
            st = self.init_state(input.size(0), input.device)
            states = (st[0], st[1])
        else:
            assert len(cache) == 2
            states = (cache[0], cache[1])
        out, _ = self.rnn(embed, states)
        return self.projection(out)

    def forward_step(self, input: torch.Tensor, padding: torch.Tensor,
                     cache: List[torch.Tensor]) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        assert len(cache) == 2
        state_m, state_c = cache[0], cache[1]
        embed = self.dropout(self.embed(input.to(self.embed.weight.device)))
        out, (m, c) = self.rnn(embed, (state_m, state_c))
        out = self.projection(out)
        pad = padding.unsqueeze(0)
        m = pad * state_m + m * (1 - pad)
        c = pad * state_c + c * (1 - pad)
        return out, [m, c]

    def batch_to_cache(self, cache: List[torch.Tensor]) -> List[List[torch.Tensor]]:
        return [[m, c] for m, c in zip(torch.split(cache[0], 1, dim=1), torch.split(cache[1], 1, dim=1))]

    def cache_to_batch(self, cache: List[List[torch.Tensor]]) -> List[torch.Tensor]:
        return [torch.cat([s[0] for s in cache], dim=1), torch.cat([s[1] for s in cache], dim=1)]
