"""RNNPredictor with the reference's constructor and parameter names
(model/component/predictor.py:11-98; wenet/transducer/predictor.py:60-210 is the same math plus the
cache batching helpers).  `self.rnn` stays an `nn.LSTM` so that checkpoints load unchanged (state_dict keys
rnn.weight_ih_l*, ...), but it is only the parameter container: the recurrence of the training forward / backward
 runs in the persistent sequence kernels of csrc/lstm_seq.cu (`functional.lstm_sequence`, one
call per layer), not in the library LSTM.  The decode loops never call forward_step per token - the on-device
decoders in decode.py consume the parameters directly."""
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import functional as CF


def lstm_stack(rnn: nn.LSTM, x: torch.Tensor, h0: torch.Tensor, c0: torch.Tensor, training: bool):
    """`out, (h_n, c_n) = rnn(x, (h0, c0))` (predictor.py:58,90) through the sequence kernels, one
    `functional.lstm_sequence` per layer: x [B,U1,E] batch-first, h0 / c0 [L,B,H].  `rnn` is only read for its
    parameters (and the inter-layer dropout rate), so it may be the reference's own module (patch.install)."""
    if not x.is_cuda:
        raise RuntimeError("ctcvr_b200.RNNPredictor runs on CUDA (B200) tensors only; there is no CPU path")
    if not rnn.batch_first or rnn.bidirectional or getattr(rnn, "proj_size", 0):
        raise RuntimeError("ctcvr_b200: only the batch-first, unidirectional nn.LSTM the reference builds is supported")
    hs, cs = [], []
    for l in range(rnn.num_layers):
        b_ih = getattr(rnn, f"bias_ih_l{l}") if rnn.bias else None
        b_hh = getattr(rnn, f"bias_hh_l{l}") if rnn.bias else None
        x, hn, cn = CF.lstm_sequence(x, getattr(rnn, f"weight_ih_l{l}"), getattr(rnn, f"weight_hh_l{l}"), b_ih, b_hh,
                                     h0[l], c0[l])
        hs.append(hn)
        cs.append(cn)
        if l + 1 < rnn.num_layers and rnn.dropout > 0.0:           # nn.LSTM's inter-layer dropout
            x = F.dropout(x, rnn.dropout, training)
    return x, torch.stack(hs), torch.stack(cs)


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

    def _rnn(self, x: torch.Tensor, h0: torch.Tensor, c0: torch.Tensor):
        return lstm_stack(self.rnn, x, h0, c0, self.training)

    def forward(self, input: torch.Tensor, cache: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        embed = self.dropout(self.embed(input))
        if cache is None:
            st = self.init_state(input.size(0), input.device)
            states = (st[0], st[1])
        else:
            assert len(cache) == 2
            states = (cache[0], cache[1])
        out, _, _ = self._rnn(embed, states[0], states[1])
        return self.projection(out)

    def forward_step(self, input: torch.Tensor, padding: torch.Tensor,
                     cache: List[torch.Tensor]) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        assert len(cache) == 2
        state_m, state_c = cache[0], cache[1]
        embed = self.dropout(self.embed(input.to(self.embed.weight.device)))
        out, m, c = self._rnn(embed, state_m, state_c)
        out = self.projection(out)
        pad = padding.unsqueeze(0)
        m = pad * state_m + m * (1 - pad)
        c = pad * state_c + c * (1 - pad)
        return out, [m, c]

    def batch_to_cache(self, cache: List[torch.Tensor]) -> List[List[torch.Tensor]]:
        return [[m, c] for m, c in zip(torch.split(cache[0], 1, dim=1), torch.split(cache[1], 1, dim=1))]

    def cache_to_batch(self, cache: List[List[torch.Tensor]]) -> List[torch.Tensor]:
        return [torch.cat([s[0] for s in cache], dim=1), torch.cat([s[1] for s in cache], dim=1)]
