"""CTC decoders with the reference's call surface (wenet/transformer/search.py:30-59,107-247).
The per-frame Python loops (one .item() per token) are replaced by one kernel launch per batch."""
from typing import List

import torch

from . import functional as CF
from . import _lib
from ._lib import call, ptr, query, stream


class DecodeResult:
    """wenet/transformer/search.py:30-59."""

    def __init__(self, tokens: List[int], score: float = 0.0, confidence: float = 0.0,
                 tokens_confidence: List[float] = None, times: List[int] = None, nbest: List[List[int]] = None,
                 nbest_scores: List[float] = None, nbest_times: List[List[int]] = None):
        self.tokens = tokens
        self.score = score
        self.confidence = confidence
        self.tokens_confidence = tokens_confidence
        self.times = times
        self.nbest = nbest
        self.nbest_scores = nbest_scores
        self.nbest_times = nbest_times


def ctc_greedy_search(ctc_probs: torch.Tensor, ctc_lens: torch.Tensor, blank_id: int = 0) -> List[DecodeResult]:
    """wenet/transformer/search.py:107-122."""
    return [DecodeResult(h) for h in CF.ctc_greedy_search(ctc_probs, ctc_lens, blank_id)]


@torch.no_grad()
def ctc_prefix_beam_search(ctc_probs: torch.Tensor, ctc_lens: torch.Tensor, beam_size: int, context_graph=None,
                           blank_id: int = 0) -> List[DecodeResult]:
    """wenet/transformer/search.py:125-247 (context_graph must be None: no caller in scope passes one)."""
    if context_graph is not None:
        raise RuntimeError("ctc_prefix_beam_search: context_graph is not supported on the device path")
    _lib.require_cuda(ctc_probs)
    x = ctc_probs.detach().float().contiguous()
    B, T, V = x.shape
    dev = x.device
    lens = ctc_lens.to(device=dev, dtype=torch.int32).contiguous()
    beam = int(beam_size)
    out_n = torch.zeros((B,), dtype=torch.int32, device=dev)
    out_tok = torch.zeros((B, beam, T), dtype=torch.int32, device=dev)
    out_len = torch.zeros((B, beam), dtype=torch.int32, device=dev)
    out_sc = torch.zeros((B, beam), dtype=torch.float64, device=dev)
    out_tm = torch.zeros((B, beam, T), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nbytes = query("ctcvr_ctc_prefix_beam_ws_bytes", B, T, V, beam)
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        call("ctcvr_ctc_prefix_beam", ptr(x), ptr(lens), B, T, V, beam, int(blank_id), ptr(out_n), ptr(out_tok),
             ptr(out_len), ptr(out_sc), ptr(out_tm), ptr(ws), ws.numel(), stream())
    n_h, len_h, sc_h = out_n.cpu().tolist(), out_len.cpu().numpy(), out_sc.cpu().numpy()
    width = max(int(len_h.max()), 1)                        # only the used part of the token / time arrays crosses PCIe
    tok_h = out_tok[:, :, :width].contiguous().cpu().numpy()
    tm_h = out_tm[:, :, :width].contiguous().cpu().numpy()
    results = []
    for b in range(B):
        n = n_h[b]
        nbest = [tuple(tok_h[b, i, :len_h[b, i]].tolist()) for i in range(n)]
        times = [tm_h[b, i, :len_h[b, i]].tolist() for i in range(n)]
        scores = [float(sc_h[b, i]) for i in range(n)]
        results.append(DecodeResult(tokens=nbest[0], score=scores[0], times=times[0], nbest=nbest,
                                    nbest_scores=scores, nbest_times=times))
    return results
