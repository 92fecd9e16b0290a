"""Evaluation helpers with the reference's call surface.

  calculate_cer(pre_tokens, gt_tokens) -> (cer, S, D, I, N)        rnnt_eval.py:11-56
  calculate_cer_batch(hyps, refs)      -> [(cer, S, D, I, N), ...]  one launch for a whole evaluation batch

The edit-distance table and the backtrace (with the reference's tie-breaking order) run in `ctcvr_cer_batch`; the
per-utterance Python O(m*n) loops of the reference's eval scripts become one kernel per batch."""
from typing import List, Sequence, Tuple

import torch

from ._lib import call, ptr, query, stream


def calculate_cer_batch(hyps: Sequence[Sequence[int]], refs: Sequence[Sequence[int]],
                        device=None) -> List[Tuple[float, int, int, int, int]]:
    if len(hyps) != len(refs):
        raise RuntimeError("calculate_cer_batch: hypotheses and references differ in number")
    n = len(hyps)
    if n == 0:
        return []
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        raise RuntimeError("ctcvr_b200 ops run on CUDA (B200) tensors only; there is no CPU path")
    lh = max(1, max(len(h) for h in hyps))
    lr = max(1, max(len(r) for r in refs))
    hyp = torch.zeros((n, lh), dtype=torch.int32)
    ref = torch.zeros((n, lr), dtype=torch.int32)
    for i, (h, r) in enumerate(zip(hyps, refs)):
        if len(h):
            hyp[i, :len(h)] = torch.as_tensor(list(h), dtype=torch.int32)
        if len(r):
            ref[i, :len(r)] = torch.as_tensor(list(r), dtype=torch.int32)
    hl = torch.tensor([len(h) for h in hyps], dtype=torch.int32)
    rl = torch.tensor([len(r) for r in refs], dtype=torch.int32)
    hyp, ref, hl, rl = hyp.to(dev), ref.to(dev), hl.to(dev), rl.to(dev)
    out = torch.empty((n, 4), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = torch.empty(max(256, query("ctcvr_cer_ws_bytes", n, lh, lr)), dtype=torch.uint8, device=dev)
        call("ctcvr_cer_batch", ptr(hyp), ptr(hl), lh, ptr(ref), ptr(rl), lr, n, ptr(ws), ws.numel(), ptr(out), stream())
    res = []
    for s, d, i, nn in out.cpu().tolist():
        res.append(((s + d + i) / nn if nn != 0 else 0.0, s, d, i, nn))
    return res


def calculate_cer(pre_tokens: list, gt_tokens: list) -> tuple:
    """Drop-in for rnnt_eval.py:11-56 (one pair; use calculate_cer_batch in an evaluation loop)."""
    return calculate_cer_batch([pre_tokens], [gt_tokens])[0]
