"""Data-parallel training step for the transducer hot path (BASELINE.json configs[3]).

Utterances are sharded across ranks (one process per GPU); every rank runs the fused joint + RNN-T loss
forward/backward on its shard with NO data-path collective; the only exchange is one all-reduce of the
parameter gradients per step (NCCL over NVLink on the GPU box; gloo in the CPU tests).  The loss is the
mean over the GLOBAL batch: each rank scales its per-utterance costs by 1/B_global, so the summed
gradients equal those of the single-process run on the concatenated batch (also for uneven shards).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int):
    """Contiguous shard [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradAllReducer:
    """All-reduces (sum) the gradients of `params` once per step.  The payload is a few MB (joint + predictor
    parameters), i.e. latency-bound on NVSwitch: on NCCL the tensors are reduced in place by one grouped collective;
    other backends (gloo in the CPU tests) go through one flat bucket."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self._flat: Optional[torch.Tensor] = None

    def reduce(self, async_op: bool = False):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return None
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        grads = [p.grad for p in self.params]
        if not async_op and grads[0].is_cuda and dist.get_backend(self.group) == "nccl" and hasattr(dist, "_coalescing_manager"):
            # NCCL: one grouped collective over the gradient tensors in place - no flatten / scatter copies
            with dist._coalescing_manager(group=self.group, device=grads[0].device, async_ops=False):
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            return None
        n = sum(g.numel() for g in grads)
        if self._flat is None or self._flat.numel() != n or self._flat.device != grads[0].device:
            self._flat = torch.empty(n, dtype=torch.float32, device=grads[0].device)
        torch._foreach_copy_(list(self._flat.split([g.numel() for g in grads])), [g.reshape(-1).float() for g in grads])
        work = dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        if async_op:
            return work
        self.scatter_back()
        return None

    def scatter_back(self):
        off = 0
        for p in self.params:
            k = p.numel()
            if p.grad is None:
                p.grad = torch.empty_like(p)
            p.grad.copy_(self._flat[off:off + k].view_as(p))
            off += k


def dp_loss_and_backward(loss_fn, global_batch: int, reducer: Optional[GradAllReducer] = None):
    """loss_fn() must return per-utterance costs [B_local].  Returns the global mean loss (all-reduced
    scalar) after backward + gradient all-reduce."""
    costs = loss_fn()
    local = costs.sum() / float(global_batch)
    local.backward()
    if reducer is not None:
        reducer.reduce()
    total = local.detach().clone()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return total
