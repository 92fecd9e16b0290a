"""Data-parallel training step for the transducer hot path (BASELINE.json configs[3]).

Utterances are sharded across ranks (one process per GPU); every rank runs the fused joint + RNN-T loss
forward/backward on its shard with NO data-path collective; the only exchange is one all-reduce of the
parameter gradients per step (NCCL over NVLink on the GPU box; gloo in the CPU tests).  The loss is the
mean over the GLOBAL batch: each rank scales its per-utterance costs by 1/B_global, so the summed
gradients equal those of the single-process run on the concatenated batch (also for uneven shards).
"""
from __future__ import annotations

import ctypes
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


class PeerGradExchange:
    """Sum of fp32 gradient tensors over the ranks of ONE node by a single kernel over NVLink peer memory
    (`csrc/peer_reduce.cu`, `ctcvr_peer_allreduce`): pack -> flag barrier -> each rank reduces its slice from all peers
    and stores it to all peers -> flag barrier -> unpack, in place on the tensors passed to `reduce()`.  Unlike an NCCL
    call it is an ordinary kernel launch, so `GraphedJointRnntStep(grad_exchange=...)` captures it inside the step graph
    (a replayed graph holding the NCCL all-reduce hung, see graph.py).  Sums are taken in rank order on the owning rank,
    so every rank holds bit-identical results.

    `torch.distributed` is used once, to pass the 64-byte CUDA IPC handles around.  Every rank must call `reduce()` with
    tensors of the same sizes in the same order.  No fallback: a node without peer access raises."""

    MAX_TENSORS = 24

    def __init__(self, max_floats: int, group=None, ctas: int = 128, _ctx=None, _rank=0, _world=1):
        from . import _lib
        self._L = _lib
        self.ctas = int(ctas)
        self._cache = {}
        if _ctx is not None:                                       # ranks simulated inside one process (tests)
            self.ctx, self.rank, self.world = _ctx, _rank, _world
            return
        if not dist.is_initialized():
            raise RuntimeError("PeerGradExchange needs an initialised torch.distributed process group (one process per GPU)")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if not torch.cuda.is_available():
            raise RuntimeError("PeerGradExchange runs on CUDA (B200) devices only; there is no CPU path")
        ctx = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        _lib.call("ctcvr_peer_create", self.rank, self.world, int(max_floats) + 4 * self.MAX_TENSORS, ctypes.byref(ctx), handle)
        self.ctx = ctx
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        blob = ctypes.create_string_buffer(b"".join(handles), 64 * self.world)
        err = None
        try:
            _lib.call("ctcvr_peer_connect", self.ctx, blob, None)
        except RuntimeError as e:
            err = str(e)
        # nobody launches before every rank has mapped its peers, and all ranks fail together if one could not
        errs: List[Optional[str]] = [None] * self.world
        dist.all_gather_object(errs, err, group=group)
        bad = [f"rank {r}: {e}" for r, e in enumerate(errs) if e]
        if bad:
            self.close()
            raise RuntimeError("PeerGradExchange: " + "; ".join(bad))

    @classmethod
    def local_group(cls, world: int, max_floats: int, ctas: int = 128) -> List["PeerGradExchange"]:
        """`world` ranks inside this process, all on the current device (buffers addressed directly, no IPC): the
        single-GPU test of the kernel - the ranks' kernels must then be launched on DIFFERENT streams."""
        from . import _lib
        ctxs = []
        for r in range(world):
            ctx = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            _lib.call("ctcvr_peer_create", r, world, int(max_floats) + 4 * cls.MAX_TENSORS, ctypes.byref(ctx), handle)
            ctxs.append(ctx)
        bufs = (ctypes.c_void_p * world)(*[_lib.lib().ctcvr_peer_local_buffer(c) for c in ctxs])
        for c in ctxs:
            _lib.call("ctcvr_peer_connect", c, None, bufs)
        return [cls(max_floats, ctas=ctas, _ctx=c, _rank=r, _world=world) for r, c in enumerate(ctxs)]

    def set_timeout_ms(self, ms: int):
        self._L.call("ctcvr_peer_set_timeout_ms", self.ctx, int(ms))

    def reduce(self, tensors: Sequence[torch.Tensor]):
        """In-place sum over ranks of `tensors` (fp32, contiguous, CUDA) on the current stream."""
        if self.world == 1:
            return
        ts = list(tensors)
        for t in ts:
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() > 0):
                raise RuntimeError("PeerGradExchange.reduce: tensors must be non-empty contiguous fp32 CUDA tensors")
        key = tuple((t.data_ptr(), t.numel()) for t in ts)
        args = self._cache.get(key)
        if args is None:
            if len(self._cache) > 64:
                self._cache.clear()
            args = ((ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts]), (ctypes.c_long * len(ts))(*[t.numel() for t in ts]))
            self._cache[key] = args
        self._L.call("ctcvr_peer_allreduce", self.ctx, args[0], args[1], len(ts), self.ctas, self._L.stream())

    def reduce_grads(self, params: Iterable[torch.nn.Parameter], extra: Sequence[torch.Tensor] = ()):
        grads = []
        for p in params:
            if not p.requires_grad:
                continue
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            grads.append(p.grad)
        self.reduce(grads + list(extra))

    def close(self):
        if getattr(self, "ctx", None) is not None:
            self._L.lib().ctcvr_peer_destroy(self.ctx)
            self.ctx = None


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
