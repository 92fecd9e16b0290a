"""Data-parallel step on real GPUs over NCCL (skipped with fewer than 2 GPUs): the summed gradients of N ranks on
shards of a batch equal the single-process gradients on the whole batch, through the NCCL branch of
`GradAllReducer` (grouped in-place collective) behind an eager step and behind a graphed step, and through the NVLink
peer exchange (`PeerGradExchange`) alone and captured inside the step graph."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch.distributed as dist
    import ctcvr_b200 as C
    from ctcvr_b200.dist import GradAllReducer, PeerGradExchange, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(5)
    B, T, U, D, V, blank = 8, 40, 9, 128, 50, 5
    joint = C.TransducerJoint(V, D, D, D).to(dev)
    enc, pred = torch.randn(B, T, D), torch.randn(B, U + 1, D)
    tgt = torch.randint(6, V, (B, U), dtype=torch.int32)
    tl, ul = torch.full((B,), T, dtype=torch.int32), torch.full((B,), U, dtype=torch.int32)
    lo, hi = shard_bounds(B, world, rank)
    sh = lambda t: t[lo:hi].to(dev)
    res = {}
    # (1) eager step + GradAllReducer (NCCL branch: grouped in-place collective)
    red = GradAllReducer(joint.parameters())
    joint.zero_grad(set_to_none=True)
    costs = joint.rnnt_loss_fused(sh(enc), sh(pred), sh(tgt), sh(tl), sh(ul), blank, reduction="none", precision="bf16")
    (costs.sum() / B).backward()
    red.reduce()
    res["eager"] = {n: p.grad.detach().float().cpu() for n, p in joint.named_parameters()}
    # an autograd graph kept alive over the same parameters pins their AccumulateGrad nodes to the default stream, which
    # a capture on another stream may not touch (cudaErrorStreamCaptureImplicit)
    del costs
    # (2) graphed step, the collective behind the replay
    g = C.GraphedJointRnntStep(joint, hi - lo, T, U, blank, global_batch=B, precision="bf16")
    g.step(sh(enc), sh(pred), sh(tgt), sh(tl), sh(ul))
    red.reduce()
    torch.cuda.synchronize()
    res["graph"] = {n: p.grad.detach().float().cpu() for n, p in joint.named_parameters()}
    # (3) the NVLink peer exchange (csrc/peer_reduce.cu) against NCCL on the same tensors: a + b is commutative, so at
    # world size 2 the sums must be bit-identical
    sizes = (412 * 512, 412, 1, 77, 512 * 512)
    ex = PeerGradExchange(max(sum(sizes), sum(p.numel() for p in joint.parameters()) + 1))
    torch.manual_seed(100 + rank)
    xs = [torch.randn(n, device=dev) for n in sizes]
    want = [x.clone() for x in xs]
    for w in want:
        dist.all_reduce(w)
    for _ in range(3):                                   # repeated calls: the flags are step counters, nothing is reset
        got = [x.clone() for x in xs]
        ex.reduce(got)
        torch.cuda.synchronize()
        res.setdefault("peer_exact", True)
        res["peer_exact"] = res["peer_exact"] and all(torch.equal(a, b) for a, b in zip(got, want))
    # (4) graphed step with the exchange captured as the last kernel of the graph, replayed twice
    g2 = C.GraphedJointRnntStep(joint, hi - lo, T, U, blank, global_batch=B, precision="bf16", grad_exchange=ex)
    for _ in range(2):
        loss = g2.step(sh(enc), sh(pred), sh(tgt), sh(tl), sh(ul))
    torch.cuda.synchronize()
    res["graph_peer"] = {n: p.grad.detach().float().cpu() for n, p in joint.named_parameters()}
    res["graph_peer_loss"] = float(loss)
    if rank == 0:
        # single-process truth on the whole batch
        joint.zero_grad(set_to_none=True)
        costs = joint.rnnt_loss_fused(enc.to(dev), pred.to(dev), tgt.to(dev), tl.to(dev), ul.to(dev), blank, reduction="none",
                                      precision="bf16")
        (costs.sum() / B).backward()
        res["single"] = {n: p.grad.detach().float().cpu() for n, p in joint.named_parameters()}
        res["single_loss"] = float(costs.sum() / B)
        torch.save(res, out)
    dist.barrier()
    ex.close()
    dist.destroy_process_group()


def test_nccl_dp_step_matches_single_process(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["peer_exact"]
    assert abs(res["graph_peer_loss"] - res["single_loss"]) < 2e-3 * abs(res["single_loss"])
    for kind in ("eager", "graph", "graph_peer"):
        for n, want in res["single"].items():
            got = res[kind][n]
            err = float((got - want).norm() / max(float(want.norm()), 1e-30))
            assert err < 1e-3, (kind, n, err)      # bf16 activation gradients: shard boundaries move single roundings
