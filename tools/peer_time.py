"""Dev tool (run under torchrun, N ranks of one node): the NVLink peer gradient exchange against NCCL on the bench's
parameter-gradient sizes - microseconds per call, calls enqueued back to back (ranks pace each other), max over ranks."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctcvr_b200.dist import PeerGradExchange
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
sizes = [512 * 512, 512, 512 * 512, 512, 412 * 512, 412, 1]
xs = [torch.randn(n, device=dev) for n in sizes]
def timeit(f, n=200):
    for _ in range(10): f()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n * 1e3], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
def nccl():
    with dist._coalescing_manager(device=dev, async_ops=False):
        for x in xs: dist.all_reduce(x)
flat = torch.randn(sum(sizes), device=dev)
res = {"nccl_grouped_us": timeit(nccl), "nccl_flat_us": timeit(lambda: dist.all_reduce(flat))}
for ctas in (16, 32, 64, 128):
    ex = PeerGradExchange(sum(sizes), ctas=ctas)
    res[f"peer_{ctas}ctas_us"] = timeit(lambda: ex.reduce(xs))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): ex.reduce(xs)
    res[f"peer_{ctas}ctas_graph_us"] = timeit(g.replay, 20) / 10
    dist.barrier(); ex.close()
if rank == 0: print({k: round(v, 1) for k, v in res.items()}, "world", dist.get_world_size(), "MB", sum(sizes) * 4 / 1e6)
dist.barrier(); dist.destroy_process_group()
