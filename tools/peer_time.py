"""Dev tool (run under torchrun, N ranks of one node): the NVLink peer gradient exchange against NCCL on the bench's
parameter-gradient sizes - microseconds per call, calls enqueued back to back (ranks pace each other), max over ranks."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctcvr_b200.dist import PeerGradExchange
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
sizes = [512 * 512, 512, 512 * 512, 512, 412 * 512, 412, 1]        # the joint's parameters (2.95 MB)
if os.environ.get("PAYLOAD") == "predictor":                         # + the predictor at H = 512: 13.2 MB
    sizes += [412 * 512, 2048 * 512, 2048 * 512, 2048, 2048, 512 * 512, 512]
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
from ctcvr_b200._lib import lib as _lib0
# experiment bits of the exchange kernel (ctcvr_debug_set_mode bits 2..): 1 = extra __threadfence_system before the release
# store, 2 = destinations in rotated order (every rank on a different peer at any moment) instead of 0..N-1
for mode in (0, 1, 2, 3):
    _lib0().ctcvr_debug_set_mode(1 | (mode << 2))
    ex = PeerGradExchange(sum(sizes), ctas=128)
    res[f"peer_mode{mode}_us"] = timeit(lambda: ex.reduce(xs))
    dist.barrier(); ex.close()
_lib0().ctcvr_debug_set_mode(1)
for ctas in (64, 128):
    ex = PeerGradExchange(sum(sizes), ctas=ctas)
    res[f"peer_{ctas}ctas_us"] = timeit(lambda: ex.reduce(xs))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): ex.reduce(xs)
    res[f"peer_{ctas}ctas_graph_us"] = timeit(g.replay, 20) / 10
    dist.barrier(); ex.close()
# phase stamps of one call (globaltimer ns, thread 0 of every CTA): scatter | barrier A | reduce | barrier B | unpack
from ctcvr_b200._lib import lib
import ctypes
ex = PeerGradExchange(sum(sizes), ctas=64)
prof = torch.zeros(128 * 8, dtype=torch.int64, device=dev)
for _ in range(5): ex.reduce(xs)
dist.barrier(); torch.cuda.synchronize()
lib().ctcvr_debug_set_prof(ctypes.c_void_p(prof.data_ptr()))
ex.reduce(xs); ex.reduce(xs); torch.cuda.synchronize()
lib().ctcvr_debug_set_prof(None)
p_ = prof.view(128, 8)[:64, :6].cpu().double()
t0 = p_[:, 0].min()
if rank == 0:
    print("phase end (us after first CTA start), mean / max over CTAs:",
          [(round(float((p_[:, k] - t0).mean()) / 1e3, 1), round(float((p_[:, k] - t0).max()) / 1e3, 1)) for k in range(6)])
dist.barrier(); ex.close()
if rank == 0: print({k: round(v, 1) for k, v in res.items()}, "world", dist.get_world_size(), "MB", sum(sizes) * 4 / 1e6)
dist.barrier(); dist.destroy_process_group()
