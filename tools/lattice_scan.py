"""Dev tool: lattice kernel time vs T (slope = per-diagonal latency, intercept = staging / copy-out)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctcvr_b200 as C  # noqa
from ctcvr_b200._lib import call, ptr, stream
dev = "cuda"
B, U1 = 32, 41
for T in (10, 60, 125, 250):
    lpb = -torch.rand(B, T, U1, device=dev) * 5; lpl = -torch.rand(B, T, U1, device=dev) * 5
    tl = torch.full((B,), T, dtype=torch.int32, device=dev); ul = torch.full((B,), U1 - 1, dtype=torch.int32, device=dev)
    al = torch.empty_like(lpb); be = torch.empty_like(lpb); costs = torch.empty(B, device=dev)
    def run():
        call("ctcvr_rnnt_lattice", ptr(lpb), ptr(lpl), ptr(tl), ptr(ul), ptr(al), ptr(be), ptr(costs), B, T, U1, stream())
    for _ in range(3): run()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): run()
    e.record(); torch.cuda.synchronize()
    print(f"T={T:4d}: {s.elapsed_time(e)/20*1e3:.1f} us  ({T+U1-1} diagonals)")
