"""Dev tool: one cfg2 fwd + lattice + bwd through the C ABI (two rounds), for `ncu -k regex:...` launch lists."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctcvr_b200 as C  # noqa
from ctcvr_b200._lib import call, ptr, query, stream
torch.manual_seed(0)
B, T, U1, D, V, blank = 32, 250, 41, 512, 412, 5
dev = "cuda"
e = torch.randn(B, T, D, device=dev).bfloat16().float(); p = torch.randn(B, U1, D, device=dev).bfloat16().float()
w = torch.randn(V, D, device=dev) / D ** 0.5; b = torch.randn(V, device=dev) * 0.1
tgt = torch.randint(6, V, (B, U1 - 1), dtype=torch.int32, device=dev)
tl = torch.full((B,), T, dtype=torch.int32, device=dev); ul = torch.full((B,), U1 - 1, dtype=torch.int32, device=dev)
lse = torch.empty(B, T, U1, device=dev); lpb = torch.empty_like(lse); lpl = torch.empty_like(lse)
al = torch.empty_like(lse); be = torch.empty_like(lse); costs = torch.empty(B, device=dev)
gc = torch.full((B,), 1.0 / B, device=dev)
d_e, d_p, d_w, d_b = torch.empty_like(e), torch.empty_like(p), torch.empty_like(w), torch.empty_like(b)
wsf = torch.empty(query("ctcvr_joint_rnnt_fwd_ws_bytes", B, T, U1, D, V, 1), dtype=torch.uint8, device=dev)
wsb = torch.empty(query("ctcvr_joint_rnnt_bwd_ws_bytes", B, T, U1, D, V, 1), dtype=torch.uint8, device=dev)
for _ in range(2):
    call("ctcvr_joint_rnnt_fwd", ptr(e), ptr(p), ptr(w), ptr(b), ptr(tgt), ptr(tl), ptr(ul), ptr(lse), ptr(lpb), ptr(lpl),
         B, T, U1, D, V, blank, 1, ptr(wsf), wsf.numel(), stream())
    call("ctcvr_rnnt_lattice", ptr(lpb), ptr(lpl), ptr(tl), ptr(ul), ptr(al), ptr(be), ptr(costs), B, T, U1, stream())
    call("ctcvr_joint_rnnt_bwd", ptr(e), ptr(p), ptr(w), ptr(b), ptr(tgt), ptr(tl), ptr(ul), ptr(lse), ptr(lpb), ptr(lpl), ptr(al), ptr(be),
         ptr(costs), ptr(gc), -1.0, ptr(d_e), ptr(d_p), ptr(d_w), ptr(d_b), B, T, U1, D, V, blank, 1, ptr(wsb), wsb.numel(), stream())
torch.cuda.synchronize()
print("ok", float(costs.sum()))
