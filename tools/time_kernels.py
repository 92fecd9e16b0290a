"""Dev tool: CUDA-event timing of the C-ABI entry points at cfg2 (B=32,T=250,U=40,D=512,V=412), L2 flushed
between launches, plus a check of the bf16 forward against the fp32 SIMT forward."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctcvr_b200 as C  # noqa: E402,F401
from ctcvr_b200._lib import call, lib, ptr, query, stream  # noqa: E402

torch.manual_seed(0)
B, T, U1, D, V, blank = 32, 250, 41, 512, 412, 5
ragged = "--ragged" in sys.argv
dev = "cuda"
e = torch.randn(B, T, D, device=dev).bfloat16().float()
p = torch.randn(B, U1, D, device=dev).bfloat16().float()
w = torch.randn(V, D, device=dev) / D ** 0.5
b = torch.randn(V, device=dev) * 0.1
tgt = torch.randint(6, V, (B, U1 - 1), dtype=torch.int32, device=dev)
if ragged:
    tl = torch.randint(125, 251, (B,), dtype=torch.int32, device=dev)
    ul = torch.randint(20, 41, (B,), dtype=torch.int32, device=dev)
    tl[0], ul[0] = T, U1 - 1
else:
    tl = torch.full((B,), T, dtype=torch.int32, device=dev)
    ul = torch.full((B,), U1 - 1, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bufs():
    lse = torch.full((B, T, U1), float("nan"), device=dev)
    return lse, lse.clone(), lse.clone()


def fwd(prec, out, ws):
    lse, lpb, lpl = out
    call("ctcvr_joint_rnnt_fwd", ptr(e), ptr(p), ptr(w), ptr(b), ptr(tgt), ptr(tl), ptr(ul), ptr(lse), ptr(lpb),
         ptr(lpl), B, T, U1, D, V, blank, prec, ptr(ws), ws.numel(), stream())


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        t.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(t)
    return tot / reps


o32, o16 = bufs(), bufs()
ws32 = torch.empty(max(256, query("ctcvr_joint_rnnt_fwd_ws_bytes", B, T, U1, D, V, 0)), dtype=torch.uint8, device=dev)
ws16 = torch.empty(max(256, query("ctcvr_joint_rnnt_fwd_ws_bytes", B, T, U1, D, V, 1)), dtype=torch.uint8, device=dev)
fwd(0, o32, ws32)
fwd(1, o16, ws16)
torch.cuda.synchronize()
print("tc error flag: 0x%08x" % lib().ctcvr_debug_tc_error())
for name, a, c in zip(("lse", "lp_blank", "lp_label"), o32, o16):
    worst = 0.0
    nbad = 0
    for bb in range(B):
        tb, ub = int(tl[bb]), int(ul[bb])
        x, y = a[bb, :tb, :ub + 1], c[bb, :tb, :ub + 1]
        if name == "lp_label":
            x, y = x[:, :ub], y[:, :ub]
        nbad += int((~torch.isfinite(y)).sum())
        if x.numel():
            worst = max(worst, float((x - y).abs().nan_to_num(1e9).max()))
    print(f"fwd bf16 vs fp32 {name}: max abs err {worst:.4e}, non-finite {nbad}")

t_f = timeit(lambda: fwd(1, o16, ws16))
M = float(tl.sum()) * 0 + sum(int(tl[i]) * (int(ul[i]) + 1) for i in range(B))
print(f"fwd bf16: {t_f*1e3:.1f} us  -> {2*M*D*V/(t_f*1e-3)/1e12:.1f} TFLOP/s algorithmic")

lse, lpb, lpl = o16
al, be = torch.empty_like(lse), torch.empty_like(lse)
costs = torch.empty(B, device=dev)


def lat():
    call("ctcvr_rnnt_lattice", ptr(lpb), ptr(lpl), ptr(tl), ptr(ul), ptr(al), ptr(be), ptr(costs), B, T, U1, stream())


t_l = timeit(lat)
print(f"lattice: {t_l*1e3:.1f} us")
gc = torch.full((B,), 1.0 / B, device=dev)
d_e, d_p, d_w, d_b = torch.empty_like(e), torch.empty_like(p), torch.empty_like(w), torch.empty_like(b)
wsb = torch.empty(query("ctcvr_joint_rnnt_bwd_ws_bytes", B, T, U1, D, V, 1), dtype=torch.uint8, device=dev)


def bwd():
    call("ctcvr_joint_rnnt_bwd", ptr(e), ptr(p), ptr(w), ptr(b), ptr(tgt), ptr(tl), ptr(ul), ptr(lse), ptr(lpb), ptr(lpl), ptr(al), ptr(be),
         ptr(costs), ptr(gc), -1.0, ptr(d_e), ptr(d_p), ptr(d_w), ptr(d_b), B, T, U1, D, V, blank, 1, ptr(wsb),
         wsb.numel(), stream())


t_b = timeit(bwd)
print(f"bwd bf16: {t_b*1e3:.1f} us  -> {4*M*D*V/(t_b*1e-3)/1e12:.1f} TFLOP/s algorithmic")
print("tc error flag: 0x%08x" % lib().ctcvr_debug_tc_error())
tot = t_f + t_l + t_b
print(f"sum: {tot*1e3:.1f} us -> {B/(tot*1e-3):.0f} utt/s (kernels only)")
