"""Dev tool: GPU time (CUDA events, 50 back-to-back calls) of the CTC entry points at cfg5 (B=32, T=500, V=412, U=40)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctcvr_b200 as C
from ctcvr_b200._lib import call, ptr, query, stream
torch.manual_seed(5)
B, T, V, U, BLANK = 32, 500, 412, 40, 5
x = torch.randn(B, T, V, device="cuda"); lp = torch.empty_like(x); grad = torch.empty_like(x)
ys = torch.randint(6, V, (B, U), device="cuda"); hl = torch.full((B,), T, dtype=torch.int32, device="cuda"); yl = torch.full((B,), U, dtype=torch.int32, device="cuda")
nll = torch.empty(B, device="cuda")
ws = torch.empty(query("ctcvr_ctc_loss_ws_bytes", B, T, U), dtype=torch.uint8, device="cuda")
def ev(f, n=50):
    for _ in range(5): f()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n * 1e3
t_ls = ev(lambda: call("ctcvr_log_softmax", ptr(x), ptr(lp), B * T, V, stream()))
t_ab = ev(lambda: call("ctcvr_ctc_loss", ptr(lp), ptr(ys), ptr(hl), ptr(yl), None, ptr(nll), None, B, T, V, U, BLANK, 1, ptr(ws), ws.numel(), stream()))
t_all = ev(lambda: call("ctcvr_ctc_loss", ptr(lp), ptr(ys), ptr(hl), ptr(yl), None, ptr(nll), ptr(grad), B, T, V, U, BLANK, 1, ptr(ws), ws.numel(), stream()))
print({"log_softmax_us": round(t_ls, 1), "alpha_beta_us": round(t_ab, 1), "alpha_beta_plus_grad_us": round(t_all, 1), "nll0": float(nll[0])})
