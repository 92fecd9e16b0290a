"""Dev tool: CTA-pair backward kernel against the single-CTA kernel (same bf16 operands) - differences and timing."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctcvr_b200 as C
import ctcvr_b200._lib as _L
if os.environ.get('CTCVR_LIB'): _L.LIB_PATH = os.environ['CTCVR_LIB']
from ctcvr_b200._lib import call, ptr, query, stream, lib
dev = 'cuda'
def rel(a, b): return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))
def run(B, T, U1, D, V, ragged, seed=0, time_it=False):
    torch.manual_seed(seed)
    blank = 5 if V > 5 else 0
    e = torch.randn(B, T, D, device=dev); p = torch.randn(B, U1, D, device=dev)
    w = torch.randn(V, D, device=dev) / D ** 0.5; b = torch.randn(V, device=dev) * 0.1
    tgt = torch.randint(0, V, (B, max(U1 - 1, 1)), dtype=torch.int32, device=dev)[:, :U1 - 1].contiguous()
    if ragged:
        tl = torch.randint(max(1, T // 2), T + 1, (B,), dtype=torch.int32, device=dev); tl[0] = T
        ul = torch.randint(max(0, (U1 - 1) // 2), U1, (B,), dtype=torch.int32, device=dev); ul[0] = U1 - 1
    else:
        tl = torch.full((B,), T, dtype=torch.int32, device=dev); ul = torch.full((B,), U1 - 1, dtype=torch.int32, device=dev)
    lse = torch.empty(B, T, U1, device=dev); lpb = torch.empty_like(lse); lpl = torch.empty_like(lse)
    ws = torch.empty(query("ctcvr_joint_rnnt_fwd_ws_bytes", B, T, U1, D, V, 1), dtype=torch.uint8, device=dev)
    call("ctcvr_joint_rnnt_fwd", ptr(e), ptr(p), ptr(w), ptr(b), ptr(tgt), ptr(tl), ptr(ul), ptr(lse), ptr(lpb), ptr(lpl),
         B, T, U1, D, V, blank, 1, ptr(ws), ws.numel(), stream())
    al = torch.empty_like(lse); be = torch.empty_like(lse); costs = torch.empty(B, device=dev)
    call("ctcvr_rnnt_lattice", ptr(lpb), ptr(lpl), ptr(tl), ptr(ul), ptr(al), ptr(be), ptr(costs), B, T, U1, stream())
    gc = torch.full((B,), 1.0 / B, device=dev)
    wsb = torch.empty(query("ctcvr_joint_rnnt_bwd_ws_bytes", B, T, U1, D, V, 1), dtype=torch.uint8, device=dev)
    outs = []
    for mode in (3, 1):                 # bit 1: single-CTA backward
        lib().ctcvr_debug_set_mode(mode)
        d_e = torch.full_like(e, float('nan')); d_p = torch.full_like(p, float('nan'))
        d_w = torch.full_like(w, float('nan')); d_b = torch.full_like(b, float('nan'))
        def f():
            call("ctcvr_joint_rnnt_bwd", ptr(e), ptr(p), ptr(w), ptr(b), ptr(tgt), ptr(tl), ptr(ul), ptr(lse), ptr(lpb), ptr(lpl),
                 ptr(al), ptr(be), ptr(costs), ptr(gc), -1.0, ptr(d_e), ptr(d_p), ptr(d_w), ptr(d_b), B, T, U1, D, V, blank, 1,
                 ptr(wsb), wsb.numel(), stream())
        f(); torch.cuda.synchronize()
        err = lib().ctcvr_debug_tc_error()
        us = None
        if time_it:
            flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            ts = []
            for _ in range(7):
                flush.zero_(); a = torch.cuda.Event(enable_timing=True); c = torch.cuda.Event(enable_timing=True)
                a.record(); f(); c.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(c) * 1e3)
            ts.sort(); us = ts[len(ts) // 2]
        outs.append((d_e, d_p, d_w, d_b, err, us))
    lib().ctcvr_debug_set_mode(1)
    d = [f"{n} rel {rel(outs[1][i], outs[0][i]):.2e} nan {int(torch.isnan(outs[1][i]).sum())}" for i, n in enumerate(("d_enc", "d_pred", "d_w", "d_b"))]
    print(f"B{B} T{T} U1{U1} D{D} V{V} ragged={ragged}: " + "; ".join(d) + f"; err single {outs[0][4]:#x} pair {outs[1][4]:#x}" +
          (f"; us single {outs[0][5]:.1f} pair {outs[1][5]:.1f}" if time_it else ""))
for a in sys.argv[1:]:
    B, T, U1, D, V, r = [int(x) for x in a.split(',')]
    run(B, T, U1, D, V, bool(r), time_it=(r == 0))
