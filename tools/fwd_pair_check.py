"""Dev tool: a forward-kernel variant against the single-CTA joint_fwd2_kernel (same bf16 operands) - max differences and
timing.  FWD_MODE selects the variant by its ctcvr_debug_set_mode value (0 = the parked CTA-pair kernel, the default here)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctcvr_b200 as C
import ctcvr_b200._lib as _L
if os.environ.get('CTCVR_LIB'): _L.LIB_PATH = os.environ['CTCVR_LIB']
from ctcvr_b200._lib import call, ptr, query, stream, lib
dev = 'cuda'
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
    ws = torch.empty(query("ctcvr_joint_rnnt_fwd_ws_bytes", B, T, U1, D, V, 1), dtype=torch.uint8, device=dev)
    outs = []
    for mode in (1, int(os.environ.get('FWD_MODE', '0'))):
        lib().ctcvr_debug_set_mode(mode)
        lse = torch.full((B, T, U1), float('nan'), device=dev); lpb = lse.clone(); lpl = lse.clone()
        def f():
            call("ctcvr_joint_rnnt_fwd", ptr(e), ptr(p), ptr(w), ptr(b), ptr(tgt), ptr(tl), ptr(ul), ptr(lse), ptr(lpb), ptr(lpl),
                 B, T, U1, D, V, blank, 1, ptr(ws), ws.numel(), stream())
        f(); torch.cuda.synchronize()
        err = lib().ctcvr_debug_tc_error()
        ms = None
        if time_it:
            flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            ts = []
            for _ in range(7):
                flush.zero_(); a = torch.cuda.Event(enable_timing=True); c = torch.cuda.Event(enable_timing=True)
                a.record(); f(); c.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(c) * 1e3)
            ts.sort(); ms = ts[len(ts) // 2]
        outs.append((lse, lpb, lpl, err, ms))
    lib().ctcvr_debug_set_mode(1)
    mask = (torch.arange(T, device=dev)[None, :, None] < tl[:, None, None]) & (torch.arange(U1, device=dev)[None, None, :] <= ul[:, None, None])
    d = []
    for i, n in enumerate(("lse", "lpb", "lpl")):
        a, c = outs[0][i], outs[1][i]
        m2 = mask if n != "lpl" else mask & (torch.arange(U1, device=dev)[None, None, :] < ul[:, None, None])
        a, c = a[m2], c[m2]
        nan = int(torch.isnan(c).sum())
        d.append(f"{n} maxdiff {float((a - c).abs().max()) if a.numel() else 0.0:.2e} nan {nan}")
    print(f"B{B} T{T} U1{U1} D{D} V{V} ragged={ragged}: " + "; ".join(d) + f"; err single {outs[0][3]:#x} pair {outs[1][3]:#x}" +
          (f"; us single {outs[0][4]:.1f} pair {outs[1][4]:.1f}" if time_it else ""))
import sys
for a in sys.argv[1:]:
    B,T,U1,D,V,r = [int(x) for x in a.split(',')]
    run(B,T,U1,D,V,bool(r & 1), time_it=bool(r & 2) or r == 0)
