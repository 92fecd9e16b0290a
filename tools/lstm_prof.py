"""Kernel-only timing of the LSTM sequence kernels through the C ABI (no autograd, no GEMMs): CUDA events around
back-to-back calls of ctcvr_lstm_seq_fwd / ctcvr_lstm_seq_bwd for a few (B, U1, H), to separate the per-step cost from
the per-launch cost (U1 = 1 against U1 = 41).

    python tools/lstm_prof.py [B U1 H ...]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctcvr_b200 as C  # noqa: E402,F401
from ctcvr_b200._lib import call, ptr, query, stream  # noqa: E402


def run(B, U1, H, iters=20):
    dev = "cuda"
    xg = torch.randn(B, U1, 4 * H, device=dev)
    w = torch.randn(4 * H, H, device=dev) * 0.05
    out = torch.empty(B, U1, H, device=dev)
    cs = torch.empty_like(out)
    act = torch.empty(B, U1, 4 * H, device=dev)
    hn, cn = torch.empty(B, H, device=dev), torch.empty(B, H, device=dev)
    d_out = torch.randn(B, U1, H, device=dev)
    dg = torch.empty_like(act)
    dh0, dc0 = torch.empty_like(hn), torch.empty_like(hn)
    ws = torch.empty(query("ctcvr_lstm_seq_ws_bytes", B, H), dtype=torch.uint8, device=dev)

    def fwd():
        call("ctcvr_lstm_seq_fwd", ptr(xg), ptr(w), None, None, ptr(out), ptr(cs), ptr(act), ptr(hn), ptr(cn), B, U1, H,
             ptr(ws), ws.numel(), stream())

    def bwd():
        call("ctcvr_lstm_seq_bwd", ptr(act), ptr(cs), None, ptr(w), ptr(d_out), None, None, ptr(dg), ptr(dh0), ptr(dc0),
             B, U1, H, ptr(ws), ws.numel(), stream())

    res = {"B": B, "U1": U1, "H": H}
    for name, fn in (("fwd_us", fwd), ("bwd_us", bwd)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = round(e0.elapsed_time(e1) / iters * 1000, 1)
    return res


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    shapes = [tuple(a[i:i + 3]) for i in range(0, len(a), 3)] or [(32, 1, 512), (32, 41, 512), (32, 81, 512), (32, 41, 256),
                                                                  (32, 41, 128), (64, 41, 512)]
    for s in shapes:
        print(json.dumps(run(*s)), flush=True)
