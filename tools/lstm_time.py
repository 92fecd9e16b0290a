"""Time the predictor LSTM sequence kernels (csrc/lstm_seq.cu) against the library LSTM (cuDNN) on the same GPU.

    python tools/lstm_time.py [B U1 H ...]   # default: cfg2 (B=32, U+1=41, H=512) and cfg1-like (B=32, U+1=25, H=256), fwd and fwd+bwd

CUDA events on the current stream, 20 iterations after 5 warm-ups; prints one JSON line per shape."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctcvr_b200 as C  # noqa: E402
from ctcvr_b200 import functional as CF  # noqa: E402


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2]


def predictor_times(B, U1, H):
    """{ours, cudnn} x {fwd, fwd+bwd} median ms of one LSTM layer [B,U1,H] -> [B,U1,H] in fp32."""
    torch.manual_seed(0)
    lstm = torch.nn.LSTM(H, H, 1, batch_first=True).cuda()
    x = torch.randn(B, U1, H, device="cuda", requires_grad=True)
    h0 = torch.zeros(1, B, H, device="cuda")
    c0 = torch.zeros(1, B, H, device="cuda")
    r = torch.randn(B, U1, H, device="cuda")
    ps = [lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0]

    def ours_f():
        with torch.no_grad():
            return CF.lstm_sequence(x, *ps, h0[0], c0[0])[0]

    def cudnn_f():
        with torch.no_grad():
            return lstm(x, (h0, c0))[0]

    def ours_fb():
        out = CF.lstm_sequence(x, *ps, h0[0], c0[0])[0]
        torch.autograd.backward(out, r)
        x.grad = None
        for p in ps:
            p.grad = None

    def cudnn_fb():
        out = lstm(x, (h0, c0))[0]
        torch.autograd.backward(out, r)
        x.grad = None
        for p in ps:
            p.grad = None

    res = {"B": B, "U1": U1, "H": H}
    for name, fn in (("ours_fwd_ms", ours_f), ("cudnn_fwd_ms", cudnn_f), ("ours_fwd_bwd_ms", ours_fb), ("cudnn_fwd_bwd_ms", cudnn_fb)):
        res[name] = round(timed(fn), 4)
    # the same work replayed from a CUDA graph (what a graphed train step pays: no host launch gaps)
    for name, fn in (("ours_fwd_bwd_graph_ms", ours_fb), ("cudnn_fwd_bwd_graph_ms", cudnn_fb)):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            res[name] = round(timed(g.replay), 4)
        except Exception as e:  # noqa: BLE001
            res[name] = f"capture failed: {type(e).__name__}"
    return res


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    shapes = [tuple(a[i:i + 3]) for i in range(0, len(a), 3)] or [(32, 41, 512), (32, 25, 256), (256, 41, 512)]
    for B, U1, H in shapes:
        print(json.dumps(predictor_times(B, U1, H)), flush=True)
