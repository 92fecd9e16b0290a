"""A few eager forward + backward passes of functional.lstm_sequence alone (for an ncu launch list):

    ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 24 --csv --log-file out.csv python tools/lstm_launches.py [B U1 H]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctcvr_b200 import functional as CF  # noqa: E402

B, U1, H = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 41, 512)
torch.manual_seed(0)
lstm = torch.nn.LSTM(H, H, 1, batch_first=True).cuda()
x = torch.randn(B, U1, H, device="cuda", requires_grad=True)
z = torch.zeros(B, H, device="cuda")
r = torch.randn(B, U1, H, device="cuda")
ps = [lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0]
for _ in range(8):
    torch.autograd.backward(CF.lstm_sequence(x, *ps, z, z)[0], r)
    x.grad = None
    for p in ps:
        p.grad = None
torch.cuda.synchronize()
