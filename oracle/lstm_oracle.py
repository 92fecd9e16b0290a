"""Oracle (TEST INFRASTRUCTURE ONLY) for the predictor's LSTM recurrence.

The reference delegates the arithmetic to `torch.nn.LSTM` (model/component/predictor.py:29-36,58: one layer,
batch_first, gate order i, f, g, o as documented for torch.nn.LSTM); this file restates one layer in plain numpy,
forward AND the hand-derived backward - the same decomposition the kernels of ctc-vr_b200/csrc/lstm_seq.cu use
(sequential part -> gate gradients; the weight / input gradients are products over all B*U1 rows afterwards) - so that
the intermediate `dgates` can be compared too.  Parity pinning: tests/test_oracle_cpu.py checks it against torch's own
LSTM autograd in fp64 and against tests/golden/predictor_small.npz (produced by the reference RNNPredictor).

`split_tf32_product` restates the operand split in front of the plain GEMMs (csrc/lstm_seq.cu::split_tf32_kernel).
"""
from __future__ import annotations

import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_layer_forward(x, w_ih, w_hh, b_ih, b_hh, h0, c0):
    """x [B,U1,E]; w_ih [4H,E]; w_hh [4H,H]; biases [4H] or None; h0 / c0 [B,H].
    Returns (out [B,U1,H], h_n, c_n, cache).  torch.nn.LSTM semantics:
        i, f, g, o = split(x_t W_ih^T + b_ih + h_{t-1} W_hh^T + b_hh);  c_t = s(f) c_{t-1} + s(i) tanh(g);  h_t = s(o) tanh(c_t)"""
    B, U1, _ = x.shape
    H = w_hh.shape[1]
    bias = 0.0 if b_ih is None else (b_ih + b_hh)
    xg = x @ w_ih.T + bias                                   # the input projection of all steps (one GEMM in the product)
    out = np.zeros((B, U1, H), dtype=x.dtype)
    cs = np.zeros_like(out)
    act = np.zeros((B, U1, 4 * H), dtype=x.dtype)
    h, c = h0, c0
    for t in range(U1):
        g = xg[:, t] + h @ w_hh.T
        gi, gf, gg, go = _sigmoid(g[:, :H]), _sigmoid(g[:, H:2 * H]), np.tanh(g[:, 2 * H:3 * H]), _sigmoid(g[:, 3 * H:])
        c = gf * c + gi * gg
        h = go * np.tanh(c)
        out[:, t], cs[:, t] = h, c
        act[:, t] = np.concatenate([gi, gf, gg, go], axis=1)
    return out, h, c, dict(x=x, w_ih=w_ih, w_hh=w_hh, h0=h0, c0=c0, out=out, cs=cs, act=act, has_bias=b_ih is not None)


def lstm_layer_backward(cache, d_out, d_hn=None, d_cn=None):
    """Gradients of sum(out * d_out) + sum(h_n * d_hn) + sum(c_n * d_cn).  Returns a dict with dgates [B,U1,4H]
    (dL / d pre-activation gates, what ctcvr_lstm_seq_bwd writes), dx, dW_ih, dW_hh, db (= db_ih = db_hh), dh0, dc0."""
    x, w_ih, w_hh, h0, c0, out, cs, act = (cache[k] for k in ("x", "w_ih", "w_hh", "h0", "c0", "out", "cs", "act"))
    B, U1, H = out.shape
    dh = np.zeros((B, H), dtype=out.dtype) if d_hn is None else d_hn.copy()
    dc = np.zeros((B, H), dtype=out.dtype) if d_cn is None else d_cn.copy()
    dgates = np.zeros_like(act)
    for t in range(U1 - 1, -1, -1):
        gi, gf, gg, go = act[:, t, :H], act[:, t, H:2 * H], act[:, t, 2 * H:3 * H], act[:, t, 3 * H:]
        c_prev = cs[:, t - 1] if t > 0 else c0
        dht = dh + d_out[:, t]
        tc = np.tanh(cs[:, t])
        dct = dc + dht * go * (1.0 - tc * tc)
        d_o = dht * tc * go * (1.0 - go)
        d_i = dct * gg * gi * (1.0 - gi)
        d_f = dct * c_prev * gf * (1.0 - gf)
        d_g = dct * gi * (1.0 - gg * gg)
        dg = np.concatenate([d_i, d_f, d_g, d_o], axis=1)
        dgates[:, t] = dg
        dh = dg @ w_hh                                        # the recurrent part: dL/dh_{t-1}
        dc = dct * gf
    dg2 = dgates.reshape(B * U1, 4 * H)
    h_prev = np.concatenate([h0[:, None], out[:, :-1]], axis=1).reshape(B * U1, H)
    res = dict(dgates=dgates, dx=(dg2 @ w_ih).reshape(x.shape), dW_ih=dg2.T @ x.reshape(B * U1, -1), dW_hh=dg2.T @ h_prev,
               dh0=dh, dc0=dc)
    res["db"] = dg2.sum(0) if cache["has_bias"] else None
    return res


def _tf32_trunc(a):
    """fp32 value with the low 13 mantissa bits cleared (what `v & 0xffffe000` does in split_tf32_kernel)."""
    return (np.asarray(a, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def split_tf32_product(a, b, terms: int = 3):
    """a [M,K] @ b [K,N] the way the product computes the plain GEMMs around the recurrence: each operand is written as
    hi + lo (hi = truncated to TF32's 10 mantissa bits, lo = the exact fp32 remainder); the tensor cores see TF32
    operands (lo is truncated once more - the worst case of their input rounding) and accumulate
    a_hi b_hi + a_hi b_lo + a_lo b_hi (terms=3).  terms=1 is a plain TF32 GEMM (what torch's cuDNN LSTM uses by default).
    Accumulation in fp64 here: the restatement bounds the OPERAND error, which is what differs between the two."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    a_hi, b_hi = _tf32_trunc(a), _tf32_trunc(b)
    if terms == 1:
        return a_hi.astype(np.float64) @ b_hi.astype(np.float64)
    a_lo, b_lo = _tf32_trunc(a - a_hi), _tf32_trunc(b - b_hi)
    a3 = np.concatenate([a_hi, a_hi, a_lo], axis=1).astype(np.float64)       # pattern 0: (hi | hi | lo) along K
    b3 = np.concatenate([b_hi, b_lo, b_hi], axis=0).astype(np.float64)       # pattern 1: (hi ; lo ; hi) along K
    return a3 @ b3
