"""Exactly reproducible synthetic tensors for the config-size decode fixtures (decode_cfg.npz).

Weights and encoder frames of BASELINE.json's cfg3 / cfg5 shapes would be tens of MB as stored fixtures, and a seeded
torch.randn is only reproducible for one torch build.  Instead every tensor is a pure integer function of (name, index)
(splitmix64 in numpy uint64 arithmetic), mapped to fp32 values k / 32768 * scale with scale a power of two - exact in
fp32, identical on every machine.  make_golden.py loads them into the UNMODIFIED reference modules to produce the
expected hypotheses; the GPU tests rebuild the same tensors and load them into the ctcvr_b200 modules.
"""
import zlib

import numpy as np

_M1, _M2, _G = np.uint64(0xBF58476D1CE4E5B9), np.uint64(0x94D049BB133111EB), np.uint64(0x9E3779B97F4A7C15)


def _mix(h):
    with np.errstate(over="ignore"):
        h = (h ^ (h >> np.uint64(30))) * _M1
        h = (h ^ (h >> np.uint64(27))) * _M2
        return h ^ (h >> np.uint64(31))


def synth(shape, name, scale):
    """fp32 array of `shape`, values uniform on the grid {-1, ..., 32767/32768} * scale, keyed by `name`."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        h = _mix(np.arange(n, dtype=np.uint64) * _G + np.uint64(zlib.crc32(name.encode())) * _M1)
    k = (h & np.uint64(0xFFFF)).astype(np.int64) - 32768
    return (k.astype(np.float32) / np.float32(32768.0) * np.float32(scale)).reshape(shape)


# scale per parameter kind (powers of two).  The joint / predictor are made "opinionated" (as in decode_small.npz: random
# default-init weights emit until the per-frame cap on every frame), the blank gets a bias so that decoding advances.
def decoder_state(shapes, tag, blank, scales=None, blank_bias=4.0):
    """shapes: {state_dict key: shape} of predictor / joint / ctc head; returns {key: fp32 array}."""
    sc = {"embed.weight": 4.0, "rnn.weight_ih_l0": 0.25, "rnn.weight_hh_l0": 0.0625, "rnn.bias_ih_l0": 0.0625,
          "rnn.bias_hh_l0": 0.0625, "projection.weight": 0.25, "projection.bias": 0.0625,
          "enc_ffn.weight": 0.125, "enc_ffn.bias": 0.0625, "pred_ffn.weight": 0.25, "pred_ffn.bias": 0.0625,
          "ffn_out.weight": 0.5, "ffn_out.bias": 0.0625, "ctc_lo.weight": 0.25, "ctc_lo.bias": 0.0625}
    sc.update(scales or {})
    out = {}
    for k, shp in shapes.items():
        out[k] = synth(tuple(shp), f"{tag}/{k}", sc[k])
        if k in ("ffn_out.bias", "ctc_lo.bias"):
            out[k][blank] += np.float32(blank_bias)
    return out


def predictor_state(shapes, tag):
    """Synthetic RNNPredictor parameters for predictor_small.npz: {state_dict key: fp32 array}, any number of layers."""
    sc = {"embed.weight": 2.0, "weight_ih": 0.25, "weight_hh": 0.125, "bias": 0.0625, "projection.weight": 0.25,
          "projection.bias": 0.0625}
    out = {}
    for k, shp in shapes.items():
        key = next(n for n in sc if n in k)
        out[k] = synth(tuple(shp), f"{tag}/{k}", sc[key])
    return out


def predictor_case(tag, V, H, B, U1):
    """Token ids [B,U1] int64 and the cotangent r [B,U1,H] of a predictor_small.npz case."""
    ys = (synth((B, U1), f"pred/{tag}/ys", 32768.0).astype(np.int64) + 32768) % V
    return ys, synth((B, U1, H), f"pred/{tag}/r", 1.0)


def ctc_logp(B, T, V, blank, name="cfg5/ctc_logp"):
    """[B,T,V] fp32 scores shaped like CTC log-posteriors (a peaked frame distribution with frequent blanks); they are
    exact grid values, not normalised - the prefix beam search only adds and compares them."""
    base = -4.0 + synth((B, T, V), name, 4.0)                                   # -8 .. 0
    peak = (synth((B, T), name + "/peak", 32768.0).astype(np.int64) + 32768) % V
    kind = (synth((B, T), name + "/kind", 32768.0).astype(np.int64) + 32768) % 8
    bi, ti = np.meshgrid(np.arange(B), np.arange(T), indexing="ij")
    out = base - 4.0
    out[bi, ti, peak] = np.where(kind < 3, -2.5, -0.25).astype(np.float32)         # a label peak, sometimes weak
    out[:, :, blank] = np.where(kind < 3, -0.25, -2.0).astype(np.float32)          # blank dominates 3 frames in 8
    return out.astype(np.float32)
