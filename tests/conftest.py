import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def cfg_decoder_weights(fx):
    """Predictor / joint / CTC-head weights and encoder frames of decode_cfg.npz: exact integer-hash tensors rebuilt by
    tests/golden/synth.py (the fixture stores only what the reference decoded from them)."""
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    import synth
    H, V, blank = int(fx["H"]), int(fx["V"]), int(fx["blank"])
    shapes = {
        "predictor": {"embed.weight": (V, H), "rnn.weight_ih_l0": (4 * H, H), "rnn.weight_hh_l0": (4 * H, H),
                      "rnn.bias_ih_l0": (4 * H,), "rnn.bias_hh_l0": (4 * H,), "projection.weight": (H, H),
                      "projection.bias": (H,)},
        "joint": {"enc_ffn.weight": (H, H), "enc_ffn.bias": (H,), "pred_ffn.weight": (H, H), "pred_ffn.bias": (H,),
                  "ffn_out.weight": (V, H), "ffn_out.bias": (V,)},
        "ctc": {"ctc_lo.weight": (V, H), "ctc_lo.bias": (V,)},
    }
    out = {k: synth.decoder_state(v, f"cfg/{k}", blank, scales={"ffn_out.weight": float(fx["ffn_out_scale"])},
                                  blank_bias=float(fx["blank_bias"])) for k, v in shapes.items()}
    out["enc"] = synth.synth((2, 500, H), "cfg/enc", 2.0)
    out["synth"] = synth
    return out


def predictor_case(fx, tag):
    """One case of predictor_small.npz: (dims, {state_dict key: fp32 array}, ys [B,U1] int64, r [B,U1,H]) - the
    weights, token ids and cotangent are rebuilt from tests/golden/synth.py, the fixture holds the reference's results."""
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    import synth
    V, H, L, B, U1 = (int(x) for x in fx[f"{tag}_dims"])
    shapes = {"embed.weight": (V, H)}
    for l in range(L):
        shapes.update({f"rnn.weight_ih_l{l}": (4 * H, H), f"rnn.weight_hh_l{l}": (4 * H, H),
                       f"rnn.bias_ih_l{l}": (4 * H,), f"rnn.bias_hh_l{l}": (4 * H,)})
    shapes.update({"projection.weight": (H, H), "projection.bias": (H,)})
    ys, r = synth.predictor_case(tag, V, H, B, U1)
    return (V, H, L, B, U1), synth.predictor_state(shapes, f"pred/{tag}"), ys, r
