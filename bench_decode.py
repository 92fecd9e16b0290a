"""Secondary measurements for the hot-path rows outside the headline metric (A4 CTC loss, A5/A6 greedy,
A7/A8 RNN-T beams (one stream, and A7b/A8b: 296 utterances per launch), A9 CTC prefix beam, A10 CTC greedy) on ONE B200, each next to the reference's CPU arithmetic timed
on a bounded sample on the same box.  One JSON line per row:

    python bench_decode.py [--quick]

Shapes follow BASELINE.json / SURVEY.md §8(d): cfg3 = streaming model H=256, T'=249 encoder frames per utterance,
chunks of 16 frames; cfg5 = beam 10 on T'=500 frames, CTC B=32 x T=500 x V=412, U=40.  RTF is the reference's
definition (online_rnnt_delay.py:50-59): wall time / (input frames x 0.01 s) with 4 input frames per encoder frame.
The CPU side calls the oracle (test infrastructure; same arithmetic as the reference's Python decoders / ATen CTC
loss) - it is a reported baseline, nothing on the GPU side touches it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

V, BLANK = 412, 5
FRAME_S = 0.04          # one encoder frame = 4 input frames of 10 ms


def _sync_time(fn, reps):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


def _decode_model(C, H, seed=0, frames=249):
    """Random-weight predictor + joint with the blank bias calibrated so that greedy decoding emits about one token per
    six frames (a speech-like rate): without a blank bias every frame hits the n_steps cap, with a large
    one nothing is ever emitted and the decoders have no work."""
    torch.manual_seed(seed)
    pred = C.RNNPredictor(V, H, H, 0.0, H, 1, dropout=0.0).cuda().eval()
    joint = C.TransducerJoint(V, H, H, H).cuda().eval()
    m = types.SimpleNamespace(predictor=pred, joint=joint, blank=BLANK)
    enc = torch.randn(8, frames, H, device="cuda")
    lens = torch.full((8,), frames, dtype=torch.int32, device="cuda")
    base = float(joint.ffn_out.bias[BLANK].detach())
    best = None
    for bias in [x * 0.1 for x in range(0, 31)]:
        with torch.no_grad():
            joint.ffn_out.bias[BLANK] = base + bias
        rate = sum(len(h) for h in C.basic_greedy_search(m, enc, lens, n_steps=64)) / (8.0 * frames)
        if best is None or abs(rate - 1 / 6) < abs(best[1] - 1 / 6):
            best = (bias, rate)
    with torch.no_grad():
        joint.ffn_out.bias[BLANK] = base + best[0]
    m.token_rate = best[1]
    return m


def _cpu_weights(m):
    pw = {k: v.detach().cpu() for k, v in m.predictor.state_dict().items()}
    jw = {k: v.detach().cpu() for k, v in m.joint.state_dict().items()}
    return pw, jw


_ROWS = []


def _line(row, what, value, unit, gpu_s, cpu_value, cpu_sample, extra=None):
    d = {"row": row, "what": what, "value": value, "unit": unit, "gpu_seconds": gpu_s, "n_gpus": 1,
         "cpu_baseline": {"value": cpu_value, "unit": unit, "cores": torch.get_num_threads(), "kind": "port",
                          "sample": cpu_sample}}
    if extra:
        d.update(extra)
    _ROWS.append(d)


def run_rows(quick=True):
    """All rows as a list of dicts (bench.py folds them into its JSON line under "decode")."""
    del _ROWS[:]
    main(types.SimpleNamespace(quick=quick))
    return list(_ROWS)


def sharded():
    """cfg3 / cfg5 decode sharded over the ranks of one node (launch under torchrun): utterances are dealt out evenly,
    every rank decodes its shard with the same weights, NO collective on the data path (replicas only);
    time = max over ranks between two barriers.  Rows: A5 (1 000 utterances x 249 frames, greedy) and A7b (1 184
    utterances x 500 frames, online beam 10)."""
    import torch.distributed as dist
    import ctcvr_b200 as C
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    m = _decode_model(C, 256)                     # seeded: identical weights on every rank

    def timed(fn, reps=2):
        fn()
        tot = 0.0
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tot += float(t)
        return tot / reps

    rows = []
    for row, total, frames, what, fn in (
            ("A5", 1000, 249, "RNN-T greedy search, 1000 utterances x 249 frames, H=256, n_steps=64",
             lambda e, l: C.basic_greedy_search(m, e, l, n_steps=64)),
            ("A7b", 1184, 500, "online RNN-T beam search, beam 10, 1184 utterances x 500 frames",
             lambda e, l: C.beam_search_batch(m, e, l, beam_size=10, n_steps=10))):
        lo, hi = rank * total // world, (rank + 1) * total // world
        g = torch.Generator().manual_seed(100 + rank)
        e = torch.randn(hi - lo, frames, 256, generator=g).to(dev)
        l = torch.full((hi - lo,), frames, dtype=torch.int32, device=dev)
        sec = timed(lambda: fn(e, l))
        rows.append({"row": row, "what": what, "value": total / sec, "unit": "utt/s", "gpu_seconds": sec, "n_gpus": world,
                     "rtf": sec / (total * frames * FRAME_S), "sharding": "utterances dealt evenly to ranks, no collective"})
    if rank == 0:
        for r in rows:
            print(json.dumps(r), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main(args=None):
    if args is None:
        ap = argparse.ArgumentParser()
        ap.add_argument("--quick", action="store_true", help="smaller CPU samples")
        ap.add_argument("--sharded", action="store_true", help="under torchrun: A5 / A7b with the utterances sharded over the ranks")
        args = ap.parse_args()
    if getattr(args, "sharded", False):
        return sharded()
    import ctcvr_b200 as C
    from oracle import ctc_oracle as CO
    from oracle import transducer_oracle as TO
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cuda")

    # ------------------------------------------------------------------ A4: CTC loss fwd+bwd at cfg5 (B=32, T=500)
    B, T, U = 32, 500, 40
    torch.manual_seed(5)
    logits = torch.randn(B, T, V)
    ys = torch.randint(6, V, (B, U))
    hl = torch.full((B,), T)
    yl = torch.full((B,), U)
    xg = logits.cuda().requires_grad_(True)
    ysd, hld, yld = ys.cuda(), hl.cuda(), yl.cuda()

    def ctc_step():
        xg.grad = None
        loss, _ = C.ctc_loss_from_logits(xg, ysd, hld, yld, BLANK, "sum")
        loss.backward()
        return loss
    g_s, _ = _sync_time(ctc_step, 20)
    nb = 8
    xr = logits[:nb].clone().requires_grad_(True)

    def ctc_cpu():
        xr.grad = None
        l = torch.nn.functional.ctc_loss(xr.transpose(0, 1).log_softmax(2), ys[:nb], hl[:nb], yl[:nb], blank=BLANK,
                                         reduction="sum", zero_infinity=True)
        l.backward()
    ctc_cpu()
    t0 = time.perf_counter(); ctc_cpu(); ctc_cpu(); c_s = (time.perf_counter() - t0) / 2
    # the same tail through ATen's CUDA kernels (the library path the reference runs with Config.device = "cuda")
    xa = logits.cuda().requires_grad_(True)

    def ctc_aten():
        xa.grad = None
        l = torch.nn.functional.ctc_loss(xa.transpose(0, 1).log_softmax(2), ysd, hld, yld, blank=BLANK, reduction="sum",
                                         zero_infinity=True)
        l.backward()
        return l
    a_s, _ = _sync_time(ctc_aten, 20)
    bytes_alg = 2.0 * B * T * V * 4 + 2.0 * B * T * (2 * U + 1) * 4
    _line("A4", "CTC log-softmax + loss fwd/bwd, cfg5 B=32 T=500 V=412 U=40", B / g_s, "utt/s", g_s, nb / c_s,
          f"{nb} of {B} utterances, ATen CPU ctc_loss fwd+bwd", {"achieved_GBps": bytes_alg / g_s / 1e9,
                                                                "chain_steps": T,
                                                                "gpu_reference_aten_utt_per_s": B / a_s})

    # ------------------------------------------------------------------ A10 / A9: CTC greedy and prefix beam (cfg5)
    lp = torch.log_softmax(logits * 2.0, dim=-1)
    lpd, lnd = lp.cuda(), hl.cuda()
    g_s, _ = _sync_time(lambda: C.ctc_greedy_hyps(lpd, lnd, BLANK), 10)
    t0 = time.perf_counter(); CO.ctc_greedy_search(lp[:4], hl[:4], BLANK); c_s = time.perf_counter() - t0
    _line("A10", "CTC greedy (argmax + collapse) B=32 T=500", B / g_s, "utt/s", g_s, 4 / c_s, "4 of 32 utterances, Python loop",
          {"rtf": g_s / (B * T * FRAME_S)})
    g_s, _ = _sync_time(lambda: C.ctc_prefix_beam_search(lpd, lnd, 10, blank_id=BLANK), 5)
    tq = 60 if args.quick else 125
    t0 = time.perf_counter(); CO.ctc_prefix_beam_search(lp[:1, :tq], torch.tensor([tq]), 10, BLANK); c_s = time.perf_counter() - t0
    _line("A9", "CTC prefix beam search, beam 10, B=32 T=500 V=412", B / g_s, "utt/s", g_s, (tq / T) / c_s,
          f"1 utterance x {tq} of 500 frames (cost is linear in frames), wenet Python algorithm",
          {"rtf": g_s / (B * T * FRAME_S)})

    # ------------------------------------------------------------------ A5: offline greedy, cfg3 model (H=256), 1000 utts
    H = 256
    m = _decode_model(C, H)
    pw, jw = _cpu_weights(m)
    N, Tp = 1000, 249
    torch.manual_seed(6)
    enc = torch.randn(N, Tp, H, device=dev)
    elens = torch.full((N,), Tp, dtype=torch.int32, device=dev)
    g_s, hyps = _sync_time(lambda: C.basic_greedy_search(m, enc, elens, n_steps=64), 3)
    t0 = time.perf_counter(); TO.greedy_search_offline(pw, jw, BLANK, enc[:1].cpu(), torch.tensor([Tp]), 64); c_s = time.perf_counter() - t0
    _line("A5", "RNN-T greedy search, 1000 utterances x 249 frames, H=256, n_steps=64", N / g_s, "utt/s", g_s, 1 / c_s,
          "1 utterance, Python loop of the reference", {"rtf": g_s / (N * Tp * FRAME_S),
                                                        "mean_tokens": sum(len(h) for h in hyps) / N,
                                                        "calibrated_tokens_per_frame": m.token_rate})

    # ------------------------------------------------------------------ A6: streaming greedy, chunks of 16 frames, batch 1
    def stream_one():
        st, last, toks = None, BLANK, []
        for s in range(0, Tp, 16):
            c, st, last = C.greedy_chunk(m, enc[:1, s:s + 16], st, last, n_steps=10)
            toks += c
        return toks
    g_s, _ = _sync_time(stream_one, 3)
    nchunks = (Tp + 15) // 16

    def stream_cpu():
        st, last = None, BLANK
        e = enc[0].cpu()
        for s in range(0, Tp, 16):
            _, st, last = TO.greedy_chunk_streaming(pw, jw, BLANK, e[s:s + 16], st, last, 10)
    t0 = time.perf_counter(); stream_cpu(); c_s = time.perf_counter() - t0
    _line("A6", "streaming greedy, 1 stream, 16-frame chunks (search only, encoder out of scope)", 1 / g_s, "utt/s", g_s,
          1 / c_s, "the same utterance, Python loop of the reference",
          {"rtf": g_s / (Tp * FRAME_S), "ms_per_chunk": 1e3 * g_s / nchunks})

    # ------------------------------------------------------------------ A7 / A8: beams, beam 10, T'=500 (cfg5)
    Tb = 500
    torch.manual_seed(7)
    e5 = torch.randn(1, Tb, H, device=dev)

    def beam_one():
        st, hy = None, None
        for s in range(0, Tb, 16):
            hy, st = C.beam_chunk_online(m, e5[:, s:s + 16], st, beam_size=10, n_steps=10)
        return hy
    g_s, _ = _sync_time(beam_one, 2)
    tq = 16 if args.quick else 48
    t0 = time.perf_counter(); TO.beam_chunk_online(pw, jw, BLANK, e5[0, :tq].cpu(), None, 10, 10); c_s = time.perf_counter() - t0
    _line("A7", "online RNN-T beam search, beam 10, 500 frames in 16-frame chunks", 1 / g_s, "utt/s", g_s, (tq / Tb) / c_s,
          f"first {tq} of 500 frames, Python algorithm of the reference", {"rtf": g_s / (Tb * FRAME_S)})
    a7_cpu = (tq / Tb) / c_s
    # the same search for S utterances in one launch (one CTA per utterance; the reference API is batch 1): whole-job
    # throughput of an offline decode, RTF = wall / total audio
    S = 296                                                      # two waves of CTAs on 148 SMs
    torch.manual_seed(8)
    eS = torch.randn(S, Tb, H, device=dev)
    lS = torch.full((S,), Tb, dtype=torch.int32, device=dev)
    g_s, hyS = _sync_time(lambda: C.beam_search_batch(m, eS, lS, beam_size=10, n_steps=10), 2)
    _line("A7b", f"online RNN-T beam search, beam 10, {S} utterances x 500 frames in ONE launch", S / g_s, "utt/s", g_s, a7_cpu,
          f"first {tq} of 500 frames of one utterance, Python algorithm of the reference",
          {"rtf": g_s / (S * Tb * FRAME_S), "mean_tokens_best": sum(len(h[0].tokens) for h in hyS) / S})
    ctc_w = torch.randn(V, H, device=dev) / H ** 0.5
    ctc_logp = torch.log_softmax(e5[0] @ ctc_w.T, dim=-1)
    g_s, _ = _sync_time(lambda: C.prefix_beam_search(m, e5[0], ctc_logp, beam_size=10), 2)
    t0 = time.perf_counter()
    TO.prefix_beam_search_wenet(pw, jw, BLANK, e5[0, :tq].cpu(), ctc_logp[:tq].cpu(), 10)
    c_s = time.perf_counter() - t0
    _line("A8", "wenet prefix beam search with CTC fusion, beam 10, 500 frames", 1 / g_s, "utt/s", g_s, (tq / Tb) / c_s,
          f"first {tq} of 500 frames, Python algorithm of the reference", {"rtf": g_s / (Tb * FRAME_S)})
    ctcS = torch.log_softmax(eS @ ctc_w.T, dim=-1)
    g_s, _ = _sync_time(lambda: C.prefix_beam_search_batch(m, eS, lS, ctcS, beam_size=10), 2)
    _line("A8b", f"wenet prefix beam search with CTC fusion, beam 10, {S} utterances x 500 frames in ONE launch", S / g_s, "utt/s",
          g_s, (tq / Tb) / c_s, f"first {tq} of 500 frames of one utterance, Python algorithm of the reference",
          {"rtf": g_s / (S * Tb * FRAME_S)})


if __name__ == "__main__":
    main()
    for r in _ROWS:
        print(json.dumps(r), flush=True)
