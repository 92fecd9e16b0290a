#!/usr/bin/env python
"""bench.py — fused joint + RNN-T loss forward/backward throughput (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload = BASELINE.json configs[1]: synthetic B=32, T=250 (post-subsampling), U=40, H=D=512, V=412 per
GPU ("weak" scaling: every rank holds its own 32 utterances; the only collective is the gradient
all-reduce of the data-parallel step).  A step = joint pre-projections + fused joint/log-softmax forward
+ lattice + fused backward (+ gradient all-reduce when N>1).  `value` is measured with inputs resident in
HBM; `e2e` goes through the same public API with HOST (pinned) inputs and a device->host read of the loss
inside the timed region.  L2 is flushed (256 MiB write) before every timed step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(B=32, T=250, U=40, D=512, V=412, blank=5)
METRIC = "fused joint+RNN-T loss fwd/bwd throughput"
# kernels of libctcvr.so inside one captured step (bf16 activations in place), as counted by ctcvr_launch_count() on
# an eager step:
KERNELS_PER_GRAPHED_STEP = 8    # fwd: prep, joint_fwd2, lattice | bwd: prep, joint_bwd2, reduce_denc, dw_gemm_rz, reduce_dw
UNIT = "utt/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def make_inputs(seed, device, pinned=False):
    g = torch.Generator().manual_seed(seed)
    B, T, U, D, V = CFG["B"], CFG["T"], CFG["U"], CFG["D"], CFG["V"]
    enc = torch.randn(B, T, D, generator=g)
    pred = torch.randn(B, U + 1, D, generator=g)
    tgt = torch.randint(6, V, (B, U), generator=g, dtype=torch.int32)
    tl = torch.full((B,), T, dtype=torch.int32)
    ul = torch.full((B,), U, dtype=torch.int32)
    ts = [enc, pred, tgt, tl, ul]
    if pinned:
        return [t.pin_memory() for t in ts]
    return [t.to(device) for t in ts]


def cpu_reference_step(nb, threads):
    """The reference's own PyTorch CPU path for this seam (joint.py:48-69 -> torchaudio rnnt_loss ->
    backward), restated in oracle/transducer_oracle.py; bounded sample of `nb` utterances."""
    from oracle import transducer_oracle as TO
    torch.set_num_threads(threads)
    torch.manual_seed(1234)
    D, V, T, U = CFG["D"], CFG["V"], CFG["T"], CFG["U"]
    lin = lambda o, i: torch.nn.Linear(i, o)
    mods = {"enc_ffn": lin(D, D), "pred_ffn": lin(D, D), "ffn_out": lin(V, D)}
    w = {f"{k}.{n}": p.detach() for k, m in mods.items() for n, p in m.named_parameters()}
    enc, pred = torch.randn(nb, T, D), torch.randn(nb, U + 1, D)
    tgt = torch.randint(6, V, (nb, U), dtype=torch.int32)
    tl, ul = torch.full((nb,), T, dtype=torch.int32), torch.full((nb,), U, dtype=torch.int32)

    def step():
        t0 = time.perf_counter()
        TO.fused_joint_rnnt_reference_call(enc, pred, w, tgt, tl, ul, CFG["blank"])
        return time.perf_counter() - t0
    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    nb = 8
    step = cpu_reference_step(nb, threads)
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    ts = [step() for _ in range(args.steps)]
    sec = sum(ts) / len(ts)
    val = nb / sec
    sample = f"{nb} of {CFG['B']} utterances per step (same T/U/H/V), {args.steps} timed steps, torch CPU + torchaudio"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: B=32,T=250,U=40,H=512,V=412 joint+rnnt_loss fwd/bwd (CPU sample)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args):
    import torch.distributed as dist
    import ctcvr_b200 as C
    from ctcvr_b200 import _lib
    from ctcvr_b200.dist import GradAllReducer
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T, U, D, V, blank = (CFG[k] for k in ("B", "T", "U", "D", "V", "blank"))
    torch.manual_seed(1234)
    joint = C.TransducerJoint(V, D, D, D).to(dev)
    reducer = GradAllReducer(joint.parameters()) if world > 1 else None
    enc, pred, tgt, tl, ul = make_inputs(1234 + rank, dev)
    enc.requires_grad_(True)
    pred.requires_grad_(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gB = B * world

    use_graph = not args.no_graph
    graphed = C.GraphedJointRnntStep(joint, B, T, U, blank, global_batch=gB, precision=args.precision) if use_graph else None
    if graphed is not None:
        graphed.load(enc.detach(), pred.detach(), tgt, tl, ul)

    def step(e, p, tg, tl_, ul_):
        """One training step of the seam through the public API.  Graph mode: the whole step (pre-projections, fused
        forward, lattice, fused backward, projection backward) is one captured CUDA graph; inputs are copied into its
        resident buffers (e is None = already resident)."""
        if graphed is not None:
            loss = graphed.step(e, p, tg, tl_, ul_) if e is not None else graphed.step()
        else:
            joint.zero_grad(set_to_none=True)
            e.grad = p.grad = None
            costs = joint.rnnt_loss_fused(e, p, tg, tl_, ul_, blank, reduction="none", precision=args.precision)
            loss = costs.sum() / gB
            loss.backward()
        if reducer is not None:
            reducer.reduce()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    resident = (None, None, None, None, None) if graphed is not None else (enc, pred, tgt, tl, ul)
    for _ in range(args.warmup):
        step(*resident)
    barrier()
    l0 = _lib.lib().ctcvr_launch_count()
    with ClockSampler(local) as clk:
        tot_ms = 0.0
        for _ in range(args.steps):
            flush.zero_()
            barrier()
            s, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            loss = step(*resident)
            e_.record()
            torch.cuda.synchronize()
            tot_ms += s.elapsed_time(e_)
        barrier()
        launches = _lib.lib().ctcvr_launch_count() - l0
        if graphed is not None:      # replays do not pass through the C ABI: count the captured launches
            launches = args.steps * KERNELS_PER_GRAPHED_STEP
        # ---- e2e: host (pinned) inputs -> H2D -> step -> D2H of the loss, every step.  As in a real input pipeline the
        # H2D copy of step i+1 runs on a copy stream while step i computes (double-buffered staging) and the loss of
        # step i is read back while step i+1 runs; every step still pays its own copy and its own loss read-back, and
        # the K steps are timed as one region on the wall clock.
        h = make_inputs(1234 + rank, dev, pinned=True)
        h2d = sum(t.numel() * t.element_size() for t in h)
        copy_stream = torch.cuda.Stream(device=dev)
        # graph mode: one captured step per staging slot, so the H2D copies land directly in the graph's own input
        # buffers (no device-to-device copy in front of the replay); eager mode: plain staging tensors
        graphs = None
        if graphed is not None:
            graphs = [graphed, C.GraphedJointRnntStep(joint, B, T, U, blank, global_batch=gB, precision=args.precision)]
            staging = [g.input_buffers() for g in graphs]
        else:
            staging = [[torch.empty_like(t, device=dev) for t in h] for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        @torch.no_grad()
        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i % 2])
                for dst, src in zip(staging[i % 2], h):
                    dst.copy_(src, non_blocking=True)
                ready[i % 2].record(copy_stream)

        def e2e_run(nsteps):
            """K steps; step i's loss is copied to pinned host memory right behind the step and read by the host one
            step later (after step i+1 has been enqueued), so the device never waits for the host between steps."""
            for ev in consumed:
                ev.record(torch.cuda.current_stream())
            prefetch(0)
            host_loss = torch.empty(nsteps, dtype=torch.float32).pin_memory()
            done = [torch.cuda.Event() for _ in range(nsteps)]
            losses = []
            for i in range(nsteps):
                if i + 1 < nsteps:
                    prefetch(i + 1)
                torch.cuda.current_stream().wait_event(ready[i % 2])
                d = staging[i % 2]
                if graphs is not None:
                    lv = graphs[i % 2].step()
                    if reducer is not None:
                        reducer.reduce()
                else:
                    e_in, p_in = d[0].detach().requires_grad_(True), d[1].detach().requires_grad_(True)
                    lv = step(e_in, p_in, d[2], d[3], d[4])
                consumed[i % 2].record(torch.cuda.current_stream())
                host_loss[i:i + 1].copy_(lv.detach().reshape(1), non_blocking=True)      # D2H read of the step's loss
                done[i].record(torch.cuda.current_stream())
                if i >= 1:
                    done[i - 1].synchronize()
                    losses.append(float(host_loss[i - 1]))
            done[nsteps - 1].synchronize()
            losses.append(float(host_loss[nsteps - 1]))
            torch.cuda.synchronize()
            return losses

        e2e_run(2)
        # the box's pinned-host -> device rate for this step's inputs (explains e2e when the copy, not the step, bounds it)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(copy_stream), torch.no_grad():
            c0.record(copy_stream)
            for dst, src in zip(staging[0], h):
                dst.copy_(src, non_blocking=True)
            c1.record(copy_stream)
        torch.cuda.synchronize()
        h2d_gbps = h2d / (c0.elapsed_time(c1) * 1e-3) / 1e9
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        e2e_run(args.steps)
        e2e_ms = (time.perf_counter() - t0) * 1e3
    ms = torch.tensor([tot_ms / args.steps, e2e_ms / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step, ms_e2e = float(ms[0]), float(ms[1])

    # ---- per-kernel roofline (rank 0): time the C-ABI entry points alone with CUDA events
    line = None
    if rank == 0:
        pk = _peaks()
        kern = time_kernels(C, joint, enc, pred, tgt, tl, ul, blank, args.precision, flush)
        M = B * T * (U + 1)
        flops_bwd = 4.0 * M * D * V
        flops_fwd = 2.0 * M * D * V
        peak = pk["tf_sust"] if args.precision == "bf16" else None
        ach_bwd = flops_bwd / (kern["bwd_ms"] * 1e-3) / 1e12
        ach_fwd = flops_fwd / (kern["fwd_ms"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "joint_rnnt_bwd (dominant)", "achieved": ach_bwd,
                "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach_bwd / pk["tf_sust"], "traffic": _ncu_traffic(),
                "peak_source": pk["src"] + " bf16 sustained",
                "step_frac": (6.0 * M * D * V / (ms_step * 1e-3) / 1e12) / pk["tf_sust"],
                "fwd": {"achieved": ach_fwd, "frac": ach_fwd / pk["tf_sust"], "ms": kern["fwd_ms"]},
                "bwd": {"achieved": ach_bwd, "frac": ach_bwd / pk["tf_sust"], "ms": kern["bwd_ms"]},
                "lattice": {"ms": kern["lat_ms"], "achieved_GBps": 24.0 * M / (kern["lat_ms"] * 1e-3) / 1e9,
                            "frac_hbm": 24.0 * M / (kern["lat_ms"] * 1e-3) / 1e9 / pk["hbm"]}}
        threads = os.cpu_count() or 1
        nb = 4
        cstep = cpu_reference_step(nb, threads)
        cstep()
        csec = min(cstep(), cstep())
        line = {"metric": METRIC, "value": gB / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": {"workload": "configs[1]: B=32/GPU,T=250,U=40,H=D=512,V=412 fused joint+rnnt_loss fwd/bwd",
                           "global_batch": gB, "parallelism": f"dp{world}", "l2": "flushed (256 MiB write) before each step",
                           "precision": args.precision, "cuda_graph": bool(graphed is not None)},
                "e2e": {"value": gB / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "h2d_GBps_measured": h2d_gbps, "h2d_ms_per_step": h2d / (h2d_gbps * 1e9) * 1e3},
                "gpu_launches": int(launches), "roofline": roof,
                "cpu_baseline": {"value": nb / csec, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": f"{nb} of {B} utterances per step, best of 2 after 1 warm-up, torch CPU + torchaudio"},
                "clocks": clk.summary(), "loss": float(loss.item()) * world}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def _ncu_traffic():
    """DRAM bytes (read + write) of the backward entry point's two tensor-core kernels per call, from the committed
    `ncu --set full` capture of this round (profiles/r1_v6_traffic.json); None when the file is absent."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_v6_traffic.json")
    try:
        k = json.load(open(path))["kernels"]
        return float(sum(k[n]["dram_read_bytes"] + k[n]["dram_write_bytes"] for n in ("joint_bwd2_kernel", "dw_gemm_rz_kernel")))
    except (OSError, KeyError, ValueError):
        return None


def time_kernels(C, joint, enc, pred, tgt, tl, ul, blank, precision, flush, reps=5):
    from ctcvr_b200._lib import call, ptr, query, stream
    prec = {"fp32": 0, "bf16": 1}[precision]
    with torch.no_grad():
        e, p = joint.project(enc.detach(), pred.detach())
        e, p = e.contiguous(), p.contiguous()
        w, b = joint.ffn_out.weight.detach().contiguous(), joint.ffn_out.bias.detach().contiguous()
    B, T, D = e.shape
    U1, V = p.shape[1], w.shape[0]
    dev = e.device
    lse = torch.empty(B, T, U1, device=dev)
    lpb, lpl, al, be = (torch.empty_like(lse) for _ in range(4))
    costs = torch.empty(B, device=dev)
    gc = torch.full((B,), 1.0 / B, device=dev)
    d_e, d_p, d_w, d_b = torch.empty_like(e), torch.empty_like(p), torch.empty_like(w), torch.empty_like(b)
    wsf = torch.empty(max(256, query("ctcvr_joint_rnnt_fwd_ws_bytes", B, T, U1, D, V, prec)), dtype=torch.uint8, device=dev)
    wsb = torch.empty(max(256, query("ctcvr_joint_rnnt_bwd_ws_bytes", B, T, U1, D, V, prec)), dtype=torch.uint8, device=dev)

    def fwd():
        call("ctcvr_joint_rnnt_fwd", ptr(e), ptr(p), ptr(w), ptr(b), ptr(tgt), ptr(tl), ptr(ul), ptr(lse), ptr(lpb),
             ptr(lpl), B, T, U1, D, V, blank, prec, ptr(wsf), wsf.numel(), stream())

    def lat():
        call("ctcvr_rnnt_lattice", ptr(lpb), ptr(lpl), ptr(tl), ptr(ul), ptr(al), ptr(be), ptr(costs), B, T, U1, stream())

    def bwd():
        call("ctcvr_joint_rnnt_bwd", ptr(e), ptr(p), ptr(w), ptr(b), ptr(tgt), ptr(tl), ptr(ul), ptr(lse), ptr(lpb), ptr(lpl), ptr(al),
             ptr(be), ptr(costs), ptr(gc), -1.0, ptr(d_e), ptr(d_p), ptr(d_w), ptr(d_b), B, T, U1, D, V, blank, prec,
             ptr(wsb), wsb.numel(), stream())

    out = {}
    for name, fn in (("fwd_ms", fwd), ("lat_ms", lat), ("bwd_ms", bwd)):
        fn()
        tot = 0.0
        for _ in range(reps):
            flush.zero_()
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            t.record()
            torch.cuda.synchronize()
            tot += s.elapsed_time(t)
        out[name] = tot / reps
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="eager step instead of the captured CUDA graph")
    ap.add_argument("--precision", default=os.environ.get("CTCVR_PRECISION", "bf16"), choices=["bf16", "fp32"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
