"""The reference's own data flow for the fused seam on the GPU (the kernel bar of SURVEY.md §2b / §8d):
joint.py:48-69 (cuBLAS GEMMs + elementwise add/tanh) -> torchaudio.functional.rnnt_loss (its sm_100 SIMT kernels)
-> autograd backward, fp32 and under bf16 autocast; plus ATen's CTC loss at cfg5.  Library ops only: nothing of
libctcvr.so and nothing of oracle/ runs here.

    python tools/gpu_reference.py [--steps 5]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def joint_rnnt_reference_gpu(B=32, T=250, U=40, D=512, V=412, blank=5, steps=5, autocast=False, seed=1234):
    import torchaudio
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(seed)
    enc = torch.randn(B, T, D, generator=g).to(dev).requires_grad_(True)
    pred = torch.randn(B, U + 1, D, generator=g).to(dev).requires_grad_(True)
    tgt = torch.randint(6, V, (B, U), generator=g, dtype=torch.int32).to(dev)
    tl = torch.full((B,), T, dtype=torch.int32, device=dev)
    ul = torch.full((B,), U, dtype=torch.int32, device=dev)
    torch.manual_seed(seed)
    enc_ffn, pred_ffn, ffn_out = (torch.nn.Linear(D, D).to(dev), torch.nn.Linear(D, D).to(dev),
                                  torch.nn.Linear(D, V).to(dev))
    params = list(enc_ffn.parameters()) + list(pred_ffn.parameters()) + list(ffn_out.parameters())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        for p in params:
            p.grad = None
        enc.grad = pred.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            e, p = enc_ffn(enc), pred_ffn(pred)
            z = torch.tanh(e.unsqueeze(2) + p.unsqueeze(1))        # joint.py:57-67
            logits = ffn_out(z)                                     # joint.py:68
        logits = logits.float()                                     # transducer.py:174-178 (rnnt_loss rejects bf16)
        loss = torchaudio.functional.rnnt_loss(logits, tgt, tl, ul, blank=blank, reduction="mean")
        loss.backward()
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    ms = []
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        loss = step()
        t.record()
        torch.cuda.synchronize()
        ms.append(s.elapsed_time(t))
    ms.sort()
    med = ms[len(ms) // 2]
    return {"ms_per_step": med, "utt_per_s": B / (med * 1e-3), "loss": float(loss), "steps": steps,
            "autocast_bf16": autocast, "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}


def ctc_reference_gpu(B=32, T=500, V=412, U=40, blank=5, steps=5, seed=1234):
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, T, V, generator=g).to(dev).requires_grad_(True)
    tgt = torch.randint(6, V, (B, U), generator=g).to(dev)
    il = torch.full((B,), T, dtype=torch.int64, device=dev)
    tl = torch.full((B,), U, dtype=torch.int64, device=dev)
    crit = torch.nn.CTCLoss(blank=blank, reduction="sum", zero_infinity=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        logits.grad = None
        lp = torch.log_softmax(logits, -1).transpose(0, 1)          # model/rnnt_model.py:55-58
        loss = crit(lp, tgt, il, tl) / B
        loss.backward()
        return loss

    for _ in range(2):
        step()
    ms = []
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        loss = step()
        t.record()
        torch.cuda.synchronize()
        ms.append(s.elapsed_time(t))
    ms.sort()
    med = ms[len(ms) // 2]
    return {"ms_per_step": med, "utt_per_s": B / (med * 1e-3), "loss": float(loss), "steps": steps}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    out = {"joint_rnnt_fp32": joint_rnnt_reference_gpu(steps=a.steps, autocast=False),
           "joint_rnnt_autocast_bf16": joint_rnnt_reference_gpu(steps=a.steps, autocast=True),
           "ctc_loss_aten": ctc_reference_gpu(steps=a.steps)}
    print(json.dumps(out))
