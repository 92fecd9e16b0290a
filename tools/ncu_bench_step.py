"""Dev tool: the bench.py training step (public API) a few times, for an ncu launch list of EVERY kernel in a step."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctcvr_b200 as C
B, T, U, D, V, blank = 32, 250, 40, 512, 412, 5
dev = torch.device("cuda")
torch.manual_seed(1234)
joint = C.TransducerJoint(V, D, D, D).to(dev)
enc = torch.randn(B, T, D, device=dev, requires_grad=True)
pred = torch.randn(B, U + 1, D, device=dev, requires_grad=True)
tgt = torch.randint(6, V, (B, U), dtype=torch.int32, device=dev)
tl = torch.full((B,), T, dtype=torch.int32, device=dev); ul = torch.full((B,), U, dtype=torch.int32, device=dev)
def step():
    joint.zero_grad(set_to_none=True); enc.grad = pred.grad = None
    costs = joint.rnnt_loss_fused(enc, pred, tgt, tl, ul, blank, reduction="none", precision="bf16")
    loss = costs.sum() / B
    loss.backward()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("timed")
step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("ok")
