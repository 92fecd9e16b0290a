"""Dev tool: one call of each non-joint hot-path kernel at its BASELINE.json size (CTC loss cfg5, lattice cfg2, greedy
cfg3, the three beams at beam 10 on shortened inputs) for `ncu --metrics gpu__time_duration.sum,dram__bytes_*`."""
import os, sys, types, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctcvr_b200 as C
torch.manual_seed(0)
V, BLANK, H = 412, 5, 256
# A4 / A10 / A9 at cfg5
B, T, U = 32, 500, 40
x = torch.randn(B, T, V, device="cuda", requires_grad=True)
ys = torch.randint(6, V, (B, U), device="cuda"); hl = torch.full((B,), T, device="cuda"); yl = torch.full((B,), U, device="cuda")
loss, lp = C.ctc_loss_from_logits(x, ys, hl, yl, BLANK, "sum"); loss.backward()
C.ctc_greedy_hyps(lp.detach(), hl, BLANK)
C.ctc_prefix_beam_search(torch.log_softmax(x.detach() * 2, -1), hl, 10, blank_id=BLANK)
# A5 at cfg3 (1000 utterances x 249 frames)
pred = C.RNNPredictor(V, H, H, 0.0, H, 1, dropout=0.0).cuda().eval(); joint = C.TransducerJoint(V, H, H, H).cuda().eval()
with torch.no_grad(): joint.ffn_out.bias[BLANK] += 1.0
m = types.SimpleNamespace(predictor=pred, joint=joint, blank=BLANK)
enc = torch.randn(1000, 249, H, device="cuda"); el = torch.full((1000,), 249, dtype=torch.int32, device="cuda")
C.basic_greedy_search(m, enc, el, n_steps=64)
# A7 / A8 on 64 frames
st = None
for s in range(0, 64, 16): hy, st = C.beam_chunk_online(m, enc[:1, s:s + 16], st, beam_size=10, n_steps=10)
ctc_logp = torch.log_softmax(enc[0, :64] @ (torch.randn(V, H, device="cuda") / 16).T, -1)
C.prefix_beam_search(m, enc[0, :64], ctc_logp, beam_size=10)
# CER (f4)
C.calculate_cer_batch([list(range(40))] * 256, [list(range(3, 43))] * 256)
torch.cuda.synchronize(); print("ok")
