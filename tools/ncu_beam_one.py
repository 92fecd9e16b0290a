"""Dev tool: one stream of the online beam decoder (A7, beam 10, 64 frames) for an `ncu --set full --import-source on` capture."""
import os, sys, types, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctcvr_b200 as C
torch.manual_seed(0)
V, BLANK, H = 412, 5, 256
pred = C.RNNPredictor(V, H, H, 0.0, H, 1, dropout=0.0).cuda().eval(); joint = C.TransducerJoint(V, H, H, H).cuda().eval()
with torch.no_grad(): joint.ffn_out.bias[BLANK] += 1.0
m = types.SimpleNamespace(predictor=pred, joint=joint, blank=BLANK)
enc = torch.randn(1, 64, H, device="cuda"); el = torch.full((1,), 64, dtype=torch.int32, device="cuda")
hy = C.beam_search_batch(m, enc, el, beam_size=10, n_steps=10)
torch.cuda.synchronize(); print("ok", len(hy[0]))
