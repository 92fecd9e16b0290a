"""Dev tool: the batched beam decoders (A7b / A8b: one CTA per utterance, 296 utterances x 64 frames, beam 10) for
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum`."""
import os, sys, types, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctcvr_b200 as C
torch.manual_seed(0)
V, BLANK, H, S, T = 412, 5, 256, 296, 64
pred = C.RNNPredictor(V, H, H, 0.0, H, 1, dropout=0.0).cuda().eval(); joint = C.TransducerJoint(V, H, H, H).cuda().eval()
with torch.no_grad(): joint.ffn_out.bias[BLANK] += 1.0
m = types.SimpleNamespace(predictor=pred, joint=joint, blank=BLANK)
enc = torch.randn(S, T, H, device="cuda"); el = torch.full((S,), T, dtype=torch.int32, device="cuda")
hy = C.beam_search_batch(m, enc, el, beam_size=10, n_steps=10)
ctc = torch.log_softmax(enc @ (torch.randn(V, H, device="cuda") / 16).T, -1)
pb = C.prefix_beam_search_batch(m, enc, el, ctc, beam_size=10)
# one stream of the same, for the per-kernel comparison
C.beam_search_batch(m, enc[:1], el[:1], beam_size=10, n_steps=10)
torch.cuda.synchronize(); print("ok", sum(len(h[0].tokens) for h in hy) / S, len(pb))
