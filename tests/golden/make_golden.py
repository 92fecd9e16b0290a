"""Generate the golden fixtures in this directory by running the UNMODIFIED reference
(/root/reference, CentaureaHO/CTC-VR) on CPU.  Run here (the build container); the fixtures
travel to the GPU box, the reference does not.

    cd /root/repo && python tests/golden/make_golden.py

Import shims: torch 2.11 no longer re-exports `Union` from
torch.nn.modules.conv (wenet/squeezeformer/conv2d.py:17) and `whisper` is absent
(wenet/utils/common.py:24).  One behavioural patch, stated in oracle/transducer_oracle.py:
wenet/transducer/search/prefix_beam_search.py:137 calls log_add([a,b]) with a list although the
vendored log_add takes varargs (TypeError on the first prefix merge); the list form of upstream
wenet is patched into that module's namespace to obtain fixtures for the merge path.
"""
import math
import os
import sys
import types
import typing

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _shims():
    import torch.nn.modules.conv as C
    C.Union = typing.Union
    if not hasattr(C, "Optional"):
        C.Optional = typing.Optional
    w, wt = types.ModuleType("whisper"), types.ModuleType("whisper.tokenizer")
    wt.LANGUAGES = {}
    w.tokenizer = wt
    sys.modules["whisper"], sys.modules["whisper.tokenizer"] = w, wt
    sys.path.insert(0, REF)
    os.chdir(REF)


def sd_np(mod, prefix=""):
    return {prefix + k: v.detach().cpu().numpy() for k, v in mod.state_dict().items()}


class StubEncoder(torch.nn.Module):
    """Identity 'encoder': the hot path starts at encoder_out, so chunks ARE encoder frames."""
    def forward_chunk(self, xs, offset, required_cache_size, att_cache, cnn_cache):
        return xs, att_cache, cnn_cache

    def forward(self, xs, lens, *a, **k):
        return xs, torch.ones(xs.size(0), 1, xs.size(1), dtype=torch.bool)


def main():
    _shims()
    import torchaudio
    from model.component.joint import TransducerJoint
    from model.component.predictor import RNNPredictor
    from model.component.transducer import basic_greedy_search
    from model.rnnt_model import TransducerModel, CTC
    from model.online_rnnt_model import OnlineRNNTModel, OnlineCTC
    from wenet.transformer.search import ctc_prefix_beam_search
    from wenet.transducer.search import prefix_beam_search as pbs_mod
    from wenet.transducer.search.greedy_search import basic_greedy_search as wenet_greedy

    blank = 5

    # ---- 1. joint + rnnt_loss, small ragged batch (A1+A2+A3) -------------------------------
    torch.manual_seed(7)
    B, T, U, E, P, D, V = 4, 13, 6, 24, 20, 32, 29
    joint = TransducerJoint(V, E, P, D)
    enc = torch.randn(B, T, E, requires_grad=True)
    pred = torch.randn(B, U + 1, P, requires_grad=True)
    tgt = torch.randint(6, V, (B, U), dtype=torch.int32)
    tl = torch.tensor([13, 9, 11, 5], dtype=torch.int32)
    ul = torch.tensor([6, 3, 0, 4], dtype=torch.int32)
    fx = {}
    for clamp in (-1.0, 0.05):
        for p_ in list(joint.parameters()) + [enc, pred]:
            p_.grad = None
        logits = joint(enc, pred)
        logits.retain_grad()
        loss = torchaudio.functional.rnnt_loss(logits, tgt, tl, ul, blank=blank, reduction="mean", clamp=clamp)
        costs = torchaudio.functional.rnnt_loss(logits.detach(), tgt, tl, ul, blank=blank, reduction="none", clamp=clamp)
        loss.backward()
        tag = "c" if clamp > 0 else "n"
        fx.update({f"{tag}_loss": loss.item(), f"{tag}_costs": costs.numpy(), f"{tag}_dlogits": logits.grad.numpy(),
                   f"{tag}_d_enc": enc.grad.numpy().copy(), f"{tag}_d_pred": pred.grad.numpy().copy()})
        for k, v in joint.named_parameters():
            fx[f"{tag}_d_{k}"] = v.grad.numpy().copy()
    fx.update(sd_np(joint, "w_"))
    fx.update(enc=enc.detach().numpy(), pred=pred.detach().numpy(), tgt=tgt.numpy(), tl=tl.numpy(), ul=ul.numpy(),
              blank=blank, clamp=0.05, logits=joint(enc, pred).detach().numpy())
    np.savez_compressed(os.path.join(OUT, "rnnt_small.npz"), **fx)
    print("rnnt_small loss", fx["n_loss"], fx["c_loss"])

    # ---- 2. cfg1: TransducerModel on example1.pt[:2] (BASELINE.json configs[0]) -------------
    torch.manual_seed(0)
    ex1 = torch.load(os.path.join(REF, "example1.pt"))
    n = 2
    alen = ex1["audio_lens"][:n]
    tlen = ex1["text_lens"][:n]
    aud = ex1["audios"][:n, :int(alen.max())]
    txt = ex1["texts"][:n, :int(tlen.max())]
    model = TransducerModel(80, 256, 412, blank, ctc_weight=0.3).eval()
    model.ctc.dropout_rate = 0.0
    cap_e = {}
    def _hook_enc(m, i, o):
        o[0].retain_grad()
        cap_e.update(enc=o[0], mask=o[1])

    def _hook_pred(m, i, o):
        o.retain_grad()
        cap_e.update(pred=o, ys_in=i[0])

    h2 = model.encoder.register_forward_hook(_hook_enc)
    h3 = model.predictor.register_forward_hook(_hook_pred)
    _, loss, ld = model(aud, alen, txt, tlen)
    loss.backward()
    h2.remove()
    h3.remove()
    enc_o, pred_o = cap_e["enc"], cap_e["pred"]
    fx = dict(encoder_out=enc_o.detach().numpy(), predictor_out=pred_o.detach().numpy(),
              encoder_out_lens=cap_e["mask"].squeeze(1).sum(1).numpy(), texts=txt.numpy(), text_lens=tlen.numpy(),
              ys_in=cap_e["ys_in"].numpy(),
              loss=loss.item(), loss_rnnt=ld["loss_rnnt"].item(), loss_ctc=ld["loss_ctc"].item(),
              d_encoder_out=enc_o.grad.numpy(), d_predictor_out=pred_o.grad.numpy(), blank=blank)
    fx.update(sd_np(model.joint, "joint."))
    fx.update(sd_np(model.ctc, "ctc."))
    for k, v in list(model.joint.named_parameters()) + [("ctc_lo.weight", model.ctc.ctc_lo.weight), ("ctc_lo.bias", model.ctc.ctc_lo.bias)]:
        fx["d_" + k] = v.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "cfg1_example1.npz"), **fx)
    print("cfg1 loss", fx["loss"], fx["loss_rnnt"], fx["loss_ctc"], "T'", enc_o.shape, "U+1", pred_o.shape)

    # ---- 3. example2.pt known answers (3.ipynb:87,159 ; 3_v2.ipynb:150) + prefix beam (A9, A10)
    ex2 = torch.load(os.path.join(REF, "example2.pt"))
    pre, lens = ex2["pre"].detach(), ex2["lens"]
    res = ctc_prefix_beam_search(pre, lens, 10, blank_id=blank)
    fx = dict(pre=pre.numpy().astype(np.float32), lens=lens.numpy(), blank=blank, beam=10,
              frame_argmax=pre.argmax(2).numpy())
    for i, r in enumerate(res):
        fx[f"best_{i}"] = np.array(r.tokens, dtype=np.int64)
        fx[f"score_{i}"] = r.score
        fx[f"times_{i}"] = np.array(r.times, dtype=np.int64)
        fx[f"nbest_scores_{i}"] = np.array(r.nbest_scores)
        fx[f"nbest_len_{i}"] = np.array([len(x) for x in r.nbest], dtype=np.int64)
        fx[f"nbest_flat_{i}"] = np.array([t for x in r.nbest for t in x], dtype=np.int64)
        fx[f"nbest_times_flat_{i}"] = np.array([t for x in r.nbest_times for t in x], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "example2.npz"), **fx)
    print("example2 best0", res[0].tokens, res[0].score, "best1", res[1].tokens, res[1].score)

    # ---- 4. decoders on a small seeded predictor/joint (A5, A6, A6', A7, A8) -----------------
    torch.manual_seed(11)
    H, V = 48, 40
    om = OnlineRNNTModel(input_dim=H, hidden_dim=H, vocab_size=V, blank_id=blank, predictor_dropout=0.0,
                         ctc_dropout_rate=0.0).eval()
    om.encoder = StubEncoder()
    with torch.no_grad():
        # blank bias so decoding terminates, strong predictor influence so emissions are mixed
        # (random weights otherwise hit the n_steps cap on every frame)
        om.joint.ffn_out.bias[blank] += 5.0
        om.joint.enc_ffn.weight.mul_(2.0)
        om.joint.pred_ffn.weight.mul_(4.0)
        om.predictor.embed.weight.mul_(3.0)
        om.predictor.rnn.weight_ih_l0.mul_(3.0)
        om.predictor.projection.weight.mul_(3.0)
        om.joint.ffn_out.weight.mul_(4.0)
    enc = torch.randn(3, 37, H) * 1.5
    elens = torch.tensor([37, 22, 30])
    fx = dict(enc=enc.numpy(), elens=elens.numpy(), blank=blank)
    fx.update(sd_np(om.predictor, "predictor."))
    fx.update(sd_np(om.joint, "joint."))
    fx.update(sd_np(om.ctc_head, "ctc."))
    # A5: offline greedy on the local predictor class sharing the same weights
    lp = RNNPredictor(V, H, H, 0.0, H, 1, dropout=0.0).eval()
    lp.load_state_dict(om.predictor.state_dict())
    holder = types.SimpleNamespace(predictor=lp, joint=om.joint, blank=blank)
    with torch.no_grad():
        hyps = basic_greedy_search(holder, enc, elens, n_steps=64)
        for b, h in enumerate(hyps):
            fx[f"a5_hyp_{b}"] = np.array(h, dtype=np.int64)
        h4 = basic_greedy_search(holder, enc, elens, n_steps=2)
        for b, h in enumerate(h4):
            fx[f"a5cap2_hyp_{b}"] = np.array(h, dtype=np.int64)
        # A6': wenet greedy (batch 1)
        holder_w = types.SimpleNamespace(predictor=om.predictor, joint=om.joint, blank=blank)
        fx["a6w_hyp_0"] = np.array(wenet_greedy(holder_w, enc[:1], torch.tensor(37), n_steps=64)[0], dtype=np.int64)
        # A6: streaming greedy, chunks of 8 encoder frames, state carried
        om.reset_streaming_cache(torch.device("cpu"))
        toks, st, last = [], None, blank
        for s in range(0, 37, 8):
            c, _, _, st, last = om._decode_chunk_streaming_logic(enc[:1, s:s + 8], 0, 0, om.streaming_att_cache,
                                                                  om.streaming_cnn_cache, st, last)
            toks += c
        fx["a6_hyp_0"] = np.array(toks, dtype=np.int64)
        fx["a6_h"], fx["a6_c"], fx["a6_last"] = st[0].numpy(), st[1].numpy(), last
        # A7: online beam, beam 4 and 10, chunked
        for beam in (4, 10):
            hy = None
            for s in range(0, 37, 8):
                hy, _, _ = om._decode_chunk_beam_search(enc[:1, s:s + 8], 0, 0, om.streaming_att_cache,
                                                        om.streaming_cnn_cache, hy, beam_size=beam)
            fx[f"a7_b{beam}_n"] = len(hy)
            for i, h in enumerate(hy):
                fx[f"a7_b{beam}_tok_{i}"] = np.array(h.tokens, dtype=np.int64)
                fx[f"a7_b{beam}_lp_{i}"] = h.log_prob
        # A8: wenet prefix beam with CTC fusion (list-form log_add patched in, see module docstring)
        def log_add_list(xs):
            if all(a == -float("inf") for a in xs):
                return -float("inf")
            m = max(xs)
            return m + math.log(sum(math.exp(a - m) for a in xs))
        pbs_mod.log_add = log_add_list
        searcher = pbs_mod.PrefixBeamSearch(om.encoder, om.predictor, om.joint, om.ctc_head, blank)
        for beam in (5, 10):
            seqs, _ = searcher.prefix_beam_search(enc[:1], torch.tensor([37]), beam_size=beam)
            fx[f"a8_b{beam}_n"] = len(seqs)
            for i, s in enumerate(seqs):
                fx[f"a8_b{beam}_tok_{i}"] = np.array(s.hyp, dtype=np.int64)
                fx[f"a8_b{beam}_sc_{i}"] = s.score
    np.savez_compressed(os.path.join(OUT, "decode_small.npz"), **fx)
    print("A5", [len(h) for h in hyps], "A6", len(toks), "A7 best", fx["a7_b4_tok_0"], "A8 best", fx["a8_b5_tok_0"])

    # ---- 5. CTC heads (A4) ------------------------------------------------------------------
    torch.manual_seed(3)
    B, T, Hc, V, U = 3, 21, 16, 17, 5
    hs = torch.randn(B, T, Hc, requires_grad=True)
    hl = torch.tensor([21, 15, 4])
    ys = torch.randint(0, V, (B, U))
    ys[ys == blank] = 7
    ys[0, 1] = ys[0, 0]                      # repeated label
    yl = torch.tensor([5, 3, 5])             # last one infeasible (T=4 < U=5) -> zero_infinity
    fx = dict(hs=hs.detach().numpy(), hl=hl.numpy(), ys=ys.numpy(), yl=yl.numpy(), blank=blank)
    for name, mod in (("off", CTC(V, Hc, 0.0, True, blank)), ("on", OnlineCTC(V, Hc, 0.0, blank))):
        hs.grad = None
        loss, ys_hat = mod(hs, hl, ys, yl)
        loss.backward()
        fx.update({f"{name}_loss": loss.item(), f"{name}_ys_hat": ys_hat.detach().numpy(),
                   f"{name}_d_hs": hs.grad.numpy().copy(), f"{name}_w": mod.ctc_lo.weight.detach().numpy(),
                   f"{name}_b": mod.ctc_lo.bias.detach().numpy(),
                   f"{name}_d_w": mod.ctc_lo.weight.grad.numpy(), f"{name}_d_b": mod.ctc_lo.bias.grad.numpy()})
    np.savez_compressed(os.path.join(OUT, "ctc_small.npz"), **fx)
    print("ctc", fx["off_loss"], fx["on_loss"])


CFG_BLANK_BIAS, CFG_SCALES = 4.0, {"ffn_out.weight": 0.25}


def decode_cfg_golden():
    """decode_cfg.npz: the reference decoders at BASELINE.json's config sizes - cfg3 (H=256, V=412, T'=249, chunk 16:
    A5 offline greedy, A6 streaming greedy) and cfg5 (beam 10, T'=500: A7 online beam, A8 wenet prefix beam with CTC
    fusion; A9 ctc_prefix_beam_search on 32 x 500 x 412).  Weights and inputs are the exact integer-hash tensors of
    synth.py (loaded into the unmodified reference modules here, rebuilt in the GPU tests), so the fixture holds only the
    expected hypotheses."""
    _shims()
    sys.path.insert(0, OUT)
    import synth
    from model.component.predictor import RNNPredictor
    from model.component.transducer import basic_greedy_search
    from model.online_rnnt_model import OnlineRNNTModel
    from wenet.transformer.search import ctc_prefix_beam_search
    from wenet.transducer.search import prefix_beam_search as pbs_mod
    H, V, blank = 256, 412, 5
    om = OnlineRNNTModel(input_dim=H, hidden_dim=H, vocab_size=V, blank_id=blank, predictor_dropout=0.0,
                         ctc_dropout_rate=0.0).eval()
    om.encoder = StubEncoder()
    for mod, tag in ((om.predictor, "cfg/predictor"), (om.joint, "cfg/joint"), (om.ctc_head, "cfg/ctc")):
        st = synth.decoder_state({k: tuple(v.shape) for k, v in mod.state_dict().items()}, tag, blank,
                                 scales=CFG_SCALES, blank_bias=CFG_BLANK_BIAS)
        mod.load_state_dict({k: torch.from_numpy(v) for k, v in st.items()})
    enc = torch.from_numpy(synth.synth((2, 500, H), "cfg/enc", 2.0))
    fx = dict(H=H, V=V, blank=blank, blank_bias=CFG_BLANK_BIAS, ffn_out_scale=CFG_SCALES["ffn_out.weight"])
    lp = RNNPredictor(V, H, H, 0.0, H, 1, dropout=0.0).eval()
    lp.load_state_dict(om.predictor.state_dict())
    holder = types.SimpleNamespace(predictor=lp, joint=om.joint, blank=blank)
    with torch.no_grad():
        # cfg3 / A5: offline greedy, two utterances of T' = 249 and 180
        hyps = basic_greedy_search(holder, enc[:, :249], torch.tensor([249, 180]), n_steps=64)
        for b, h in enumerate(hyps):
            fx[f"a5_hyp_{b}"] = np.array(h, dtype=np.int64)
        # cfg3 / A6: streaming greedy over T' = 249 in chunks of 16 encoder frames, predictor state carried
        om.reset_streaming_cache(torch.device("cpu"))
        toks, st, last = [], None, blank
        for s in range(0, 249, 16):
            c, _, _, st, last = om._decode_chunk_streaming_logic(enc[1:2, s:min(s + 16, 249)], 0, 0, om.streaming_att_cache,
                                                                  om.streaming_cnn_cache, st, last)
            toks += c
        fx["a6_hyp"], fx["a6_last"] = np.array(toks, dtype=np.int64), last
        # cfg5 / A7: online beam search, beam 10, T' = 500 in chunks of 16
        hy = None
        for s in range(0, 500, 16):
            hy, _, _ = om._decode_chunk_beam_search(enc[:1, s:s + 16], 0, 0, om.streaming_att_cache, om.streaming_cnn_cache,
                                                    hy, beam_size=10)
        fx["a7_n"] = len(hy)
        for i, h in enumerate(hy):
            fx[f"a7_tok_{i}"] = np.array(h.tokens, dtype=np.int64)
            fx[f"a7_lp_{i}"] = h.log_prob
        # cfg5 / A8: wenet prefix beam with CTC fusion, beam 10, T' = 500 (list-form log_add patched in, module docstring)
        def log_add_list(xs):
            if all(a == -float("inf") for a in xs):
                return -float("inf")
            m = max(xs)
            return m + math.log(sum(math.exp(a - m) for a in xs))
        pbs_mod.log_add = log_add_list
        searcher = pbs_mod.PrefixBeamSearch(om.encoder, om.predictor, om.joint, om.ctc_head, blank)
        seqs, _ = searcher.prefix_beam_search(enc[1:2], torch.tensor([500]), beam_size=10)
        fx["a8_n"] = len(seqs)
        for i, s in enumerate(seqs):
            fx[f"a8_tok_{i}"] = np.array(s.hyp, dtype=np.int64)
            fx[f"a8_sc_{i}"] = s.score
    # cfg5 / A9: CTC prefix beam search on 32 x 500 x 412 scores, ragged lengths
    B, T = 32, 500
    lens = np.array([T - 13 * (b % 7) if b % 5 else T for b in range(B)], dtype=np.int64)
    res = ctc_prefix_beam_search(torch.from_numpy(synth.ctc_logp(B, T, V, blank)), torch.from_numpy(lens), 10, blank_id=blank)
    fx["a9_lens"] = lens
    for i, r in enumerate(res):
        fx[f"a9_nbest_len_{i}"] = np.array([len(x) for x in r.nbest], dtype=np.int64)
        fx[f"a9_nbest_flat_{i}"] = np.array([t for x in r.nbest for t in x], dtype=np.int64)
        fx[f"a9_nbest_scores_{i}"] = np.array(r.nbest_scores)
        fx[f"a9_times_{i}"] = np.array(r.times, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "decode_cfg.npz"), **fx)
    print("cfg A5", [len(h) for h in hyps], "A6", len(toks), "A7", len(hy), len(hy[0].tokens), "A8", len(seqs), len(seqs[0].hyp),
          "A9 best lens", [int(fx[f"a9_nbest_len_{i}"][0]) for i in range(4)])


def cer_golden():
    """Section 8f row 4: calculate_cer of the reference (rnnt_eval.py:11-56).  rnnt_eval.py cannot be imported here (it
    pulls data.dataloader -> librosa), so the function is taken out of the file by its AST node and executed as is."""
    import ast
    src = open(os.path.join(REF, "rnnt_eval.py")).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "calculate_cer")
    ns = {}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "rnnt_eval.py", "exec"), ns)
    calculate_cer = ns["calculate_cer"]
    rng = np.random.default_rng(17)
    hyps, refs = [], []
    for i in range(64):
        n = int(rng.integers(0, 40))
        ref = rng.integers(6, 30, n).tolist()          # small alphabet: many ties in the backtrace
        hyp = list(ref)
        for _ in range(int(rng.integers(0, 12))):       # random edits
            op = int(rng.integers(0, 3))
            pos = int(rng.integers(0, len(hyp) + 1))
            if op == 0 and hyp:
                hyp[min(pos, len(hyp) - 1)] = int(rng.integers(6, 30))
            elif op == 1 and hyp:
                del hyp[min(pos, len(hyp) - 1)]
            else:
                hyp.insert(pos, int(rng.integers(6, 30)))
        if i % 16 == 0:
            hyp = []
        hyps.append(hyp)
        refs.append(ref)
    fx = dict(n=len(hyps))
    for i, (h, r) in enumerate(zip(hyps, refs)):
        cer, S, D, I, N = calculate_cer(h, r)
        fx[f"hyp_{i}"] = np.array(h, dtype=np.int64)
        fx[f"ref_{i}"] = np.array(r, dtype=np.int64)
        fx[f"res_{i}"] = np.array([S, D, I, N], dtype=np.int64)
        fx[f"cer_{i}"] = cer
    np.savez_compressed(os.path.join(OUT, "cer_small.npz"), **fx)
    print("cer", sum(int(fx[f"res_{i}"][:3].sum()) for i in range(len(hyps))), "edits over", len(hyps), "pairs")


def predictor_golden():
    """predictor_small.npz: the reference RNNPredictor (model/component/predictor.py:11-63)
    forward + backward on synth.py weights - one layer at V=412, H=128 (the models' layout at half width) and a
    two-layer H=32 case for the layer stacking.  The loss is sum(out * r) with a synthetic cotangent r, so the fixture
    holds out and the gradient of every parameter."""
    _shims()
    sys.path.insert(0, OUT)
    import synth
    from model.component.predictor import RNNPredictor
    fx = {}
    for tag, (V, H, L, B, U1) in {"p1": (412, 128, 1, 3, 9), "p2": (20, 32, 2, 5, 7)}.items():
        pr = RNNPredictor(V, H, H, 0.0, H, L, dropout=0.0).eval()
        shapes = {k: tuple(v.shape) for k, v in pr.state_dict().items()}
        st = synth.predictor_state(shapes, f"pred/{tag}")
        pr.load_state_dict({k: torch.from_numpy(v) for k, v in st.items()})
        ys, r = synth.predictor_case(tag, V, H, B, U1)
        out = pr(torch.from_numpy(ys))
        (out * torch.from_numpy(r)).sum().backward()
        fx.update({f"{tag}_dims": np.array([V, H, L, B, U1]), f"{tag}_out": out.detach().numpy()})
        for k, v in pr.named_parameters():
            fx[f"{tag}_d_{k}"] = v.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "predictor_small.npz"), **fx)
    print("predictor", {k: float(np.abs(v).max()) for k, v in fx.items() if k.endswith("_out")})


if __name__ == "__main__":
    if "--cer-only" in sys.argv:
        cer_golden()
    elif "--decode-cfg-only" in sys.argv:
        decode_cfg_golden()
    elif "--predictor-only" in sys.argv:
        predictor_golden()
    else:
        main()
        cer_golden()
        decode_cfg_golden()
        predictor_golden()
