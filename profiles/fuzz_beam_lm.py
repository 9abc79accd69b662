"""Random-configuration check of the LM-fused beam kernel against oracle/lm.py + oracle/beam.py (restatements of kenlm /
pyctcdecode, unpinned): transcripts identical, ranking scores within 1e-4.
    timeout 900 python profiles/fuzz_beam_lm.py [cases] [seed] [--emu]
--emu runs the kernel on the CPU SIMT emulator (tests/emu) instead of the GPU."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import beam as ob, lm as olm
from silent_speech_decoder_b200 import synth
from silent_speech_decoder_b200.data.vocab import Vocab
from silent_speech_decoder_b200.decoding import ctc, lm as plm
from tests import lm_cases

EMU = "--emu" in sys.argv
args = [a for a in sys.argv[1:] if not a.startswith("--")]
cases = int(args[0]) if args else 30
rng = np.random.default_rng(int(args[1]) if len(args) > 1 else 0)
V = Vocab.default()
text, space = ctc._beam_label_strings(V)
keep = [i for i in range(V.size) if i not in {V.blank_id, V.pad_id}]
labels = [""] + [text[i] for i in keep]
if EMU:
    from tests.emu import harness
    lib = harness.library()
else:
    import torch
    from silent_speech_decoder_b200 import ops
tmp = tempfile.mkdtemp()
arpa = {}
for order in (1, 2, 3, 4, 5, 6):
    arpa[order] = os.path.join(tmp, f"w{order}.arpa")
    lm_cases.write_lm(arpa[order], order=order, seed=order)
pool = lm_cases.corpus(300, seed=77) + lm_cases.SENTENCES
bad = 0
for c in range(cases):
    order = int(rng.choice([1, 2, 3, 5, 6]))
    B = int(rng.choice([1, 3, 5]))
    T = int(rng.choice([1, 3, 30, 70, 120] if EMU else [1, 3, 30, 70, 120, 300]))
    width = int(rng.choice([1, 2, 5, 20, 50] if EMU else [1, 2, 5, 20, 50, 100]))
    prune = float(rng.choice([-10.0, -12.0, -5.0, -20.0]))
    bias = float(rng.choice([0.0, 0.0, 0.12, 0.2, -0.5]))
    alpha = float(rng.choice([0.0, 0.3, 0.6, 0.9, 2.0]))
    beta = float(rng.choice([0.0, 0.05, 0.5, 1.5, -0.5]))
    kind = str(rng.choice(["spelled", "spelled_confident", "flat", "peaked"]))
    if kind.startswith("spelled"):
        sents = [pool[int(rng.integers(len(pool)))][: max(1, T // 3)].strip() or "a" for _ in range(B)]
        lp, _ = lm_cases.spelled_log_probs(sents, T, seed=1000 + c, boost=3.0 if kind == "spelled" else 7.0) if T >= 3 else (synth.flat(B, T, c).numpy(), None)
    else:
        lp = (synth.flat(B, T, 100 + c) if kind == "flat" else synth.peaked(B, T, 100 + c)).numpy()
    ln = synth.ragged_lengths(B, T, 200 + c).tolist()
    if B > 1:
        ln[1] = 0
    L = olm.LanguageModel(olm.ArpaModel(arpa[order]), alpha=alpha, beta=beta)
    want = ob.BeamDecoder(V.tokens, V.blank_id, V.pad_id, width, prune, blank_bias=bias, lm=L).decode_with_lm_scores(lp, ln)
    tab = plm.build_tables(arpa[order], labels, alpha, beta)
    if EMU:
        (ids, score), = harness.beam_lm(lib, lp, ln, V.blank_id, V.pad_id, space, tab, [(width, prune, bias, -5.0)])
        got = ["".join(ob.UNK_GLYPH if i == V.unk_id else V.tokens[i] for i in s) for s in ids]
        got = [" ".join(g.split()) for g in got]
    else:
        tok, cnt, sc = ops.beam_lm_ids(torch.from_numpy(lp), torch.tensor(ln), V.blank_id, V.pad_id, space, tab, [(width, prune, bias, -5.0)])
        got = ctc._rows_to_strings(tok[0], cnt[0], text, normalise_space=True)
        score = sc[0].cpu().numpy()
    ok = all(g == w[0] for g, w in zip(got, want)) and all(abs(s - w[2]) <= 1e-4 for s, w in zip(score, want) if np.isfinite(w[2]))
    bad += not ok
    print(("ok  " if ok else "BAD ") + f"case {c}: order={order} B={B} T={T} width={width} prune={prune} bias={bias} alpha={alpha} beta={beta} {kind}", flush=True)
    if not ok:
        for g, w, s in zip(got, want, score):
            if g != w[0] or abs(s - w[2]) > 1e-4:
                print("   got ", repr(g[:70]), s, "\n   want", repr(w[0][:70]), w[2])
print(f"{cases - bad} of {cases} cases identical to the restatement")
sys.exit(1 if bad else 0)
