"""Random-configuration check of the beam kernel against oracle/beam.py (the CPU restatement of pyctcdecode):
transcripts identical, scores within 1e-4.   timeout 600 python profiles/fuzz_beam.py [cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import beam as ob
from silent_speech_decoder_b200 import ops, synth
from silent_speech_decoder_b200.data.vocab import Vocab
from silent_speech_decoder_b200.decoding import ctc

V = Vocab.default()
_, space = ctc._beam_label_strings(V)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for c in range(cases):
    B = int(rng.choice([1, 3, 6]))
    T = int(rng.choice([1, 2, 17, 60, 130, 260]))
    width = int(rng.choice([1, 2, 5, 20, 50, 100, 200]))
    prune = float(rng.choice([-10.0, -12.0, -5.0, -20.0]))
    bias = float(rng.choice([0.0, 0.0, 0.12, 0.2, -0.5, 2.5]))
    kind = rng.choice(["flat", "peaked", "mild"])
    lp = synth.flat(B, T, 100 + c) if kind == "flat" else synth.peaked(B, T, 100 + c, boost=8.0 if kind == "peaked" else 3.0)
    ln = synth.ragged_lengths(B, T, 200 + c)
    if B > 1:
        ln[1] = 0
    want = ob.BeamDecoder(V.tokens, V.blank_id, V.pad_id, width, prune, bias).decode_with_scores(lp.numpy(), ln.tolist())
    text = ctc.build_beam_decoder(V, None, width, beam_prune_logp=prune, blank_bias=bias)(lp, ln)
    _, _, score = ops.beam_ids(lp, ln, V.blank_id, V.pad_id, space, width, prune, bias)
    score = score.cpu().numpy()
    ok = all(t == w for t, (w, _) in zip(text, want)) and all(abs(s - ws) <= 1e-4 for s, (_, ws) in zip(score, want) if np.isfinite(ws))
    bad += not ok
    print(("ok  " if ok else "BAD ") + f"case {c}: B={B} T={T} width={width} prune={prune} bias={bias} {kind}", flush=True)
    if not ok:
        for t, (w, ws), s in zip(text, want, score):
            if t != w or abs(s - ws) > 1e-4:
                print("   got ", repr(t[:70]), s, "\n   want", repr(w[:70]), ws)
print(f"{cases - bad} of {cases} cases identical to the restatement")
sys.exit(1 if bad else 0)
