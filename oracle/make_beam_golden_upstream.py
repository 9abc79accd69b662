"""Pin the beam oracle to the REAL library: run where ``pyctcdecode`` is installed (it is not in the build image).

    pip install pyctcdecode          # on any machine with network; numpy<2 as the library requires
    python oracle/make_beam_golden_upstream.py

Calls ``pyctcdecode.build_ctcdecoder(...).decode(...)`` exactly as the reference's wrapper does
(``src/decoding/ctc.py:61-97``: pad->blank ``logaddexp`` merge, label permutation with blank first,
``kenlm_model_path=None``, one ``decode(lp[:len], beam_width, beam_prune_logp)`` call per utterance; every other
argument at the library's default) on seeded inputs that the tests regenerate, and writes
``tests/golden/beam_upstream.npz`` with the transcripts and, through ``decode_beams``, the top beam's ``logit_score``.
``tests/test_oracle_beam.py::test_restatement_matches_upstream`` and ``tests/test_gpu_decoders.py::test_beam_upstream``
consume the file when it exists; until someone commits it, beam parity stays "unpinned" and every report says so.

Only numpy, pyctcdecode and this repo's ``synth`` module are needed (no reference checkout, no GPU).
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

CASES = [  # name, generator, B, T, seed, beam_width, beam_prune_logp, blank_bias
    ("flat50", "flat", 6, 120, 5101, 50, -10.0, 0.0),
    ("peaked50", "peaked", 6, 200, 5102, 50, -10.0, 0.0),
    ("flat20_bias", "flat", 4, 90, 5103, 20, -10.0, 0.2),
    ("peaked100_p12", "peaked", 4, 160, 5104, 100, -12.0, 0.05),
    ("flat1", "flat", 3, 80, 5105, 1, -10.0, 0.0),
    ("flat200", "flat", 2, 100, 5106, 200, -10.0, 0.0),
    ("peaked50_long", "peaked", 2, 1500, 5107, 50, -10.0, 0.0),
]


def case_inputs(gen: str, B: int, T: int, seed: int):
    """Seeded inputs and ragged lengths; shared with the consuming tests."""
    import torch

    from silent_speech_decoder_b200 import synth

    x = getattr(synth, gen)(B, T, seed)
    ln = synth.ragged_lengths(B, T, seed + 1)
    ln[0] = T
    return x.numpy(), [int(v) for v in ln]


def main() -> None:
    from pyctcdecode import build_ctcdecoder  # the library the reference imports lazily at ctc.py:56-59

    with open(os.path.join(REPO, "silent-speech-decoder_b200", "configs", "vocab.json")) as f:
        vj = json.load(f)
    tokens, blank, pad = vj["tokens"], int(vj["blank_id"]), int(vj["pad_id"])
    keep = [i for i in range(len(tokens)) if i not in {blank, pad}]       # ctc.py:64
    labels = [""] + [tokens[i] for i in keep]                              # ctc.py:65
    decoder = build_ctcdecoder(labels=labels, kenlm_model_path=None, alpha=0.6, beta=0.0)   # ctc.py:67-72

    pack = {}
    for name, gen, B, T, seed, width, prune, bias in CASES:
        lp, ln = case_inputs(gen, B, T, seed)
        blank_logp = lp[:, :, blank] + (bias if bias != 0.0 else 0.0)      # ctc.py:76-78
        blank_logp = np.logaddexp(blank_logp, lp[:, :, pad])               # ctc.py:79-81
        merged = np.concatenate([blank_logp[:, :, None], lp[:, :, keep]], axis=-1)   # ctc.py:84-87
        texts, scores = [], []
        for i in range(B):                                                 # ctc.py:89-96
            seq = merged[i, : ln[i]]
            texts.append(decoder.decode(seq, beam_width=width, beam_prune_logp=prune))
            beams = decoder.decode_beams(seq, beam_width=width, beam_prune_logp=prune)
            assert beams[0][0] == texts[-1]
            scores.append(float(beams[0][-2]))                             # (text, lm_state, frames, logit_score, lm_score)
        pack[f"{name}.text"] = np.array(texts)
        pack[f"{name}.score"] = np.array(scores, dtype=np.float64)
        pack[f"{name}.cfg"] = np.array([width, prune, bias], dtype=np.float64)
    import pyctcdecode

    pack["pyctcdecode_version"] = np.array(getattr(pyctcdecode, "__version__", "unknown"))
    pack["numpy_version"] = np.array(np.__version__)
    out = os.path.join(REPO, "tests", "golden", "beam_upstream.npz")
    np.savez_compressed(out, **pack)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
