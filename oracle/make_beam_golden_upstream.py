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

With ``kenlm`` installed too (``pip install kenlm``), the LM-fused cases (scope row f-4) are added: the same call
with ``kenlm_model_path=<an ARPA file written by synth.write_arpa>`` and the reference's ``alpha`` / ``beta``
(``ctc.py:67-72``); ``tests/test_lm_fusion.py::test_lm_restatement_matches_upstream`` and
``tests/test_gpu_lm.py::test_lm_beam_upstream`` consume them.

Only numpy, pyctcdecode (+ kenlm) and this repo's ``synth`` module are needed (no reference checkout, no GPU).
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


LM_CASES = [  # name, n-gram order, words, B, T, seed, beam_width, beam_prune_logp, blank_bias, alpha, beta
    ("lm5_w50", 5, 300, 8, 200, 6101, 50, -10.0, 0.0, 0.6, 0.0),      # build_char_lm.py's order, ctc.py's defaults
    ("lm3_w20_bias", 3, 300, 6, 150, 6102, 20, -12.0, 0.1, 0.9, 1.5),
    ("lm5_w100", 5, 1000, 4, 300, 6103, 100, -10.0, 0.0, 0.5, 0.05),
    ("lm2_w1", 2, 100, 4, 100, 6104, 1, -10.0, 0.0, 0.6, 0.0),
]


def lm_case_inputs(order: int, words: int, B: int, T: int, seed: int, arpa_path: str):
    """Writes the case's ARPA file to ``arpa_path`` and returns (log_probs, lengths); shared with the consuming tests."""
    from silent_speech_decoder_b200 import synth

    with open(os.path.join(REPO, "silent-speech-decoder_b200", "configs", "vocab.json")) as f:
        tokens = json.load(f)["tokens"]
    wl = synth.word_list(words, seed)
    synth.write_arpa(arpa_path, synth.word_corpus(wl, 10 * words, seed + 1), order)
    sents = synth.word_corpus(wl, B, seed + 2, lo=2, hi=6)
    sents[-1] = sents[-1] + " zzqx"   # an out-of-vocabulary word
    lp, ln = synth.spelled(sents, {t: i for i, t in enumerate(tokens)}, T, seed + 3)
    ln = [int(v) for v in ln]
    ln[1] = max(1, T - 13)
    return lp.numpy(), ln


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
    try:
        import kenlm  # noqa: F401
    except ImportError:
        print("kenlm not installed: LM cases skipped (LM parity stays unpinned)")
    else:
        import tempfile

        for name, order, words, B, T, seed, width, prune, bias, alpha, beta in LM_CASES:
            arpa = os.path.join(tempfile.mkdtemp(), f"{name}.arpa")
            lp, ln = lm_case_inputs(order, words, B, T, seed, arpa)
            dec = build_ctcdecoder(labels=labels, kenlm_model_path=arpa, alpha=alpha, beta=beta)   # ctc.py:67-72
            blank_logp = lp[:, :, blank] + (bias if bias != 0.0 else 0.0)
            blank_logp = np.logaddexp(blank_logp, lp[:, :, pad])
            merged = np.concatenate([blank_logp[:, :, None], lp[:, :, keep]], axis=-1)
            texts, logit, fused = [], [], []
            for i in range(B):
                seq = merged[i, : ln[i]]
                texts.append(dec.decode(seq, beam_width=width, beam_prune_logp=prune))
                beams = dec.decode_beams(seq, beam_width=width, beam_prune_logp=prune)
                assert beams[0][0] == texts[-1]
                logit.append(float(beams[0][-2]))
                fused.append(float(beams[0][-1]))
            pack[f"{name}.text"] = np.array(texts)
            pack[f"{name}.logit_score"] = np.array(logit, dtype=np.float64)
            pack[f"{name}.lm_score"] = np.array(fused, dtype=np.float64)
    import pyctcdecode

    pack["pyctcdecode_version"] = np.array(getattr(pyctcdecode, "__version__", "unknown"))
    pack["numpy_version"] = np.array(np.__version__)
    out = os.path.join(REPO, "tests", "golden", "beam_upstream.npz")
    np.savez_compressed(out, **pack)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
