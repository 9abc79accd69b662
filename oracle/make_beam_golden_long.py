"""Config-4 beam fixtures: T=4000, widths 1/10/50/200, peaked and flat posteriors (test infrastructure only).

    python oracle/make_beam_golden_long.py        # minutes of CPU: the restatement at width 200 on flat rows is slow

Inputs are NOT stored: tests regenerate them from the seeds with ``silent_speech_decoder_b200.synth`` (bit-identical,
CPU generators).  Outputs come from ``oracle/beam.py`` -- the restatement of pyctcdecode, PARITY UNPINNED like every
beam fixture here -- and freeze it at the long-utterance shape so the GPU parity test does not have to spend minutes
in the Python search.  ``tests/golden/beam_t4000.npz``: per case ``<kind>_w<width>.text`` / ``.score`` for
``N_UTT`` utterances.
"""

from __future__ import annotations

import multiprocessing as mp
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import beam as ob  # noqa: E402
from silent_speech_decoder_b200 import synth  # noqa: E402
from silent_speech_decoder_b200.data.vocab import Vocab  # noqa: E402

T, SEED, N_UTT = 4000, 1004, 2
WIDTHS = (1, 10, 50, 200)
LENGTHS = (4000, 3777)   # one full-length row, one ragged


def inputs(kind: str):
    x = synth.flat(N_UTT, T, SEED) if kind == "flat" else synth.peaked(N_UTT, T, SEED + 1)
    return x.numpy()


def _one(job):
    kind, w, i = job
    v = Vocab.default()
    dec = ob.BeamDecoder(v.tokens, v.blank_id, v.pad_id, w, -10.0)
    text, score = dec.decode_with_scores(inputs(kind)[i:i + 1], [LENGTHS[i]])[0]
    return kind, w, i, text, score


def main() -> None:
    jobs = [(k, w, i) for k in ("flat", "peaked") for w in WIDTHS[::-1] for i in range(N_UTT)]
    with mp.get_context("fork").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        res = pool.map(_one, jobs, chunksize=1)
    pack = {}
    for kind in ("flat", "peaked"):
        for w in WIDTHS:
            rows = sorted((r for r in res if r[0] == kind and r[1] == w), key=lambda r: r[2])
            pack[f"{kind}_w{w}.text"] = np.array([r[3] for r in rows])
            pack[f"{kind}_w{w}.score"] = np.array([r[4] for r in rows], dtype=np.float64)
    pack["lengths"] = np.array(LENGTHS, dtype=np.int64)
    out = os.path.join(REPO, "tests", "golden", "beam_t4000.npz")
    np.savez_compressed(out, **pack)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
