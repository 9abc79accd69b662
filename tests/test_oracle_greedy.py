"""Pins oracle/greedy.py against the reference's _greedy_decode / build_greedy_decoder outputs
(tests/golden/greedy_ref.npz) and its known-answer test (reference tests/test_evaluation.py:6-14)."""
import os

import numpy as np
import pytest

from oracle import greedy as og
from silent_speech_decoder_b200.data.vocab import Vocab


def _unragged(arr):
    return [[int(x) for x in row if x >= 0] for row in arr]


def test_known_answer_vector(golden_dir):
    z = np.load(os.path.join(golden_dir, "greedy_ref.npz"))
    assert og.greedy_ids(z["kat.lp"], z["kat.len"], int(z["kat.blank"])) == [[1, 2], [1, 2]]
    assert og.greedy_ids_loop(z["kat.lp"], z["kat.len"], int(z["kat.blank"])) == [[1, 2], [1, 2]]


@pytest.mark.parametrize("name", ["flat", "peaked", "peaked_bias", "flat_negbias"])
def test_matches_reference_outputs(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "greedy_ref.npz"))
    v = Vocab.default()
    lp, ln, bias = z[f"{name}.lp"], z[f"{name}.len"], float(z[f"{name}.bias"])
    want = _unragged(z[f"{name}.ids"])
    assert og.greedy_ids(lp, ln, v.blank_id, bias) == want
    assert og.greedy_ids_loop(lp, ln, v.blank_id, bias) == want
    assert og.greedy_text(lp, ln, v.tokens, v.blank_id, v.pad_id, bias) == list(z[f"{name}.text"])
    # Vocab mirror renders ids exactly like the reference's Vocab.decode
    assert [v.decode(s) for s in want] == list(z[f"{name}.text"])


def test_cfg1(golden_dir):
    z = np.load(os.path.join(golden_dir, "cfg1_ref.npz"))
    assert og.greedy_ids(z["lp"], z["il"], 1) == _unragged(z["greedy_ids"])
