"""CPU tier: oracle/edit.py against the outputs of the reference's own compute_error_breakdown
(tests/golden/edit_ref.json, made by oracle/make_golden.py with jiwer stubbed out so that the
reference runs its in-tree Levenshtein branch, evaluate.py:61-99)."""
import json
import os

import pytest

from oracle import edit


@pytest.fixture(scope="module")
def golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, "edit_ref.json")))


def _vec(c):
    return [c["substitutions"], c["deletions"], c["insertions"], c["hits"]]


def test_counts_match_reference(golden):
    for row in golden["pairs"]:
        assert _vec(edit.levenshtein_counts(row["ref"].split(), row["hyp"].split())) == row["word"], row
        r, h = row["ref"].strip().replace(" ", "_"), row["hyp"].strip().replace(" ", "_")
        assert _vec(edit.levenshtein_counts(list(r), list(h))) == row["char"], row


def test_corpus_breakdown_matches_reference(golden):
    refs = [r["ref"] for r in golden["pairs"]]
    hyps = [r["hyp"] for r in golden["pairs"]]
    assert edit.error_breakdown(refs, hyps) == golden["corpus_breakdown"]


def test_identities_and_rates():
    # S, D, I follow from (cost, hits, n, m): n = S+D+H, m = S+I+H
    c = edit.levenshtein_counts("a b c d".split(), "a x c".split())
    assert c["substitutions"] + c["deletions"] + c["hits"] == 4
    assert c["substitutions"] + c["insertions"] + c["hits"] == 3
    assert edit.wer(["a b c d"], ["a x c"]) == pytest.approx(2 / 4)
    assert edit.cer(["abcd"], ["abd"]) == pytest.approx(1 / 4)
    assert edit.cer([" ab cd "], ["ab cd"]) == 0.0  # stripped; the inner space is a character
    with pytest.raises(ValueError):
        edit.wer(["  "], ["a"])
