"""CPU restatement of the reference's error-rate arithmetic (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module; the product path (``silent_speech_decoder_b200.evaluation``) never does.

* ``levenshtein_counts`` follows ``_levenshtein_counts`` of the reference,
  ``src/evaluation/evaluate.py:61-83``: a table of ``(cost, ins, del, sub, hits)`` tuples, each cell
  the best of [insertion, deletion, diagonal] under ``key=(cost, -hits)`` (first wins on a tie).
  PINNED: ``tests/golden/edit_ref.json`` holds outputs of the reference's own
  ``compute_error_breakdown`` (its fallback branch, jiwer being absent here) on seeded inputs;
  ``tests/test_oracle_edit.py`` checks this restatement against them.
* ``error_breakdown`` follows ``compute_error_breakdown``'s fallback branch, ``evaluate.py:85-99``.
* ``wer`` / ``cer`` restate what ``jiwer.wer`` / ``jiwer.cer`` (``evaluate.py:34-38``) compute with
  their default transforms, from memory of jiwer 3.x (the package is not installed: PARITY UNPINNED
  for these two aggregates): WER = word edits / reference words after collapsing runs of spaces and
  stripping; CER = character edits / reference characters after stripping (inner spaces count).
"""

from __future__ import annotations

from typing import Dict, List, Sequence


def levenshtein_counts(ref_tokens: Sequence, hyp_tokens: Sequence) -> Dict[str, int]:
    n, m = len(ref_tokens), len(hyp_tokens)
    prev = [(j, j, 0, 0, 0) for j in range(m + 1)]  # row 0: j insertions   (evaluate.py:68-69)
    for i in range(1, n + 1):
        cur = [(i, 0, i, 0, 0)]  # column 0: i deletions                   (evaluate.py:66-67)
        for j in range(1, m + 1):
            a = cur[j - 1]
            ins = (a[0] + 1, a[1] + 1, a[2], a[3], a[4])
            b = prev[j]
            dele = (b[0] + 1, b[1], b[2] + 1, b[3], b[4])
            d = prev[j - 1]
            if ref_tokens[i - 1] == hyp_tokens[j - 1]:
                diag = (d[0], d[1], d[2], d[3], d[4] + 1)
            else:
                diag = (d[0] + 1, d[1], d[2], d[3] + 1, d[4])
            cur.append(min([ins, dele, diag], key=lambda t: (t[0], -t[4])))  # evaluate.py:81
        prev = cur
    _, i_c, d_c, s_c, h_c = prev[m]
    return {"insertions": i_c, "deletions": d_c, "substitutions": s_c, "hits": h_c}


def error_breakdown(refs: Sequence[str], hyps: Sequence[str]) -> Dict[str, float]:
    totals = {"insertions": 0, "deletions": 0, "substitutions": 0, "hits": 0}
    for ref, hyp in zip(refs, hyps):
        c = levenshtein_counts(ref.split(), hyp.split())
        for k in totals:
            totals[k] += c[k]
    total_words = max(1.0, float(totals["substitutions"] + totals["deletions"] + totals["hits"]))
    return {
        "substitutions": float(totals["substitutions"]),
        "deletions": float(totals["deletions"]),
        "insertions": float(totals["insertions"]),
        "hits": float(totals["hits"]),
        "substitution_rate": float(totals["substitutions"]) / total_words,
        "deletion_rate": float(totals["deletions"]) / total_words,
        "insertion_rate": float(totals["insertions"]) / total_words,
    }


def _rate(pairs: List) -> float:
    edits = total = 0
    for r, h in pairs:
        c = levenshtein_counts(r, h)
        edits += c["substitutions"] + c["deletions"] + c["insertions"]
        total += c["substitutions"] + c["deletions"] + c["hits"]
    return edits / total


def wer(refs: Sequence[str], hyps: Sequence[str]) -> float:
    if any(len(r.split()) == 0 for r in refs):
        raise ValueError("one or more references are empty strings")
    return _rate([(r.split(), h.split()) for r, h in zip(refs, hyps)])


def cer(refs: Sequence[str], hyps: Sequence[str]) -> float:
    if any(len(r.strip()) == 0 for r in refs):
        raise ValueError("one or more references are empty strings")
    return _rate([(list(r.strip()), list(h.strip())) for r, h in zip(refs, hyps)])
