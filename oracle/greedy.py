"""Greedy CTC decode oracle (test infrastructure only -- see oracle/__init__.py).

Restates reference ``src/decoding/ctc.py:16-44`` (``_greedy_decode`` and the
``build_greedy_decoder`` closure) and ``src/data/vocab.py:44-53`` (``Vocab.decode``).
Points the reference fixes and this keeps:
  * the bias is added to the blank column in fp32 before the argmax (ctc.py:18-20)
  * argmax ties go to the lowest index (torch.argmax on CPU)
  * a frame emits iff its argmax is not blank and differs from the previous
    frame's argmax; a blank resets "previous" (ctc.py:26-34)
  * pad (id 0) is an ordinary token for the collapse; it is dropped only when
    ids are turned into a string (vocab.py:48-49)
"""

from __future__ import annotations

from typing import List, Sequence

import numpy as np


def greedy_ids(log_probs: np.ndarray, lengths: Sequence[int], blank_id: int, blank_bias: float = 0.0) -> List[List[int]]:
    lp = np.asarray(log_probs, dtype=np.float32)
    if blank_bias != 0.0:
        lp = lp.copy()
        lp[:, :, blank_id] = lp[:, :, blank_id] + np.float32(blank_bias)
    best = lp.argmax(axis=-1)  # first maximum, like torch.argmax on CPU
    out: List[List[int]] = []
    for row, n in zip(best, lengths):
        n = int(n)
        a = row[:n]
        if n == 0:
            out.append([])
            continue
        prev = np.concatenate(([-1], a[:-1]))
        keep = (a != blank_id) & (a != prev)
        out.append([int(x) for x in a[keep]])
    return out


def greedy_ids_loop(log_probs, lengths, blank_id: int, blank_bias: float = 0.0) -> List[List[int]]:
    """Frame-by-frame statement with the reference's explicit ``prev`` variable; used to
    check that the vectorised rule in ``greedy_ids`` is the same function."""
    lp = np.asarray(log_probs, dtype=np.float32)
    if blank_bias != 0.0:
        lp = lp.copy()
        lp[:, :, blank_id] = lp[:, :, blank_id] + np.float32(blank_bias)
    out = []
    for b, n in enumerate(lengths):
        toks: List[int] = []
        prev = None
        for t in range(int(n)):
            k = int(np.argmax(lp[b, t]))
            if k == blank_id:
                prev = k
            elif k != prev:
                toks.append(k)
                prev = k
        out.append(toks)
    return out


def ids_to_text(ids: Sequence[int], tokens: Sequence[str], blank_id: int, pad_id: int) -> str:
    return "".join(tokens[i] for i in ids if i != blank_id and i != pad_id and 0 <= i < len(tokens))


def greedy_text(log_probs, lengths, tokens, blank_id: int, pad_id: int, blank_bias: float = 0.0) -> List[str]:
    return [ids_to_text(s, tokens, blank_id, pad_id) for s in greedy_ids(log_probs, lengths, blank_id, blank_bias)]
