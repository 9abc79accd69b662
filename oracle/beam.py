"""Prefix beam search oracle (test infrastructure only -- see oracle/__init__.py).

PARITY UNPINNED.  The reference delegates the search to third-party
``pyctcdecode`` (``src/decoding/ctc.py:57,67-72,91-95``; un-pinned in
``environment.yml:59``, latest release 0.5.0), which is not present in
``/root/reference`` and cannot be installed in this image.  This file restates
(a) the reference's own wrapper, ``src/decoding/ctc.py:61-97`` [read in tree], and
(b) the published behaviour of ``pyctcdecode.BeamSearchDecoderCTC.decode`` with no
language model and no hotwords [from memory of 0.5.0; SURVEY.md section 8c], namely

  alphabet      ``<unk>``-like labels become U+2047, label 0 ("") is blank, " " closes a word
  conditioning  rows that do not sum to 1 are log-softmaxed in fp32 and clipped to [ln 1e-15, 0]
  per frame     candidates = {i : row[i] >= token_min_logp} | {argmax}; every live beam is
                extended by every candidate: blank / repeat keep the text, " " closes the open
                word, anything else appends to the open word
  merge         beams with equal (closed text, open word, last char) are summed (logaddexp, fp64)
  prune         drop below best + beam_prune_logp, keep the beam_width best
  history       ``decode`` forces prune_history=True; without an LM the key is
                (last closed word, open word, last char) and only the best beam per key survives
  finish        close the open word, merge, prune, best text with whitespace normalised

Scores accumulate in float64 (pyctcdecode needs numpy<2, where ``float + np.float32`` is
float64); per-frame log-probs are fp32 values.
Exact score ties are broken by generation order (ascending label index, then beam rank),
which upstream leaves to Python ``set`` iteration order; parity inputs avoid exact ties.
"""

from __future__ import annotations

import heapq
import math
import re
from typing import List, Optional, Sequence, Tuple

import numpy as np

UNK_GLYPH = "⁇"
MIN_TOKEN_CLIP_P = 1e-15
DEFAULT_TOKEN_MIN_LOGP = -5.0
_UNK_RE = re.compile(r"^[<\[]unk[>\]]$", flags=re.IGNORECASE)
_PAD_RE = re.compile(r"^[<\[]pad[>\]]$", flags=re.IGNORECASE)


# --------------------------------------------------------------------------- wrapper
def beam_labels(tokens: Sequence[str], blank_id: int, pad_id: int) -> Tuple[List[int], List[str]]:
    """ctc.py:64-65 -- blank first, pad dropped, the rest in vocabulary order."""
    keep = [i for i in range(len(tokens)) if i not in {blank_id, pad_id}]
    return keep, [""] + [tokens[i] for i in keep]


def wrapper_matrix(log_probs: np.ndarray, blank_id: int, pad_id: int, blank_bias: float, keep: Sequence[int]) -> np.ndarray:
    """ctc.py:75-87 -- (B,T,V) fp32 -> (B,T,1+len(keep)) fp32 with pad folded into blank.
    The bias is added before the pad merge, all in fp32."""
    lp = np.asarray(log_probs, dtype=np.float32)
    blank_col = lp[:, :, blank_id]
    if blank_bias != 0.0:
        blank_col = blank_col + float(blank_bias)
        blank_col = blank_col.astype(np.float32)
    if pad_id != blank_id and 0 <= pad_id < lp.shape[-1]:
        blank_col = np.logaddexp(blank_col, lp[:, :, pad_id])
    parts = [blank_col[:, :, None]]
    if len(keep):
        parts.append(lp[:, :, list(keep)])
    return np.concatenate(parts, axis=2)


# --------------------------------------------------------------------------- alphabet
def normalize_labels(labels: Sequence[str]) -> List[str]:
    """Non-BPE alphabet normalisation of ``build_ctcdecoder`` as it applies to the
    reference's label list: "|" stands in for a missing " ", pad-like labels and a lone
    "_" become the blank "", unk-like labels become U+2047."""
    out = list(labels)
    if "|" in out and " " not in out:
        out[out.index("|")] = " "
    for k, lab in enumerate(out):
        if _PAD_RE.match(lab):
            out[k] = ""
    if "_" in out and "" not in out:
        out[out.index("_")] = ""
    if "" not in out:
        out.append("")
    for k, lab in enumerate(out):
        if _UNK_RE.match(lab):
            out[k] = UNK_GLYPH
    return out


# --------------------------------------------------------------------------- conditioning
def _log_softmax_rows(x: np.ndarray) -> np.ndarray:
    m = np.amax(x, axis=1, keepdims=True)
    m[~np.isfinite(m)] = 0
    sh = x - m
    with np.errstate(divide="ignore"):
        norm = np.log(np.sum(np.exp(sh), axis=1, keepdims=True))
    return sh - norm


def condition_rows(rows: np.ndarray) -> np.ndarray:
    """decode_beams' input handling: probabilities -> log, anything else -> fp32 log-softmax,
    both clipped to [ln 1e-15, 0].  Stays in the input dtype (fp32 on the reference path)."""
    rows = np.asarray(rows)
    if rows.ndim != 2:
        raise ValueError(f"Input logits have {rows.ndim} dimensions, but need 2: (time, vocabulary)")
    if rows.shape[0] and math.isclose(float(rows.sum(axis=1).mean()), 1.0):
        return np.log(np.clip(rows, rows.dtype.type(MIN_TOKEN_CLIP_P), rows.dtype.type(1)))
    # numpy<2 (which pyctcdecode requires) keeps the array dtype when clipping against scalars
    lo = rows.dtype.type(np.log(MIN_TOKEN_CLIP_P))
    return np.clip(_log_softmax_rows(rows), lo, rows.dtype.type(0))


# --------------------------------------------------------------------------- search
def _logadd(a: float, b: float) -> float:
    if a >= b:
        return a + math.log(1 + math.exp(b - a))
    return b + math.log(1 + math.exp(a - b))


def _join(text: str, word: str) -> str:
    if not word:
        return text
    if not text:
        return word
    return text + " " + word


# A beam is (text, next_word, word_part, last_char, score); frame bookkeeping that
# upstream carries for word timestamps does not influence text or score and is omitted.
def _merge(beams):
    table = {}
    for text, nxt, part, last, score in beams:
        key = (_join(text, nxt), part, last)
        if key in table:
            table[key] = (text, nxt, part, last, _logadd(table[key][4], score))
        else:
            table[key] = (text, nxt, part, last, score)
    return list(table.values())


def _score_no_lm(beams):
    # without LM / hotwords the ranking score is the acoustic score; text absorbs next_word
    return [(_join(text, nxt), "", part, last, score, score) for text, nxt, part, last, score in beams]


# Which key upstream's score cache uses.  "text" (the release this restatement follows, and what the kernel implements): a
# text that was scored before keeps its score, so the final pass adds no </s> term to it.  "text_eos" (the library's main
# branch): the final pass re-scores every beam with the </s> term, an already closed text scoring an EMPTY last word.
# Flip it when the upstream fixtures (make_beam_golden_upstream.py) say so; the per-frame search does not depend on it.
LM_CACHE_KEY = "text"


def _score_lm(beams, lm, cache, pcache, is_eos: bool = False):
    """``_get_lm_beams``: ranking score = acoustic score + LM score of the closed text + partial-word score.
    ``cache`` maps a closed text to (accumulated LM score, end state) and is shared by every frame of an
    utterance AND by the final end-of-sentence pass -- so, as upstream, a text that was scored before keeps
    its score without the </s> term."""
    out = []
    for text, nxt, part, last, score in beams:
        new_text = _join(text, nxt)
        if LM_CACHE_KEY == "text_eos" and is_eos:
            prev, state = cache[text]
            s, _ = lm.score(state, nxt, is_last_word=True)
            out.append((new_text, "", part, last, score, score + prev + s))  # (no open word in the final pass)
            continue
        if new_text not in cache:
            prev, state = cache[text]
            s, end_state = lm.score(state, nxt, is_last_word=is_eos)
            cache[new_text] = (prev + s, end_state)
        lm_score = cache[new_text][0]
        if part:
            if part not in pcache:
                pcache[part] = lm.score_partial_token(part)
            lm_score += pcache[part]
        out.append((new_text, "", part, last, score, score + lm_score))
    return out


def _cut(beams, beam_width: int, beam_prune_logp: float):
    # ranked and pruned by the LAST field (acoustic + LM); the acoustic score [4] is what the beam carries on
    top = max(b[-1] for b in beams)
    alive = [b for b in beams if b[-1] >= top + beam_prune_logp]
    return heapq.nlargest(beam_width, alive, key=lambda b: b[-1])


def _history_prune(beams, keep_words: int = 1):
    seen = set()
    out = []
    for b in beams:
        text, nxt, part, last, score = b[:5]
        key = (tuple(text.split()[-keep_words:]), part, last)
        if key not in seen:
            seen.add(key)
            out.append((text, nxt, part, last, score))
    return out


def search(
    rows: np.ndarray,
    labels: Sequence[str],
    beam_width: int,
    beam_prune_logp: float,
    token_min_logp: float = DEFAULT_TOKEN_MIN_LOGP,
    prune_history: bool = True,
    lm=None,
    with_lm_score: bool = False,
) -> List[Tuple]:
    """``_decode_logits``.  rows: conditioned (T,len(labels)) log-probs; ``lm``: an ``oracle.lm.LanguageModel``
    or None.  Returns the final beams as (text, acoustic score), best first -- with ``with_lm_score`` as
    (text, acoustic score, ranking score)."""
    cache = {"": (0.0, lm.start_state())} if lm is not None else None
    pcache: dict = {}
    keep_words = max(1, lm.order - 1) if lm is not None else 1

    def rank(merged, is_eos=False):
        return _score_no_lm(merged) if lm is None else _score_lm(merged, lm, cache, pcache, is_eos)

    beams = [("", "", "", None, 0.0)]
    for row in rows:
        cand = set(np.where(row >= token_min_logp)[0].tolist()) | {int(row.argmax())}
        grown = []
        for k in sorted(cand):
            p = float(row[k])  # fp32 value, fp64 accumulation
            ch = labels[k]
            for text, nxt, part, last, score in beams:
                if ch == "" or last == ch:
                    grown.append((text, nxt, part, ch, score + p))
                elif ch == " ":
                    grown.append((text, part, "", ch, score + p))
                else:
                    grown.append((text, nxt, part + ch, ch, score + p))
        kept = _cut(rank(_merge(grown)), beam_width, beam_prune_logp)
        beams = _history_prune(kept, keep_words) if prune_history else [b[:5] for b in kept]
    closed = [(text, part, "", None, score) for text, _, part, _, score in beams]
    final = _cut(rank(_merge(closed), is_eos=True), beam_width, beam_prune_logp)
    if with_lm_score:
        return [(" ".join(b[0].split()), b[4], b[5]) for b in final]
    return [(" ".join(b[0].split()), b[4]) for b in final]


# --------------------------------------------------------------------------- reference-shaped API
class BeamDecoder:
    """What ``build_beam_decoder(vocab, None, ...)`` returns, as a class so tests can also
    read scores.  ``tokens``/ids come from the reference ``Vocab`` (vocab.py:11-30)."""

    def __init__(
        self,
        tokens: Sequence[str],
        blank_id: int,
        pad_id: int,
        beam_width: int = 50,
        beam_prune_logp: float = -10.0,
        blank_bias: float = 0.0,
        token_min_logp: float = DEFAULT_TOKEN_MIN_LOGP,
        lm=None,
    ):
        self.lm = lm  # oracle.lm.LanguageModel, what build_ctcdecoder(kenlm_model_path=..., alpha, beta) sets up
        self.blank_id, self.pad_id = blank_id, pad_id
        self.beam_width, self.beam_prune_logp = beam_width, beam_prune_logp
        self.blank_bias, self.token_min_logp = blank_bias, token_min_logp
        self.keep, raw = beam_labels(tokens, blank_id, pad_id)
        self.labels = normalize_labels(raw)

    def decode_with_scores(self, log_probs, lengths) -> List[Tuple[str, float]]:
        mat = wrapper_matrix(np.asarray(log_probs), self.blank_id, self.pad_id, self.blank_bias, self.keep)
        if mat.shape[-1] != len(self.labels):
            raise ValueError("Input logits shape does not match alphabet size")
        out = []
        for i, n in enumerate(lengths):
            rows = condition_rows(mat[i, : int(n)])
            out.append(search(rows, self.labels, self.beam_width, self.beam_prune_logp, self.token_min_logp, lm=self.lm)[0])
        return out

    def decode_with_lm_scores(self, log_probs, lengths) -> List[Tuple[str, float, float]]:
        """(text, acoustic score, acoustic + LM score) of the best beam per utterance."""
        mat = wrapper_matrix(np.asarray(log_probs), self.blank_id, self.pad_id, self.blank_bias, self.keep)
        return [search(condition_rows(mat[i, : int(n)]), self.labels, self.beam_width, self.beam_prune_logp,
                       self.token_min_logp, lm=self.lm, with_lm_score=True)[0] for i, n in enumerate(lengths)]

    def __call__(self, log_probs, lengths) -> List[str]:
        return [t for t, _ in self.decode_with_scores(log_probs, lengths)]


def decode_one(args) -> Tuple[str, float]:
    """Picklable single-utterance entry for multiprocessing CPU baselines."""
    rows47, labels, width, prune, tmin = args
    return search(condition_rows(rows47), labels, width, prune, tmin)[0]
