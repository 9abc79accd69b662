"""n-gram LM fusion oracle (test infrastructure only -- see oracle/__init__.py).

PARITY UNPINNED, twice over.  The reference hands ``lm_path`` to third-party ``pyctcdecode``
(``src/decoding/ctc.py:49-52,67-72``), which scores words with third-party ``kenlm``; neither is in
``/root/reference`` nor installable here.  The LM file is what ``src/decoding/build_char_lm.py:35-40``
makes: ``lmplz -o 5 --text corpus --arpa out`` over whitespace-separated transcripts, i.e. a WORD
n-gram model in ARPA text format (the script calls it a character LM; ``lmplz`` tokenises on spaces).

Restated here [from memory of pyctcdecode 0.5.0 ``language_model.py`` / ``decoder.py`` and of KenLM's
``lm/model.cc``]:

  ArpaModel.base_score   kenlm.Model.BaseScore: log10 p(word | state) by back-off -- the longest n-gram
                         ending in the word, plus the back-off weights of every context longer than the
                         matched one, accumulated in fp32 from the shorter context upwards; unknown words
                         score as <unk> (``lmplz`` always writes one)
  LanguageModel.score    BaseScore (+ unk_score_offset for out-of-vocabulary words) (+ log10 p(</s> | end
                         state) for the last word) times alpha * ln 10, plus beta -- in Python floats
  score_partial_token    unk_score_offset if the open word is not a prefix of any unigram, times
                         len/6 once it is longer than 6 characters
  build_ctcdecoder       an ``.arpa`` path makes the decoder read its unigram list from the file;
                         unk_score_offset = -10, lm_score_boundary = True (start state <s>)

How the search uses them is in ``oracle/beam.py`` (``_score_lm``), including upstream's cache quirk: the
end-of-sentence term is only added for texts that were never scored as closed texts before.

Known uncertainty to settle when the upstream fixtures exist (``make_beam_golden_upstream.py``): the restatement
follows the release whose score cache is keyed by the text alone (``Dict[str, ...]``).  The library's main branch
keys it by ``(text, is_eos)``; under that variant EVERY final beam is re-scored with the ``</s>`` term, and a final
beam whose text ends closed scores an EMPTY last word (an out-of-vocabulary word to kenlm: ``<unk>`` + the unk offset).
If the fixtures disagree on end-of-utterance scores, this is the first place to look; the per-frame search is the
same under both.
"""

from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

LOG_BASE_CHANGE_FACTOR = 1.0 / math.log10(math.e)
AVG_TOKEN_LEN = 6
DEFAULT_UNK_LOGP_OFFSET = -10.0
UNKNOWN_MISSING_LOGPROB = np.float32(-100.0)  # KenLM's default when the file has no <unk>


class ArpaModel:
    """The ARPA file as kenlm.Model sees it: fp32 log10 probabilities and back-offs per n-gram."""

    def __init__(self, path):
        self.ngrams: Dict[Tuple[str, ...], Tuple[np.float32, np.float32]] = {}
        self.order = 0
        section = 0
        with open(path, "r", encoding="utf-8") as f:
            for line in f:
                line = line.rstrip("\n")
                if not line.strip():
                    continue
                if line.startswith("\\") and line.endswith("-grams:"):
                    section = int(line[1 : line.index("-")])
                    self.order = max(self.order, section)
                    continue
                if line.startswith("\\"):  # \data\ , \end\
                    section = 0
                    continue
                if section == 0:
                    continue  # "ngram n=count" header lines
                cols = line.split("\t") if "\t" in line else line.split()
                if "\t" in line:
                    prob, words = cols[0], cols[1].split(" ")
                    backoff = cols[2] if len(cols) > 2 else "0"
                else:
                    prob, words = cols[0], cols[1 : 1 + section]
                    backoff = cols[1 + section] if len(cols) > 1 + section else "0"
                self.ngrams[tuple(words)] = (np.float32(prob), np.float32(backoff))
        self.vocab = {k[0] for k in self.ngrams if len(k) == 1}

    def __contains__(self, word: str) -> bool:
        return word in self.vocab

    def unigrams(self) -> List[str]:
        return sorted(self.vocab)

    def begin_state(self) -> Tuple[str, ...]:
        return ("<s>",)

    def base_score(self, state: Tuple[str, ...], word: str) -> Tuple[np.float32, Tuple[str, ...]]:
        w = word if word in self.vocab else "<unk>"
        ctx = tuple(state[-(self.order - 1) :]) if self.order > 1 else ()
        if (w,) in self.ngrams:
            prob = self.ngrams[(w,)][0]
        else:
            prob = UNKNOWN_MISSING_LOGPROB
        matched = 1
        for n in range(1, len(ctx) + 1):
            key = ctx[len(ctx) - n :] + (w,)
            if key not in self.ngrams:
                break
            prob = self.ngrams[key][0]
            matched = n + 1
        for n in range(matched, len(ctx) + 1):
            c = ctx[len(ctx) - n :]
            if c in self.ngrams:
                prob = np.float32(prob + self.ngrams[c][1])
        new_state = (ctx + (w,))[-(self.order - 1) :] if self.order > 1 else ()
        return np.float32(prob), new_state


class LanguageModel:
    """pyctcdecode.language_model.LanguageModel over an ArpaModel, unigrams read from the same file."""

    def __init__(self, arpa: ArpaModel, alpha: float = 0.5, beta: float = 1.5,
                 unk_score_offset: float = DEFAULT_UNK_LOGP_OFFSET, score_boundary: bool = True):
        self.model = arpa
        self.alpha, self.beta = float(alpha), float(beta)
        self.unk_score_offset = float(unk_score_offset)
        self.score_boundary = score_boundary
        self.unigram_set = set(arpa.vocab)
        self.prefixes = {w[:k] for w in self.unigram_set for k in range(1, len(w) + 1)}

    @property
    def order(self) -> int:
        return self.model.order

    def start_state(self):
        return self.model.begin_state() if self.score_boundary else ()

    def score_partial_token(self, partial: str) -> float:
        is_oov = int(partial not in self.prefixes)
        unk = self.unk_score_offset * is_oov
        if len(partial) > AVG_TOKEN_LEN:
            unk = unk * len(partial) / AVG_TOKEN_LEN
        return unk

    def score(self, prev_state, word: str, is_last_word: bool = False):
        s, end_state = self.model.base_score(prev_state, word)
        lm = float(s)
        if word not in self.unigram_set:
            lm += self.unk_score_offset
        if is_last_word and self.score_boundary:
            e, _ = self.model.base_score(end_state, "</s>")
            lm = lm + float(e)
        return self.alpha * lm * LOG_BASE_CHANGE_FACTOR + self.beta, end_state


# --------------------------------------------------------------------------- a tiny ARPA writer for tests
def write_arpa(path, corpus: Iterable[str], order: int = 3, discount: float = 0.4, seed: int = 0) -> None:
    """Count-based back-off model over ``corpus`` lines (absolute discounting, interpolation left out; the
    numbers need not be a good LM, only a well-formed one with the closure properties ``lmplz`` output has:
    every prefix and every suffix of a listed n-gram is listed).  Stand-in for ``lmplz`` (absent here)."""
    rng = np.random.default_rng(seed)
    counts: List[Dict[Tuple[str, ...], int]] = [dict() for _ in range(order)]
    for line in corpus:
        words = ["<s>"] + line.split() + ["</s>"]
        for n in range(1, order + 1):
            for i in range(len(words) - n + 1):
                g = tuple(words[i : i + n])
                counts[n - 1][g] = counts[n - 1].get(g, 0) + 1
    counts[0].setdefault(("<unk>",), 0)
    total = sum(c for g, c in counts[0].items() if g != ("<s>",))
    with open(path, "w", encoding="utf-8") as f:
        f.write("\\data\\\n")
        for n in range(order):
            f.write(f"ngram {n + 1}={len(counts[n])}\n")
        for n in range(order):
            f.write(f"\n\\{n + 1}-grams:\n")
            extended = {h[:-1] for h in counts[n + 1]} if n + 1 < order else set()
            for g in sorted(counts[n]):
                c = counts[n][g]
                if n == 0:
                    p = max(c - discount, 0.05) / (total + 1.0)
                    if g == ("<s>",):
                        p = 1e-99  # lmplz writes -99 for <s>
                else:
                    p = max(c - discount, 0.05) / counts[n - 1][g[:-1]]
                logp = math.log10(p) if g != ("<s>",) else -99.0
                has_ext = g in extended
                if n + 1 < order:
                    bo = -float(rng.uniform(0.05, 0.9)) if has_ext else 0.0
                    f.write(f"{logp:.6f}\t{' '.join(g)}\t{bo:.6f}\n")
                else:
                    f.write(f"{logp:.6f}\t{' '.join(g)}\n")
        f.write("\n\\end\\\n")
