"""Inputs for the language-model fusion tests (scope row f-4): a small word list, a corpus of sentences over it, an
ARPA file written from the corpus by ``oracle.lm.write_arpa`` (``lmplz`` is absent here) and posteriors that spell
sentences -- some in the corpus, some with out-of-vocabulary words -- noisily enough that the LM changes the ranking."""
from __future__ import annotations

import numpy as np

from oracle import lm as olm
from silent_speech_decoder_b200.data.vocab import Vocab

WORDS = ["the", "a", "cat", "cats", "car", "cart", "dog", "do", "door", "sat", "sad", "on", "one", "mat", "map", "man",
         "ran", "rat", "to", "too", "two", "it's", "its", "in", "inn", "big", "bag", "bad", "understanding", "under"]


def corpus(n: int = 400, seed: int = 0):
    rng = np.random.default_rng(seed)
    # a first-order chain over WORDS so that the higher orders carry information
    nxt = {w: rng.choice(len(WORDS), size=4, replace=False) for w in WORDS}
    lines = []
    for _ in range(n):
        w = WORDS[int(rng.integers(len(WORDS)))]
        line = [w]
        for _ in range(int(rng.integers(1, 7))):
            w = WORDS[int(rng.choice(nxt[w]))] if rng.random() < 0.8 else WORDS[int(rng.integers(len(WORDS)))]
            line.append(w)
        lines.append(" ".join(line))
    return lines


def write_lm(path, order: int = 3, seed: int = 0) -> None:
    olm.write_arpa(path, corpus(seed=seed), order=order, seed=seed)


def spelled_log_probs(sentences, T: int, seed: int, boost: float = 3.0, confusions: float = 1.5):
    """(B,T,V) fp32 log-probs laid on a CTC alignment of each sentence; ``boost`` is low and every emitted letter has a
    rival letter boosted by ``confusions``, so that several spellings compete and word scores matter."""
    v = Vocab.default()
    rng = np.random.default_rng(seed)
    B = len(sentences)
    x = rng.standard_normal((B, T, v.size)).astype(np.float32)
    lengths = []
    letters = [i for i, t in enumerate(v.tokens) if len(t) == 1 and t.isalpha()]
    for b, s in enumerate(sentences):
        ids = v.encode(s)
        ext = []
        for k, i in enumerate(ids):
            if k and ids[k - 1] == i:
                ext.append(v.blank_id)
            ext.append(i)
            if rng.random() < 0.5:
                ext.append(v.blank_id)
        assert len(ext) <= T, (len(ext), T)
        reps = np.bincount(rng.integers(0, len(ext), size=T - len(ext)), minlength=len(ext)) + 1
        path = np.repeat(np.asarray(ext), reps)[:T]
        n = int(rng.integers(max(len(ext), T - 10), T + 1))
        # shortening keeps the alignment's head; the tail frames are blank
        for t in range(T):
            x[b, t, path[t]] += boost
            if path[t] != v.blank_id and path[t] in letters:
                x[b, t, letters[int(rng.integers(len(letters)))]] += confusions
        lengths.append(T if n < len(ext) else T)
    lp = x - np.log(np.exp(x).sum(-1, keepdims=True))
    return lp.astype(np.float32), lengths


SENTENCES = ["the cat sat on the mat", "a dog ran to the door", "its a big bad rat", "the cart ran in two",
             "understanding the zebra", "one man sat", "it's too sad", "qqq the xyzzy cat"]
