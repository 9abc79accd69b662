"""Seeded synthetic inputs for the CTC hot path.

``flat``   : log_softmax of white noise -- what a random-init CTC head emits; nearly
             every label passes the beam's per-frame -5 filter (worst case).
``peaked`` : a random transcript laid on a random monotone CTC alignment with
             +8 logit on the aligned label -- what a trained model emits.
All generators are deterministic functions of ``seed`` and run on CPU so the
oracle and the GPU path see bit-identical inputs.
"""

from __future__ import annotations

from typing import Tuple

import torch

V_DEFAULT = 48
BLANK_DEFAULT = 1
SPACE_DEFAULT = 3


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def flat(B: int, T: int, seed: int, V: int = V_DEFAULT) -> torch.Tensor:
    """(B,T,V) fp32 log-probs, near-uniform."""
    x = torch.randn(B, T, V, generator=_gen(seed), dtype=torch.float32)
    return torch.log_softmax(x, dim=-1)


def peaked(
    B: int,
    T: int,
    seed: int,
    V: int = V_DEFAULT,
    blank: int = BLANK_DEFAULT,
    space: int = SPACE_DEFAULT,
    boost: float = 8.0,
) -> torch.Tensor:
    """(B,T,V) fp32 log-probs concentrated on a random valid CTC alignment."""
    g = _gen(seed)
    x = torch.randn(B, T, V, generator=g, dtype=torch.float32)
    n_tok = max(1, T // 5)
    for b in range(B):
        # transcript over ids 3..V-1, space with prob ~0.17, no leading/double space
        ids = torch.randint(4, V, (n_tok,), generator=g)
        is_sp = torch.rand(n_tok, generator=g) < 0.17
        toks = []
        for k in range(n_tok):
            t = space if bool(is_sp[k]) and toks and toks[-1] != space else int(ids[k])
            toks.append(t)
        # extended sequence with blanks where needed between repeats, else random
        ext = []
        for k, t in enumerate(toks):
            if k > 0 and (toks[k - 1] == t or float(torch.rand((), generator=g)) < 0.5):
                ext.append(blank)
            ext.append(t)
        ext = ext[:T]
        # monotone alignment: distribute the remaining frames as repeats
        extra = T - len(ext)
        cuts = torch.sort(torch.randint(0, len(ext), (extra,), generator=g)).values
        reps = torch.bincount(cuts, minlength=len(ext)) + 1
        align = torch.repeat_interleave(torch.tensor(ext), reps)[:T]
        x[b, torch.arange(T), align] += boost
    return torch.log_softmax(x, dim=-1)


def targets(
    B: int, S: int, seed: int, V: int = V_DEFAULT, lo: int = 2, s_min: int | None = None
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Zero-padded (B,S) int64 label matrix with ids in [lo,V) and lengths ~U[S/4,S]."""
    g = _gen(seed)
    s_min = max(1, S // 4) if s_min is None else s_min
    lens = torch.randint(s_min, S + 1, (B,), generator=g, dtype=torch.int64)
    tg = torch.randint(lo, V, (B, S), generator=g, dtype=torch.int64)
    mask = torch.arange(S)[None, :] < lens[:, None]
    return tg * mask, lens


def ragged_lengths(B: int, T: int, seed: int) -> torch.Tensor:
    """(B,) int64 frame counts ~U[T/2,T]."""
    return torch.randint(max(1, T // 2), T + 1, (B,), generator=_gen(seed), dtype=torch.int64)


# Relative frequencies of English text over the reference's character inventory (space 0.17, then e t a o i n ...),
# used for "trained-regime" inputs: real transcripts repeat a few labels often, uniform synthetic targets never do.
_ENGLISH = {" ": 0.170, "e": 0.104, "t": 0.075, "a": 0.067, "o": 0.062, "i": 0.058, "n": 0.056, "s": 0.052, "h": 0.050,
            "r": 0.049, "d": 0.035, "l": 0.033, "u": 0.023, "c": 0.023, "m": 0.020, "w": 0.019, "f": 0.018, "g": 0.016,
            "y": 0.016, "p": 0.016, "b": 0.012, "v": 0.008, "k": 0.006, "'": 0.003, "x": 0.002, "j": 0.001, "q": 0.001,
            "z": 0.001}


def english_label_weights(tokens, lo: int = 2) -> torch.Tensor:
    """(V,) sampling weights over vocabulary ids: English letter/space frequencies where the token is one of them, a
    small floor elsewhere, zero for the special ids below ``lo``."""
    w = torch.full((len(tokens),), 2e-4, dtype=torch.float64)
    for i, t in enumerate(tokens):
        if t in _ENGLISH:
            w[i] = _ENGLISH[t]
    w[:lo] = 0.0
    return w / w.sum()


def peaked_aligned(
    B: int,
    T: int,
    seed: int,
    V: int = V_DEFAULT,
    blank: int = BLANK_DEFAULT,
    space: int = SPACE_DEFAULT,
    boost: float = 8.0,
    n_labels=None,
    label_weights=None,
    return_alignment: bool = False,
):
    """Peaked posteriors TOGETHER WITH the transcript they were laid on: ``(log_probs (B,T,V), targets (B,S) int64
    zero-padded, target_lengths (B,))`` -- what a trained model emits for its own training pair.

    ``n_labels``: labels per utterance, an int or a ``(lo, hi)`` range sampled per utterance (default ``T // 5``);
    always capped so the alignment fits (``2 * n - 1 <= T`` covers the worst case of all-repeats).
    ``label_weights``: (V,) sampling weights (e.g. ``english_label_weights``); default uniform over ids 4..V-1 with
    spaces at probability 0.17 like ``peaked``.  No leading space, no double space.  Vectorised per utterance
    (10 000 x T=500 in a few seconds); its random stream differs from ``peaked``'s, whose fixtures stay as they are."""
    import numpy as np

    rng = np.random.default_rng(int(seed))
    noise = torch.randn(B, T, V, generator=_gen(seed), dtype=torch.float32)
    if n_labels is None:
        n_labels = max(1, T // 5)
    lo_n, hi_n = (n_labels, n_labels) if isinstance(n_labels, int) else n_labels
    cap = max(1, (T + 1) // 2)
    if label_weights is not None:
        wts = np.asarray(label_weights, dtype=np.float64)
        wts = wts / wts.sum()
    S_max = min(hi_n, cap)
    tg = np.zeros((B, S_max), dtype=np.int64)
    tl = np.zeros((B,), dtype=np.int64)
    align = np.empty((B, T), dtype=np.int64)
    for b in range(B):
        n = min(int(rng.integers(lo_n, hi_n + 1)), cap)
        if label_weights is not None:
            toks = rng.choice(V, size=n, p=wts)
        else:
            toks = rng.integers(4, V, size=n)
            toks[rng.random(n) < 0.17] = space
        # no leading space, no double space: such a space becomes a letter
        bad = (toks == space) & np.concatenate(([True], toks[:-1] == space))
        while bad.any():
            toks[bad] = rng.integers(4, V, size=int(bad.sum()))
            bad = (toks == space) & np.concatenate(([True], toks[:-1] == space))
        need_blank = np.concatenate(([False], toks[1:] == toks[:-1]))
        room = T - n - int(need_blank.sum())           # frames left after one per label and the mandatory blanks
        opt = (~need_blank) & (np.arange(n) > 0) & (rng.random(n) < 0.5)
        if int(opt.sum()) > room:                      # drop optional blanks that do not fit
            idx = np.flatnonzero(opt)
            opt[idx[room:]] = False
        with_blank = need_blank | opt
        ext = np.empty(n + int(with_blank.sum()), dtype=np.int64)
        pos = np.arange(n) + np.cumsum(with_blank)
        ext[:] = blank
        ext[pos] = toks
        extra = T - ext.shape[0]
        reps = np.bincount(rng.integers(0, ext.shape[0], size=extra), minlength=ext.shape[0]) + 1
        align[b] = np.repeat(ext, reps)[:T]
        tg[b, :n] = toks
        tl[b] = n
    a = torch.from_numpy(align)
    noise.scatter_add_(2, a.unsqueeze(-1), torch.full((B, T, 1), float(boost)))
    lp = torch.log_softmax(noise, dim=-1)
    out = (lp, torch.from_numpy(tg), torch.from_numpy(tl))
    return out + (a,) if return_alignment else out


# ------------------------------------------------------------------ language-model fusion inputs (scope row f-4)
def word_list(n_words: int, seed: int):
    """``n_words`` distinct lower-case "words" (2-9 letters, English letter frequencies)."""
    import numpy as np

    rng = np.random.default_rng(int(seed))
    letters = [c for c in _ENGLISH if c.isalpha()]
    p = np.asarray([_ENGLISH[c] for c in letters])
    p = p / p.sum()
    out, seen = [], set()
    while len(out) < n_words:
        w = "".join(rng.choice(letters, size=int(rng.integers(2, 10)), p=p))
        if w not in seen:
            seen.add(w)
            out.append(w)
    return out


def word_corpus(words, n_lines: int, seed: int, lo: int = 3, hi: int = 12):
    """Sentences over ``words``: Zipf-distributed unigrams with a sparse first-order chain, so that higher n-gram
    orders carry information (what transcripts of a small domain look like to ``lmplz``)."""
    import numpy as np

    rng = np.random.default_rng(int(seed))
    n = len(words)
    zipf = 1.0 / np.arange(1, n + 1)
    zipf /= zipf.sum()
    follow = rng.integers(0, n, size=(n, 4))
    lines = []
    for _ in range(n_lines):
        k = int(rng.choice(n, p=zipf))
        line = [words[k]]
        for _ in range(int(rng.integers(lo, hi + 1)) - 1):
            k = int(follow[k, int(rng.integers(4))]) if rng.random() < 0.75 else int(rng.choice(n, p=zipf))
            line.append(words[k])
        lines.append(" ".join(line))
    return lines


def write_arpa(path, corpus, order: int = 5, discount: float = 0.5) -> None:
    """A well-formed ARPA file from n-gram counts of ``corpus`` (absolute discounting; every prefix and suffix of a
    listed n-gram is listed, as in ``lmplz`` output -- ``lmplz`` itself, which the reference shells out to at
    ``build_char_lm.py:35-40``, is not in this image).  The probabilities are plausible, not a tuned model."""
    import math

    counts = [dict() for _ in range(order)]
    for line in corpus:
        w = ["<s>"] + line.split() + ["</s>"]
        for n in range(1, order + 1):
            for i in range(len(w) - n + 1):
                g = tuple(w[i: i + n])
                counts[n - 1][g] = counts[n - 1].get(g, 0) + 1
    counts[0].setdefault(("<unk>",), 0)
    total = sum(c for g, c in counts[0].items() if g != ("<s>",)) + 1.0
    with open(path, "w", encoding="utf-8") as f:
        f.write("\\data\\\n")
        for n in range(order):
            f.write(f"ngram {n + 1}={len(counts[n])}\n")
        for n in range(order):
            f.write(f"\n\\{n + 1}-grams:\n")
            longer = {}
            if n + 1 < order:
                for h in counts[n + 1]:
                    longer[h[:-1]] = longer.get(h[:-1], 0) + 1
            for g in sorted(counts[n]):
                c = counts[n][g]
                denom = total if n == 0 else counts[n - 1][g[:-1]]
                logp = -99.0 if g == ("<s>",) else math.log10(max(c - discount, 0.1) / denom)
                if n + 1 < order:
                    types = longer.get(g, 0)
                    bo = math.log10(min(1.0, discount * types / max(c, 1)) + 1e-3) if types else 0.0
                    f.write(f"{logp:.6f}\t{' '.join(g)}\t{bo:.6f}\n")
                else:
                    f.write(f"{logp:.6f}\t{' '.join(g)}\n")
        f.write("\n\\end\\\n")


def spelled(sentences, token_to_id, T: int, seed: int, V: int = V_DEFAULT, blank: int = BLANK_DEFAULT, boost: float = 5.0,
            rival: float = 2.5):
    """(log_probs (B,T,V) fp32, lengths (B,) int64): posteriors laid on a random CTC alignment of each sentence's
    characters, with a random rival letter boosted on every emitting frame so that several spellings compete and an LM
    has something to decide.  Sentences are cut to what fits T frames."""
    import numpy as np

    rng = np.random.default_rng(int(seed))
    B = len(sentences)
    x = torch.randn(B, T, V, generator=_gen(seed), dtype=torch.float32).numpy()
    letter_ids = np.asarray([i for t, i in token_to_id.items() if len(t) == 1 and t.isalpha()])
    for b, s in enumerate(sentences):
        ids = np.asarray([token_to_id[c] for c in s if c in token_to_id][: max(1, (T - 2) // 2)], dtype=np.int64)
        n = len(ids)
        need_blank = np.concatenate(([False], ids[1:] == ids[:-1]))
        opt = (~need_blank) & (np.arange(n) > 0) & (rng.random(n) < 0.5)
        room = T - n - int(need_blank.sum())
        if int(opt.sum()) > room:
            idx = np.flatnonzero(opt)
            opt[idx[room:]] = False
        with_blank = need_blank | opt
        ext = np.full(n + int(with_blank.sum()), blank, dtype=np.int64)
        ext[np.arange(n) + np.cumsum(with_blank)] = ids
        reps = np.bincount(rng.integers(0, ext.shape[0], size=T - ext.shape[0]), minlength=ext.shape[0]) + 1
        path = np.repeat(ext, reps)[:T]
        x[b, np.arange(T), path] += boost
        emit = np.flatnonzero(path != blank)
        x[b, emit, letter_ids[rng.integers(0, len(letter_ids), size=len(emit))]] += rival
    lp = torch.log_softmax(torch.from_numpy(x), dim=-1)
    return lp, torch.full((B,), T, dtype=torch.int64)
