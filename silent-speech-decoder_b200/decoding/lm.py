"""Host side of language-model fusion (scope row f-4): ARPA file -> the device tables of ``ssd_ctc_beam_lm``.

The reference passes ``lm_path`` (an ARPA file written by ``src/decoding/build_char_lm.py:35-40``: ``lmplz -o 5``
over whitespace-separated transcripts, i.e. a word n-gram model) to ``pyctcdecode.build_ctcdecoder`` together with
``alpha`` and ``beta`` (``src/decoding/ctc.py:67-72``).  Here the file is parsed once into

* a dense character trie of the unigrams' spellings, indexed by the decoder's label numbers (what pyctcdecode's
  ``CharTrie`` answers: "is this open word a prefix of any unigram"; the node also gives the word id);
* a forward trie of all n-grams as one open-addressing hash table: ``(id of (w1..w_{n-1}), w_n) -> id of (w1..w_n)``
  with fp32 log10 probability and back-off weight per id (what ``kenlm.Model.BaseScore`` walks);

which the beam kernel reads from HBM/L2 (``include/ssd_ctc.h``, ``csrc/beam.cu``).  No CPU scoring path exists here:
``LanguageModelTables`` only builds and owns buffers.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

MAX_ORDER = 6
UNK_SCORE_OFFSET = -10.0  # pyctcdecode DEFAULT_UNK_LOGP_OFFSET
_M64 = (1 << 64) - 1


def lm_hash(keys: np.ndarray) -> np.ndarray:
    """``ssd_lm_hash`` on a uint64 array."""
    with np.errstate(over="ignore"):
        h = keys.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    return h ^ (h >> np.uint64(29))


def parse_arpa(path) -> Tuple[int, List[Dict[Tuple[str, ...], Tuple[float, float]]]]:
    """-> (order, [ {ngram words: (log10 prob, log10 backoff)} per order ]); values as the text gives them."""
    tables: List[Dict[Tuple[str, ...], Tuple[float, float]]] = []
    section = 0
    with open(path, "r", encoding="utf-8") as f:
        for raw in f:
            line = raw.rstrip("\r\n")
            if not line.strip():
                continue
            if line[0] == "\\":
                if line.endswith("-grams:"):
                    section = int(line[1 : line.index("-")])
                    while len(tables) < section:
                        tables.append({})
                else:
                    section = 0
                continue
            if section == 0:
                continue
            cols = line.split("\t")
            if len(cols) >= 2:
                words = tuple(cols[1].split(" "))
                backoff = float(cols[2]) if len(cols) > 2 and cols[2].strip() else 0.0
            else:  # space-separated file
                cols = line.split()
                words = tuple(cols[1 : 1 + section])
                backoff = float(cols[1 + section]) if len(cols) > 1 + section else 0.0
            if len(words) != section:
                raise ValueError(f"{path}: malformed {section}-gram line: {line!r}")
            tables[section - 1][words] = (float(cols[0]), backoff)
    if not tables or not tables[0]:
        raise ValueError(f"{path}: no unigrams found (not an ARPA file?)")
    return len(tables), tables


@dataclass
class LanguageModelTables:
    """Numpy tables + (after ``to``) their device copies and the ``ssd_lm_tables`` struct."""

    order: int
    words: List[str]
    trie_child: np.ndarray      # int32 [nodes, n_labels]
    trie_word: np.ndarray       # int32 [nodes]
    ngram_keys: np.ndarray      # uint64 [slots]
    ngram_vals: np.ndarray      # int32 [slots]
    ngram_prob: np.ndarray      # float32 [n_ngrams]
    ngram_backoff: np.ndarray   # float32 [n_ngrams]
    bos: int
    eos: int
    unk: int
    alpha: float
    beta: float
    unk_score_offset: float = UNK_SCORE_OFFSET
    eos_rescore: int = 0  # see ssd_lm_tables.eos_rescore: which of upstream's two score-cache variants the final pass follows
    _dev: Dict[int, tuple] = field(default_factory=dict, repr=False)

    @property
    def n_labels(self) -> int:
        return int(self.trie_child.shape[1])

    def struct_for(self, dev: torch.device):
        """Device copies (made once per device) wrapped in the C struct."""
        from .. import _native

        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        held = self._dev.get(idx)
        if held is None:
            arrs = [self.trie_child, self.trie_word, self.ngram_keys.view(np.int64), self.ngram_vals, self.ngram_prob,
                    self.ngram_backoff]
            held = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in arrs)
            self._dev[idx] = held
        tc, tw, nk, nv, npb, nbo = held
        return _native.LmTables(tc.data_ptr(), tw.data_ptr(), nk.data_ptr(), nv.data_ptr(), npb.data_ptr(), nbo.data_ptr(),
                                ctypes.c_uint64(len(self.ngram_keys) - 1), int(self.trie_child.shape[0]), self.n_labels,
                                int(self.ngram_prob.shape[0]), int(self.order), int(self.bos), int(self.eos), int(self.unk), int(self.eos_rescore),
                                float(self.alpha), float(self.beta), float(self.unk_score_offset))


def _insert_all(keys: np.ndarray, vals: np.ndarray, slots: int) -> Tuple[np.ndarray, np.ndarray]:
    """Linear-probing insertion of distinct non-zero uint64 keys, vectorised: every round places the keys whose
    current slot is free and unclaimed by an earlier key of the same round; the rest move one slot on."""
    tk = np.zeros(slots, dtype=np.uint64)
    tv = np.full(slots, -1, dtype=np.int32)
    mask = np.uint64(slots - 1)
    pos = lm_hash(keys) & mask
    todo = np.arange(len(keys))
    while len(todo):
        p = pos[todo]
        free = tk[p] == 0
        # first claimant per slot among the free ones
        cand = todo[free]
        _, first = np.unique(pos[cand], return_index=True)
        win = cand[first]
        tk[pos[win]] = keys[win]
        tv[pos[win]] = vals[win]
        placed = np.zeros(len(keys), dtype=bool)
        placed[win] = True
        todo = todo[~placed[todo]]
        pos[todo] = (pos[todo] + np.uint64(1)) & mask
    return tk, tv


def build_tables(arpa_path, label_text: Sequence[str], alpha: float, beta: float,
                 unk_score_offset: float = UNK_SCORE_OFFSET) -> LanguageModelTables:
    """``label_text[k]``: the text beam label k emits (label 0 = blank = ""), numbered as the kernel numbers them."""
    order, tables = parse_arpa(arpa_path)
    if order > MAX_ORDER:
        raise ValueError(f"{arpa_path}: order {order} exceeds the kernel's limit of {MAX_ORDER}")
    for k, t in enumerate(label_text):
        if k > 0 and len(t) != 1:
            raise ValueError(f"label {k} emits {t!r}: LM fusion needs one character per label")
    words = [w[0] for w in tables[0]]
    wid = {w: i for i, w in enumerate(words)}
    # ---- n-gram ids: unigram id = word id, higher orders numbered on
    ids: Dict[Tuple[int, ...], int] = {(i,): i for i in range(len(words))}
    prob = [tables[0][(w,)][0] for w in words]
    back = [tables[0][(w,)][1] for w in words]
    keys: List[int] = []
    vals: List[int] = []
    for n in range(2, order + 1):
        for gram, (p, b) in tables[n - 1].items():
            try:
                g = tuple(wid[w] for w in gram)
            except KeyError:
                continue  # an n-gram over a word without a unigram entry cannot be reached
            parent = ids.get(g[:-1])
            if parent is None:
                continue  # no (w1..w_{n-1}) entry: unreachable by a forward walk (kenlm refuses such files)
            i = len(prob)
            ids[g] = i
            prob.append(p)
            back.append(b)
            keys.append((((parent + 1) << 32) | g[-1]) & _M64)
            vals.append(i)
    slots = 64
    while slots < 2 * max(1, len(keys)):
        slots <<= 1
    tk, tv = _insert_all(np.asarray(keys, dtype=np.uint64), np.asarray(vals, dtype=np.int32), slots)
    # ---- character trie of the unigram spellings over the decoder's labels
    char2lab = {t: k for k, t in enumerate(label_text) if k > 0}
    n_labels = len(label_text)
    child: List[List[int]] = [[-1] * n_labels]
    word_at: List[int] = [-1]
    for w, i in wid.items():
        node = 0
        whole = True
        for ch in w:
            k = char2lab.get(ch)
            if k is None:
                whole = False  # the rest of this word cannot be typed; its typed prefix still counts as a prefix
                break
            nxt = child[node][k]
            if nxt < 0:
                nxt = len(child)
                child[node][k] = nxt
                child.append([-1] * n_labels)
                word_at.append(-1)
            node = nxt
        if whole and node > 0:
            word_at[node] = i
    return LanguageModelTables(
        order=order, words=words,
        trie_child=np.asarray(child, dtype=np.int32).reshape(len(child), n_labels),
        trie_word=np.asarray(word_at, dtype=np.int32),
        ngram_keys=tk, ngram_vals=tv,
        ngram_prob=np.asarray(prob, dtype=np.float32), ngram_backoff=np.asarray(back, dtype=np.float32),
        bos=wid.get("<s>", -1), eos=wid.get("</s>", -1), unk=wid.get("<unk>", -1),
        alpha=float(alpha), beta=float(beta), unk_score_offset=float(unk_score_offset),
    )
