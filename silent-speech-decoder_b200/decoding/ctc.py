"""CTC decoders with the reference's call signatures, running on sm_100a kernels.

Drop-in for reference ``src/decoding/ctc.py``: same public names (``DecoderFn``,
``_greedy_decode``, ``build_greedy_decoder``, ``build_beam_decoder``, ``build_decoder``),
same keyword arguments and defaults, same return types.  Differences a caller can see:

* tensors may be on the host or on the GPU; host tensors are copied to the device
  (the reference copies the other way, ``evaluate.py:299``);
* ``pyctcdecode`` is not needed: the search is the native kernel ``ssd_ctc_beam``;
* a language model is out of scope: ``lm_path`` other than None raises.
"""

from __future__ import annotations

import re
from pathlib import Path
from typing import Callable, List, Sequence

import numpy as np
import torch

from ..data.vocab import Vocab
from .. import ops

DecoderFn = Callable[[torch.Tensor, torch.Tensor], List[str]]

_UNK_LIKE = re.compile(r"^[<\[]unk[>\]]$", flags=re.IGNORECASE)
_PAD_LIKE = re.compile(r"^[<\[]pad[>\]]$", flags=re.IGNORECASE)
UNK_GLYPH = "⁇"


def _rows_to_lists(tok: torch.Tensor, cnt: torch.Tensor) -> List[List[int]]:
    """(B,T) int32 + (B,) counts on the device -> python lists, one D2H copy each."""
    n = cnt.cpu().numpy()
    B = n.shape[0]
    if B == 0:
        return []
    width = int(n.max()) if B else 0
    rows = tok[:, : max(width, 1)].cpu().numpy()
    keep = np.arange(rows.shape[1])[None, :] < n[:, None]
    flat = rows[keep].tolist()
    ends = np.cumsum(n).tolist()
    out, lo = [], 0
    for hi in ends:
        out.append(flat[lo:hi])
        lo = hi
    return out


def _greedy_decode(log_probs: torch.Tensor, lengths: torch.Tensor, blank_id: int, blank_bias: float = 0.0) -> List[List[int]]:
    """Argmax per frame, collapse repeats/blanks (reference ctc.py:16-36). Returns id lists."""
    tok, cnt = ops.greedy_ids(log_probs, lengths, blank=blank_id, blank_bias=blank_bias)
    return _rows_to_lists(tok, cnt)


def build_greedy_decoder(vocab: Vocab, blank_bias: float = 0.0) -> DecoderFn:
    def decode(log_probs: torch.Tensor, lengths: torch.Tensor) -> List[str]:
        seqs = _greedy_decode(log_probs, lengths, blank_id=vocab.blank_id, blank_bias=blank_bias)
        return [vocab.decode(seq) for seq in seqs]

    return decode


def _beam_label_strings(vocab: Vocab) -> tuple[List[str], int]:
    """Text emitted per vocabulary id and the id acting as word separator, following what
    ``build_ctcdecoder`` does to the label list the reference builds at ctc.py:64-65."""
    keep = [i for i in range(vocab.size) if i not in {vocab.blank_id, vocab.pad_id}]
    labels = [vocab.tokens[i] for i in keep]
    if "|" in labels and " " not in labels:
        labels[labels.index("|")] = " "
    text = [""] * vocab.size
    space = -1
    for i, lab in zip(keep, labels):
        if _PAD_LIKE.match(lab):
            lab = ""
        elif _UNK_LIKE.match(lab):
            lab = UNK_GLYPH
        text[i] = lab
        if lab == " " and space < 0:
            space = i
    return text, space


def build_beam_decoder(
    vocab: Vocab,
    lm_path: Path | None = None,
    beam_width: int = 50,
    alpha: float = 0.6,
    beta: float = 0.0,
    beam_prune_logp: float = -10.0,
    blank_bias: float = 0.0,
) -> DecoderFn:
    if lm_path:
        raise NotImplementedError(
            "language-model fusion is outside the native CTC hot path (lm_path must be None); "
            "alpha/beta have no effect without an LM, exactly as in pyctcdecode"
        )
    del alpha, beta  # inert without an LM (ctc.py:67-72)
    text, space = _beam_label_strings(vocab)
    pad = vocab.pad_id if (vocab.pad_id != vocab.blank_id and 0 <= vocab.pad_id < vocab.size) else -1

    def decode(log_probs: torch.Tensor, lengths: torch.Tensor) -> List[str]:
        if log_probs.dim() == 3 and log_probs.size(-1) != vocab.size:
            raise ValueError(
                f"Input logits shape is {tuple(log_probs.shape[1:])}, but vocabulary is size {vocab.size}"
            )
        tok, cnt, _ = ops.beam_ids(
            log_probs, lengths, blank=vocab.blank_id, pad=pad, space=space, beam_width=beam_width,
            beam_prune_logp=beam_prune_logp, blank_bias=blank_bias,
        )
        hyps = []
        for seq in _rows_to_lists(tok, cnt):
            hyps.append(" ".join("".join(text[i] for i in seq).split()))
        return hyps

    return decode


def build_decoder(
    method: str,
    vocab: Vocab,
    lm_path: Path | None = None,
    beam_width: int = 50,
    alpha: float = 0.6,
    beta: float = 0.0,
    beam_prune_logp: float = -10.0,
    blank_bias: float = 0.0,
) -> DecoderFn:
    # anything that is not "beam" means greedy, as in the reference (ctc.py:112-122)
    if method.lower() == "beam":
        return build_beam_decoder(vocab=vocab, lm_path=lm_path, beam_width=beam_width, alpha=alpha, beta=beta,
                                  beam_prune_logp=beam_prune_logp, blank_bias=blank_bias)
    return build_greedy_decoder(vocab, blank_bias=blank_bias)
