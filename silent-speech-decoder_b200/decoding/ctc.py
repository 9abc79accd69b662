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


_PUA = 0xE000  # private-use code points stand in for multi-character tokens while rows are assembled


def _rows_to_strings(tok: torch.Tensor, cnt: torch.Tensor, text: Sequence[str], normalise_space: bool) -> List[str]:
    """(B,T) int32 ids + (B,) counts on the device -> one string per row, ``text[id]`` concatenated.

    The reference builds every transcript token by token in Python; at config 3 (10 000 utterances) that costs
    more than the beam kernel itself.  Here the id matrix becomes ONE utf-32 buffer through a code-point table and
    is decoded once; rows are slices of that string.  Tokens whose text is not a single character (``<unk>`` in
    the greedy path) travel as private-use code points and are substituted per row only where they occur; empty
    texts are dropped before the decode."""
    n = cnt.cpu().numpy().astype(np.int64)
    B = n.shape[0]
    if B == 0:
        return []
    width = int(n.max())
    if width == 0:
        return [""] * B
    rows = tok[:, :width].cpu().numpy()
    lut = np.zeros(len(text), dtype="<u4")
    multi = {}
    for i, s in enumerate(text):
        if len(s) == 1:
            lut[i] = ord(s)
        elif len(s) > 1:
            lut[i] = _PUA + len(multi)
            multi[chr(_PUA + len(multi))] = s
    ids = np.clip(rows, 0, len(text) - 1)
    cp = lut[ids]
    valid = np.arange(width)[None, :] < n[:, None]
    empties = bool(((cp == 0) & valid).any())
    if empties:
        # rare (ids whose text is empty): compact row by row semantics through a mask
        keep = valid & (cp != 0)
        kept = keep.sum(axis=1)
        ends = np.cumsum(kept)
        big = cp[keep].tobytes().decode("utf-32-le")
        out = [big[lo:hi] for lo, hi in zip((ends - kept).tolist(), ends.tolist())]
    else:
        cp = np.where(valid, cp, 0x20)
        big = np.ascontiguousarray(cp).tobytes().decode("utf-32-le")
        out = [big[i * width: i * width + k] for i, k in enumerate(n.tolist())]
    if multi:
        marks = "".join(multi)
        for i, s in enumerate(out):
            if any(m in s for m in marks):
                for m, rep in multi.items():
                    s = s.replace(m, rep)
                out[i] = s
    if normalise_space:
        out = [s if (s[:1] != " " and s[-1:] != " " and "  " not in s) else " ".join(s.split()) for s in out]
    return out


# Host-resident batches above this size are decoded in chunks on two alternating streams: chunk k+1's H2D copy and
# kernel run while chunk k's ids come back and its strings are assembled on the host (the reference hands its
# decoders CPU tensors, evaluate.py:299; at config 3 that is a 960 MB copy and ~20 ms of string work per call).
PIPELINE_MIN_BYTES = 96 << 20
PIPELINE_CHUNK = 0  # utterances per chunk; 0 = eight per SM (one wave of resident beam CTAs at width 50: measured best on B200)


_PIPE_STREAMS: dict = {}


def _pipelined(log_probs: torch.Tensor, lengths: torch.Tensor, launch, to_strings) -> List[str]:
    """``launch(x_dev, len_dev) -> tensors``; ``to_strings(*tensors) -> List[str]`` (does the D2H copies)."""
    B = log_probs.size(0)
    chunk = PIPELINE_CHUNK
    if chunk <= 0 and not log_probs.is_cuda and torch.cuda.is_available():
        chunk = 8 * torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    big = (not log_probs.is_cuda) and log_probs.dim() == 3 and log_probs.numel() * 4 >= PIPELINE_MIN_BYTES and chunk > 0 and B >= 2 * chunk
    if not big:
        return to_strings(*launch(log_probs, lengths))
    dev = ops._device(log_probs)
    ln = torch.as_tensor(lengths)
    streams = _PIPE_STREAMS.get(dev.index)
    if streams is None:  # two streams per device, kept: the scratch buffers in ops are cached per stream
        streams = _PIPE_STREAMS[dev.index] = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    for s in streams:
        s.wait_stream(torch.cuda.current_stream(dev))
    out: List[str] = []
    pending = None
    for k, lo in enumerate(range(0, B, chunk)):
        hi = min(B, lo + chunk)
        s = streams[k & 1]
        with torch.cuda.stream(s):
            x = log_probs[lo:hi].to(dev, non_blocking=True)
            res = launch(x, ln[lo:hi])
        if pending is not None:
            with torch.cuda.stream(pending[0]):
                out.extend(to_strings(*pending[1]))  # blocks on the OTHER stream only
        pending = (s, res)
    with torch.cuda.stream(pending[0]):
        out.extend(to_strings(*pending[1]))
    for s in streams:
        torch.cuda.current_stream(dev).wait_stream(s)
    return out


def _greedy_decode(log_probs: torch.Tensor, lengths: torch.Tensor, blank_id: int, blank_bias: float = 0.0) -> List[List[int]]:
    """Argmax per frame, collapse repeats/blanks (reference ctc.py:16-36). Returns id lists."""
    tok, cnt = ops.greedy_ids(log_probs, lengths, blank=blank_id, blank_bias=blank_bias)
    return _rows_to_lists(tok, cnt)


def build_greedy_decoder(vocab: Vocab, blank_bias: float = 0.0) -> DecoderFn:
    # Vocab.decode (vocab.py:44-53): blank and pad render as nothing, every other id as its token string
    greedy_text = ["" if i in (vocab.blank_id, vocab.pad_id) else vocab.tokens[i] for i in range(vocab.size)]

    def decode(log_probs: torch.Tensor, lengths: torch.Tensor) -> List[str]:
        return _pipelined(log_probs, lengths,
                          lambda x, n: ops.greedy_ids(x, n, blank=vocab.blank_id, blank_bias=blank_bias),
                          lambda tok, cnt: _rows_to_strings(tok, cnt, greedy_text, normalise_space=False))

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
    text, space = _beam_label_strings(vocab)
    pad = vocab.pad_id if (vocab.pad_id != vocab.blank_id and 0 <= vocab.pad_id < vocab.size) else -1
    if lm_path:
        # ctc.py:67-72: build_ctcdecoder(labels, kenlm_model_path=str(lm_path), alpha, beta) -- word n-gram fusion
        # (scope row f-4).  The ARPA file becomes device tables once, here; every call then runs ssd_ctc_beam_lm.
        from .lm import build_tables

        if not str(lm_path).endswith(".arpa"):
            raise ValueError(f"lm_path must be an ARPA text file (what build_char_lm.py writes); got {lm_path}")
        keep = [i for i in range(vocab.size) if i not in {vocab.blank_id, vocab.pad_id}]
        tables = build_tables(lm_path, [""] + [text[i] for i in keep], alpha=alpha, beta=beta)

        def decode_lm(log_probs: torch.Tensor, lengths: torch.Tensor) -> List[str]:
            if log_probs.dim() == 3 and log_probs.size(-1) != vocab.size:
                raise ValueError(
                    f"Input logits shape is {tuple(log_probs.shape[1:])}, but vocabulary is size {vocab.size}"
                )

            def launch(x, n):
                tok, cnt, _ = ops.beam_lm_ids(x, n, vocab.blank_id, pad, space, tables,
                                              [(beam_width, beam_prune_logp, blank_bias, -5.0)])
                return tok[0], cnt[0]

            return _pipelined(log_probs, lengths, launch,
                              lambda tok, cnt: _rows_to_strings(tok, cnt, text, normalise_space=True))

        decode_lm.lm_tables = tables  # for callers that want the scores (ops.beam_lm_ids)
        return decode_lm
    del alpha, beta  # inert without an LM (ctc.py:67-72)

    def decode(log_probs: torch.Tensor, lengths: torch.Tensor) -> List[str]:
        if log_probs.dim() == 3 and log_probs.size(-1) != vocab.size:
            raise ValueError(
                f"Input logits shape is {tuple(log_probs.shape[1:])}, but vocabulary is size {vocab.size}"
            )
        return _pipelined(log_probs, lengths,
                          lambda x, n: ops.beam_ids(x, n, blank=vocab.blank_id, pad=pad, space=space, beam_width=beam_width,
                                                    beam_prune_logp=beam_prune_logp, blank_bias=blank_bias)[:2],
                          lambda tok, cnt: _rows_to_strings(tok, cnt, text, normalise_space=True))

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
