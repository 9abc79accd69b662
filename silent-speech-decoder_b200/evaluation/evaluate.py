"""Decode-and-score loop with the reference's metric names and output files, log-probs staying on the GPU.

Mirror of reference ``src/evaluation/evaluate.py`` for the part that sits around the decoder:

* ``compute_metrics(refs, hyps) -> {"wer", "cer"}``                      (reference :34-38, jiwer.wer / jiwer.cer)
* ``compute_error_breakdown(refs, hyps) -> {...}``                        (reference :41-99, same keys)
* ``DecoderSetting`` / ``GridEvaluator``: the reference's decoder sweep (``src/experiments/config_builder.py:42-93``,
  ``orchestrate.py:129-153,235-249``) re-runs the whole evaluation -- encoder included -- once per
  ``(method, width, prune, blank_bias)`` setting; here every setting decodes the same device-resident batch.
* ``Evaluator``: the body of the reference's batch loop (:290-328) -- decode, collect
  ``{"utterance_id", "ref", "hyp"}`` records, then ``metrics.json`` / ``predictions.jsonl`` in the reference's
  schema.  The reference copies every ``(B,T,48)`` log-prob batch to the host before decoding (:299) and
  defaults to batches of 4 (:158); here the decoder reads the batch where the model left it.

All edit-distance arithmetic runs in ``ssd_edit_counts`` (csrc/edit.cu); the host only turns strings into
integer ids (words through a dictionary, characters through their code points).  jiwer is not needed.
Checkpoint loading, the data loader and the CLI of the reference are outside this path.
"""

from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import ops
from ..data.vocab import Vocab
from ..decoding.ctc import DecoderFn, build_decoder

_KEYS = ("substitutions", "deletions", "insertions", "hits")


def _csr(seqs: List[np.ndarray]) -> Tuple[torch.Tensor, torch.Tensor, int]:
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    if seqs:
        off[1:] = np.cumsum([s.shape[0] for s in seqs])
    tok = np.concatenate(seqs).astype(np.int32, copy=False) if int(off[-1]) else np.zeros(0, dtype=np.int32)
    longest = int(np.diff(off).max()) if seqs else 0
    return torch.from_numpy(tok), torch.from_numpy(off), longest


def _word_ids(refs: Sequence[str], hyps: Sequence[str]) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    table: Dict[str, int] = {}
    enc = lambda s: np.fromiter((table.setdefault(w, len(table)) for w in s.split()), dtype=np.int32)  # noqa: E731
    return [enc(r) for r in refs], [enc(h) for h in hyps]


def _char_ids(strings: Sequence[str]) -> List[np.ndarray]:
    # code points of the stripped string; inner spaces are characters (jiwer's cer_default: Strip + ReduceToListOfListOfChars)
    return [np.frombuffer(s.strip().encode("utf-32-le"), dtype=np.int32) for s in strings]


def alignment_counts(ref_seqs: List[np.ndarray], hyp_seqs: List[np.ndarray]) -> np.ndarray:
    """(N,4) int64 [substitutions, deletions, insertions, hits] per pair, computed on the GPU."""
    if len(ref_seqs) != len(hyp_seqs):
        raise ValueError(f"{len(ref_seqs)} references for {len(hyp_seqs)} hypotheses")
    rt, ro, rmax = _csr(ref_seqs)
    ht, ho, hmax = _csr(hyp_seqs)
    return ops.edit_counts(rt, ro, ht, ho, rmax, hmax).cpu().numpy().astype(np.int64)


def _check_refs(ref_seqs: List[np.ndarray]) -> None:
    if any(r.shape[0] == 0 for r in ref_seqs):
        raise ValueError("one or more references are empty strings")  # jiwer refuses them too


def compute_metrics(refs: Sequence[str], hyps: Sequence[str]) -> Dict[str, float]:
    """WER and CER over the corpus: total edits / total reference tokens (reference :34-38)."""
    rw, hw = _word_ids(refs, hyps)
    rc, hc = _char_ids(refs), _char_ids(hyps)
    _check_refs(rw)
    n = len(rw)
    c = alignment_counts(rw + rc, hw + hc)
    w, ch = c[:n].sum(0), c[n:].sum(0)
    rate = lambda t: float(t[0] + t[1] + t[2]) / float(t[0] + t[1] + t[3])  # noqa: E731
    return {"wer": rate(w), "cer": rate(ch)}


def compute_error_breakdown(refs: Sequence[str], hyps: Sequence[str]) -> Dict[str, float]:
    """Word-level substitution / deletion / insertion / hit totals and rates (reference :85-99)."""
    rw, hw = _word_ids(refs, hyps)
    t = alignment_counts(rw, hw).sum(0) if len(rw) else np.zeros(4, dtype=np.int64)
    total_words = max(1.0, float(t[0] + t[1] + t[3]))
    out = {k: float(v) for k, v in zip(_KEYS, t)}
    out.update({"substitution_rate": float(t[0]) / total_words, "deletion_rate": float(t[1]) / total_words,
                "insertion_rate": float(t[2]) / total_words})
    return out


class Evaluator:
    """The reference's evaluation loop body (:290-328) as an object: ``add`` one batch at a time, ``finish`` once.

    ``decoder`` is a ``DecoderFn`` from ``decoding.ctc.build_decoder``; ``decoder_info`` is what the reference
    stores under ``metrics["decoder"]`` (:309-317)."""

    def __init__(self, decoder, decoder_info: Optional[Dict] = None):
        self.decoder = decoder
        self.decoder_info = dict(decoder_info or {})
        self.refs: List[str] = []
        self.hyps: List[str] = []
        self.records: List[Dict] = []

    def add(self, utterance_ids: Iterable, transcripts: Sequence[str], log_probs: torch.Tensor, lengths: torch.Tensor) -> List[str]:
        hyps = self.decoder(log_probs, lengths)  # device tensors go straight to the kernel: no .cpu() hop
        return self.add_decoded(utterance_ids, transcripts, hyps)

    def add_decoded(self, utterance_ids: Iterable, transcripts: Sequence[str], hyps: Sequence[str]) -> List[str]:
        """Record transcripts that were decoded elsewhere (the grid evaluator's single launch, a sharded decode)."""
        for uid, ref, hyp in zip(utterance_ids, transcripts, hyps):
            self.refs.append(ref)
            self.hyps.append(hyp)
            self.records.append({"utterance_id": uid, "ref": ref, "hyp": hyp})
        return hyps

    def metrics(self, splits: Sequence[str] = (), subsets: Optional[Sequence[str]] = None, run_name: str = "eval_run") -> Dict:
        m: Dict = compute_metrics(self.refs, self.hyps)
        m["error_breakdown"] = compute_error_breakdown(self.refs, self.hyps)
        m["decoder"] = self.decoder_info
        m["data"] = {"splits": list(splits), "subsets": list(subsets) if subsets else None, "num_samples": len(self.refs)}
        m["run_name"] = run_name
        return m

    def finish(self, out_dir, splits: Sequence[str] = (), subsets: Optional[Sequence[str]] = None,
               run_name: str = "eval_run", config: Optional[Dict] = None) -> Dict:
        """Write ``metrics.json`` and ``predictions.jsonl`` (and ``config_used.json`` when given) like :319-323."""
        out = Path(out_dir)
        out.mkdir(parents=True, exist_ok=True)
        m = self.metrics(splits, subsets, run_name)
        if config is not None:
            (out / "config_used.json").write_text(json.dumps(config, indent=2))
        (out / "metrics.json").write_text(json.dumps(m, indent=2))
        with (out / "predictions.jsonl").open("w") as f:
            for rec in self.records:
                f.write(json.dumps(rec) + "\n")
        return m


@dataclass
class DecoderSetting:
    """One point of the decoder sweep; same fields as reference ``config_builder.py:42-52``."""

    name: str
    method: str = "greedy"
    beam_width: Optional[int] = None
    alpha: Optional[float] = None
    beta: Optional[float] = None
    beam_prune_logp: Optional[float] = None
    blank_bias: float = 0.0
    use_lm: bool = False
    lm_path: Optional[Path] = None

    def resolved(self) -> Dict:
        """The values the reference ends up decoding with: ``orchestrate.py:141-153`` builds the command line,
        ``evaluate.py:197-212`` fills what is missing (width 50, alpha 0.6, beta 0, prune -10 for beam)."""
        beam = self.method.lower() == "beam"
        return {
            "type": self.method,
            "beam_width": (self.beam_width or 50) if beam else None,
            "alpha": (0.6 if self.alpha is None else self.alpha) if beam else None,
            "beta": (0.0 if self.beta is None else self.beta) if beam else None,
            "beam_prune_logp": (-10.0 if self.beam_prune_logp is None else self.beam_prune_logp) if beam else None,
            "blank_bias": float(self.blank_bias),
            "lm_path": str(self.lm_path) if (beam and self.use_lm and self.lm_path) else None,  # evaluate.py:316
        }

    def build(self, vocab: Vocab) -> DecoderFn:
        r = self.resolved()
        return build_decoder(method=self.method, vocab=vocab, lm_path=r["lm_path"], beam_width=int(r["beam_width"] or 50),
                             alpha=float(r["alpha"] or 0.0), beta=float(r["beta"] or 0.0),
                             beam_prune_logp=float(-10.0 if r["beam_prune_logp"] is None else r["beam_prune_logp"]),
                             blank_bias=float(self.blank_bias))


class GridEvaluator:
    """Every setting of a decoder grid over the same batches: log-probs are moved to the GPU once per batch and ALL
    beam settings of the grid are decoded by ONE kernel launch (``ssd_ctc_beam_grid``: a work item is
    (setting, utterance)); greedy settings take one greedy launch each.  The reference runs its whole evaluation --
    encoder included -- once per setting (``orchestrate.py:129-153,235-249``).

    Settings that ask for a language model run through the LM-fused kernel (``ssd_ctc_beam_lm``, one launch per
    setting: alpha and beta belong to the tables) when their ARPA file exists, and are reported under ``skipped``
    when it does not -- what the reference does (``orchestrate.py:117-119,428``)."""

    MAX_PER_LAUNCH = 8  # SSD_MAX_BEAM_SETTINGS

    def __init__(self, vocab: Vocab, grid: Sequence[DecoderSetting]):
        from ..decoding import ctc as _ctc

        names = [g.name for g in grid]
        if len(set(names)) != len(names):
            raise ValueError(f"decoder setting names must be unique: {names}")
        self.vocab = vocab
        def lm_missing(g):
            return g.use_lm and not (g.lm_path and Path(g.lm_path).exists())

        self.skipped = [g.name for g in grid if lm_missing(g)]
        self.settings = [g for g in grid if not lm_missing(g)]
        self.evaluators: Dict[str, Evaluator] = {g.name: Evaluator(g.build(vocab), g.resolved()) for g in self.settings}
        self._beam_text, self._space = _ctc._beam_label_strings(vocab)
        self._pad = vocab.pad_id if (vocab.pad_id != vocab.blank_id and 0 <= vocab.pad_id < vocab.size) else -1
        self._rows_to_strings = _ctc._rows_to_strings

    def add(self, utterance_ids: Iterable, transcripts: Sequence[str], log_probs: torch.Tensor, lengths: torch.Tensor) -> Dict[str, List[str]]:
        dev = ops._device(log_probs)
        lp = log_probs.detach().to(dev, non_blocking=True)  # one host->device copy shared by all settings
        ln = torch.as_tensor(lengths).to(dev, non_blocking=True)
        ids = list(utterance_ids)
        out: Dict[str, List[str]] = {}
        beams = [g for g in self.settings if g.method.lower() == "beam" and not g.use_lm]
        for g in self.settings:
            if g.method.lower() == "beam" and g.use_lm:  # the decoder built for it holds the LM's device tables
                out[g.name] = self.evaluators[g.name].add(ids, transcripts, lp, ln)
        for lo in range(0, len(beams), self.MAX_PER_LAUNCH):
            part = beams[lo: lo + self.MAX_PER_LAUNCH]
            spec = []
            for g in part:
                r = g.resolved()
                spec.append((int(r["beam_width"]), float(r["beam_prune_logp"]), float(g.blank_bias), -5.0))
            tok, cnt, _ = ops.beam_grid_ids(lp, ln, self.vocab.blank_id, self._pad, self._space, spec)
            for i, g in enumerate(part):
                hyps = self._rows_to_strings(tok[i], cnt[i], self._beam_text, normalise_space=True)
                out[g.name] = self.evaluators[g.name].add_decoded(ids, transcripts, hyps)
        for g in self.settings:
            if g.method.lower() != "beam":
                out[g.name] = self.evaluators[g.name].add(ids, transcripts, lp, ln)
        return out

    def finish(self, out_root=None, **kw) -> Dict[str, Dict]:
        """-> {setting name: metrics}; with ``out_root`` each setting also gets ``<out_root>/<name>/`` files."""
        if out_root is None:
            return {name: e.metrics(**kw) for name, e in self.evaluators.items()}
        return {name: e.finish(Path(out_root) / name, **kw) for name, e in self.evaluators.items()}
