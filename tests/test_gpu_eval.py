"""GPU parity tests for the evaluation loop around the decoders (scope row f-1): ssd_edit_counts through
the C ABI against the reference's own alignment counts (tests/golden/edit_ref.json) and oracle/edit.py,
and the Evaluator's files against the reference's schema (evaluate.py:290-328)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import beam as ob
from oracle import edit as oe
from oracle import greedy as og
from silent_speech_decoder_b200 import ops, synth
from silent_speech_decoder_b200.data.vocab import Vocab
from silent_speech_decoder_b200.decoding import ctc
from silent_speech_decoder_b200.evaluation import evaluate as ev

pytestmark = pytest.mark.gpu
V = Vocab.default()


def _vec(c):
    return [c["substitutions"], c["deletions"], c["insertions"], c["hits"]]


def _counts(refs, hyps):
    return ev.alignment_counts([np.asarray(r, dtype=np.int32) for r in refs], [np.asarray(h, dtype=np.int32) for h in hyps]).tolist()


def test_counts_match_reference_outputs(golden_dir):
    z = json.load(open(os.path.join(golden_dir, "edit_ref.json")))
    refs, hyps = [r["ref"] for r in z["pairs"]], [r["hyp"] for r in z["pairs"]]
    rw, hw = ev._word_ids(refs, hyps)
    assert ev.alignment_counts(rw, hw).tolist() == [r["word"] for r in z["pairs"]]
    enc = lambda s: np.frombuffer(s.strip().replace(" ", "_").encode("utf-32-le"), dtype=np.int32)  # noqa: E731
    assert ev.alignment_counts([enc(r) for r in refs], [enc(h) for h in hyps]).tolist() == [r["char"] for r in z["pairs"]]
    assert ev.compute_error_breakdown(refs, hyps) == z["corpus_breakdown"]


def test_counts_random_and_column_tiles():
    rng = np.random.default_rng(11)
    refs, hyps = [], []
    for n, m in [(300, 257), (40, 600), (520, 256), (1, 513), (257, 1), (0, 5), (5, 0), (0, 0), (256, 258), (90, 1100)]:
        r = rng.integers(0, 5, n).tolist()
        h = (r[: m // 2] + rng.integers(0, 5, m).tolist())[:m]
        refs.append(r); hyps.append(h)
    for _ in range(200):  # many small pairs: more pairs than warps in a CTA, ragged tail
        n, m = rng.integers(0, 30, 2)
        refs.append(rng.integers(0, 4, n).tolist()); hyps.append(rng.integers(0, 4, m).tolist())
    want = [_vec(oe.levenshtein_counts(r, h)) for r, h in zip(refs, hyps)]
    assert _counts(refs, hyps) == want


def test_counts_properties_at_scale():
    # size-independent properties at a size the Python oracle cannot do: identity, symmetry of the cost,
    # and the (cost, hits) identities n = S+D+H, m = S+I+H
    rng = np.random.default_rng(12)
    N = 20000
    refs = [rng.integers(0, 40, rng.integers(1, 200)).astype(np.int32) for _ in range(N)]
    hyps = [np.delete(r, rng.integers(0, r.shape[0], r.shape[0] // 7)) for r in refs]
    c = ev.alignment_counts(refs, hyps)
    n = np.array([r.shape[0] for r in refs]); m = np.array([h.shape[0] for h in hyps])
    assert (c[:, 0] + c[:, 1] + c[:, 3] == n).all() and (c[:, 0] + c[:, 2] + c[:, 3] == m).all()
    assert (c[:, 0] == 0).all() and (c[:, 2] == 0).all() and (c[:, 1] == n - m).all()  # pure deletions
    back = ev.alignment_counts(hyps, refs)
    assert (back[:, 2] == c[:, 1]).all() and (back[:, 3] == c[:, 3]).all()  # deletions become insertions
    same = ev.alignment_counts(refs, refs)
    assert (same[:, :3] == 0).all() and (same[:, 3] == n).all()


def test_metrics_against_oracle():
    lp = synth.peaked(48, 300, 4101)
    ln = synth.ragged_lengths(48, 300, 4102)
    hyps = ctc.build_decoder("beam", V, beam_width=20)(lp, ln)
    noisy = torch.log_softmax(lp + 2.5 * torch.randn(lp.shape, generator=torch.Generator().manual_seed(5)), -1)
    refs = [r if r.strip() else "x" for r in ctc.build_decoder("greedy", V)(noisy, ln)]
    m = ev.compute_metrics(refs, hyps)
    assert m["wer"] == pytest.approx(oe.wer(refs, hyps), abs=1e-12) and m["cer"] == pytest.approx(oe.cer(refs, hyps), abs=1e-12)
    assert 0 < m["cer"] < m["wer"]
    assert ev.compute_error_breakdown(refs, hyps) == oe.error_breakdown(refs, hyps)
    with pytest.raises(ValueError):
        ev.compute_metrics(["", "a"], ["b", "a"])


def test_evaluator_files(tmp_path):
    dec = ctc.build_decoder("beam", V, beam_width=50)
    e = ev.Evaluator(dec, {"type": "beam", "beam_width": 50, "alpha": 0.6, "beta": 0.0, "beam_prune_logp": -10.0,
                           "blank_bias": 0.0, "lm_path": None})
    refs_all = []
    for k in range(3):  # three batches, log-probs resident on the device
        lp = synth.peaked(16, 200, 4200 + k).cuda()
        ln = synth.ragged_lengths(16, 200, 4300 + k).cuda()
        refs = [t if t.strip() else "x" for t in ctc.build_decoder("greedy", V)(lp, ln)]
        hyps = e.add([f"utt{k}_{i}" for i in range(16)], refs, lp, ln)
        assert hyps == dec(lp.cpu(), ln.cpu())  # same transcripts as the reference's host-tensor call
        refs_all += refs
    m = e.finish(tmp_path / "run", splits=["silent_parallel_data"], subsets=["test"], run_name="r", config={"a": 1})
    on_disk = json.loads((tmp_path / "run" / "metrics.json").read_text())
    assert on_disk == m and set(m) == {"wer", "cer", "error_breakdown", "decoder", "data", "run_name"}
    assert m["data"] == {"splits": ["silent_parallel_data"], "subsets": ["test"], "num_samples": 48}
    assert m["wer"] == pytest.approx(oe.wer(e.refs, e.hyps), abs=1e-12)
    recs = [json.loads(l) for l in (tmp_path / "run" / "predictions.jsonl").read_text().splitlines()]
    assert [r["utterance_id"] for r in recs][:2] == ["utt0_0", "utt0_1"] and len(recs) == 48
    assert all(set(r) == {"utterance_id", "ref", "hyp"} for r in recs) and [r["ref"] for r in recs] == refs_all
    assert json.loads((tmp_path / "run" / "config_used.json").read_text()) == {"a": 1}


def test_decoder_grid_matches_single_runs(tmp_path):
    # the reference's FULL_DECODERS_SILENT grid (config_builder.py:86-92); the LM entry is skipped
    grid = [ev.DecoderSetting(name="greedy", method="greedy"),
            ev.DecoderSetting(name="beam20_bias", method="beam", beam_width=20, alpha=0.45, beta=0.0, beam_prune_logp=-10.0, blank_bias=0.2),
            ev.DecoderSetting(name="beam50", method="beam", beam_width=50, alpha=0.5, beta=0.0, beam_prune_logp=-12.0, blank_bias=0.1),
            ev.DecoderSetting(name="beam100", method="beam", beam_width=100, alpha=0.55, beta=0.05, beam_prune_logp=-12.0, blank_bias=0.05),
            ev.DecoderSetting(name="beam50_lm", method="beam", beam_width=50, use_lm=True, lm_path="results/lm/char_5gram.arpa")]
    g = ev.GridEvaluator(V, grid)
    assert g.skipped == ["beam50_lm"]
    lp = synth.peaked(12, 250, 4400)  # host tensors, as the reference's loop holds them
    ln = synth.ragged_lengths(12, 250, 4401)
    refs = [t if t.strip() else "x" for t in ctc.build_decoder("greedy", V)(synth.peaked(12, 250, 4402), ln)]
    n0 = ops.LAUNCHES["beam"]
    got = g.add(range(12), refs, lp, ln)
    assert ops.LAUNCHES["beam"] == n0 + 1  # the three beam settings shared ONE launch
    # every setting against the CPU oracle (greedy: pinned to the reference; beam: the restatement)
    assert got["greedy"] == og.greedy_text(lp.numpy(), ln.tolist(), V.tokens, V.blank_id, V.pad_id)
    for name, (w, p, b) in {"beam20_bias": (20, -10.0, 0.2), "beam50": (50, -12.0, 0.1), "beam100": (100, -12.0, 0.05)}.items():
        assert got[name] == ob.BeamDecoder(V.tokens, V.blank_id, V.pad_id, w, p, b)(lp.numpy(), ln.tolist()), name
    assert got["beam20_bias"] == ctc.build_decoder("beam", V, beam_width=20, beam_prune_logp=-10.0, blank_bias=0.2)(lp, ln)
    m = g.finish(tmp_path)
    assert set(m) == {"greedy", "beam20_bias", "beam50", "beam100"}
    assert m["beam50"]["decoder"] == {"type": "beam", "beam_width": 50, "alpha": 0.5, "beta": 0.0, "beam_prune_logp": -12.0,
                                      "blank_bias": 0.1, "lm_path": None}
    assert m["greedy"]["decoder"]["beam_width"] is None and (tmp_path / "beam100" / "predictions.jsonl").exists()
    with pytest.raises(ValueError):
        ev.GridEvaluator(V, [grid[0], grid[0]])
