"""CPU tier: the kernel SOURCES, compiled by g++ against the SIMT emulator (tests/emu/), against the
oracle on small cases.  Checks kernel logic (barriers, mbarrier pipelines, index arithmetic) without a
GPU; the product never loads this build.  Also run under reversed / random thread schedules, which is
what exposes missing barriers."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import beam as ob
from oracle import ctc_loss as ol
from oracle import edit as oe
from oracle import greedy as og
from silent_speech_decoder_b200 import synth
from silent_speech_decoder_b200.data.vocab import Vocab
from tests.emu import harness

V = Vocab.default()
SPACE = V.token_to_id[" "]


@pytest.fixture(scope="module")
def lib():
    return harness.library()


def test_greedy(lib, golden_dir):
    z = np.load(os.path.join(golden_dir, "greedy_ref.npz"))
    assert harness.greedy(lib, z["kat.lp"], [5, 5], 0) == [[1, 2], [1, 2]]
    for gen, B, T, bias in (("peaked", 4, 300, 0.0), ("flat", 3, 130, 0.5)):
        x = getattr(synth, gen)(B, T, 3 + T).numpy()
        ln = [T, T - 1, T // 2, 0][:B]
        assert harness.greedy(lib, x, ln, 1, bias) == og.greedy_ids(x, ln, 1, bias)


@pytest.mark.parametrize("name", ["peaked50", "flat50", "flat20_bias", "flat1", "flat200"])
def test_beam_frozen(lib, golden_dir, name):
    z = np.load(os.path.join(golden_dir, "beam_restatement.npz"))
    width, prune, bias = z[f"{name}.cfg"]
    ids, score = harness.beam(lib, z[f"{name}.lp"], z[f"{name}.len"], V.blank_id, V.pad_id, SPACE, int(width), float(prune), float(bias))
    text = ["".join(ob.UNK_GLYPH if i == V.unk_id else V.tokens[i] for i in s) for s in ids]
    assert text == list(z[f"{name}.text"])
    assert np.abs(score - z[f"{name}.score"]).max() <= 1e-4


@pytest.mark.parametrize("name", ["flat_edges", "peaked", "test_models_shape", "tight"])
def test_loss_golden(lib, golden_dir, name):
    z = np.load(os.path.join(golden_dir, "loss_ref.npz"))
    c = {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith(name + ".")}
    nll, grad = harness.loss(lib, c["lp"], c["il"], c["tg"], c["tl"], int(c["blank"]), True, mean_den=c["lp"].shape[0])
    np.testing.assert_allclose(nll, c["nll64"], rtol=1e-5, atol=1e-6)
    assert np.abs(grad - c["grad64"]).max() <= 1e-5 * np.abs(c["grad64"]).max() + 1e-12
    fwd, none = harness.loss(lib, c["lp"], c["il"], c["tg"], c["tl"], int(c["blank"]), True, want_grad=False)
    assert none is None
    np.testing.assert_allclose(fwd, c["nll64"], rtol=1e-5, atol=1e-6)


def test_loss_wide_lattice_and_extreme_range(lib):
    # NC=20 code path, and a peaked input against unrelated targets (cells thousands of binades apart)
    for lp, il, S in ((synth.flat(1, 300, 17), [300], 290), (synth.peaked(1, 600, 77, boost=10.0), [600], 120)):
        tg, tl = synth.targets(1, S, 18)
        tl[:] = S
        nll, grad = harness.loss(lib, lp.numpy(), il, tg.numpy(), tl.numpy(), 1, True, mean_den=1)
        n64, _, g64 = ol.torch_ctc(lp, torch.tensor(il), tg, tl, 1, True, torch.float64)
        np.testing.assert_allclose(nll, n64.numpy(), rtol=1e-5)
        assert np.abs(grad - g64.numpy()).max() <= 1e-5 * g64.abs().max().item()


def test_loss_fast_path_flags_and_fallback(lib):
    """The fp32 fast path checks its own result per utterance; what it flags is recomputed by the fp64 kernel.
    Flags (FastParams::flags): 0 fast result kept, 1 not covered, 2 log-prob range, 4 non-finite strip,
    8 posterior mass, 16 Z."""
    # plain utterances stay on the fast path and meet the tolerance on their own
    lp = synth.flat(3, 200, 41)
    tg, tl = synth.targets(3, 60, 42)
    il = torch.tensor([200, 150, 97])
    nll, grad, flags = harness.loss(lib, lp.numpy(), il.numpy(), tg.numpy(), tl.numpy(), 1, True, mean_den=3, with_fallbacks="flags")
    assert flags.tolist() == [0, 0, 0]
    n64, _, g64 = ol.torch_ctc(lp, il, tg, tl, 1, True, torch.float64)
    np.testing.assert_allclose(nll, n64.numpy(), rtol=1e-5)
    assert np.abs(grad - g64.numpy()).max() <= 1e-5 * g64.abs().max().item()
    # rows the fast path must hand over: no frames, infeasible (S > T), a log-prob below the exp range
    lp = synth.flat(5, 40, 43).clone()
    tg, tl = synth.targets(5, 12, 44)
    il = torch.tensor([40, 0, 5, 40, 40])
    tl[:] = torch.tensor([12, 3, 12, 6, 6])
    lp[3, 7, 2] = -95.0
    lp[4, 9, 5] = -120.0
    nll, grad, flags = harness.loss(lib, lp.numpy(), il.numpy(), tg.numpy(), tl.numpy(), 1, True, mean_den=5, with_fallbacks="flags")
    assert flags[0] == 0 and flags[1] == 1 and (flags[2] & 16) and (flags[3] & 2) and (flags[4] & 2)
    n64, _, g64 = ol.torch_ctc(lp, il, tg, tl, 1, True, torch.float64)
    np.testing.assert_allclose(nll, n64.numpy(), rtol=1e-5, atol=1e-6)
    assert np.abs(grad - g64.numpy()).max() <= 1e-5 * g64.abs().max().item()
    # confident posteriors that disagree with the target: cells of one lane thousands of binades apart.  Whatever
    # fp32 cannot hold there is seen by the per-frame mass check (flag 8) and recomputed by the fp64 kernel; since
    # strips behind a mass front take the front's basis (renorm_exchange) most such rows stay on the fast path.
    # Flagged or not, what comes back meets the tolerance.
    lp = synth.peaked(2, 700, 77, boost=10.0)
    tg, tl = synth.targets(2, 120, 18)
    tl[:] = 120
    il = torch.tensor([700, 700])
    nll, grad, flags = harness.loss(lib, lp.numpy(), il.numpy(), tg.numpy(), tl.numpy(), 1, True, mean_den=2, with_fallbacks="flags")
    n64, _, g64 = ol.torch_ctc(lp, il, tg, tl, 1, True, torch.float64)
    np.testing.assert_allclose(nll, n64.numpy(), rtol=1e-5)
    assert np.abs(grad - g64.numpy()).max() <= 1e-5 * g64.abs().max().item()


def test_loss_fast_path_keeps_the_trained_regime(lib):
    """What a training run feeds the loss once the model has learnt something: confident posteriors laid on the
    utterance's OWN transcript, English label frequencies (space ~0.14-0.17, 'e' ~0.10: long class ranges in the
    class-sum stage).  The mass front is then a cliff of ~2^150 per strip; every row must stay on the fp32 path
    (flag 0 -- a flagged row costs a second, fp64, pass) and meet the tolerance.  Shapes: dense (T ~ 2S), sparse
    (T ~ 20 S) and in between."""
    w = synth.english_label_weights(V.tokens)
    for B, T, nl, seed in ((4, 600, (100, 120), 1), (4, 400, (100, 130), 2), (4, 500, (30, 200), 3), (4, 240, (100, 120), 5),
                           (4, 800, (20, 60), 6)):
        lp, tg, tl = synth.peaked_aligned(B, T, seed + T, n_labels=nl, label_weights=w)
        il = torch.full((B,), T)
        nll, grad, flags = harness.loss(lib, lp.numpy(), il.numpy(), tg.numpy(), tl.numpy(), 1, True, mean_den=B, with_fallbacks="flags")
        assert flags.tolist() == [0] * B, (T, nl, flags.tolist())
        n64, _, g64 = ol.torch_ctc(lp, il, tg, tl, 1, True, torch.float64)
        np.testing.assert_allclose(nll, n64.numpy(), rtol=1e-5)
        assert np.abs(grad - g64.numpy()).max() <= 1e-5 * g64.abs().max().item()


def test_loss_fast_path_short_targets_in_a_wide_strip(lib):
    """The strip width is chosen per launch from the longest target, so short targets share it: few live lanes, steep
    opposite alpha / beta slopes inside a lane, lane exponents more than 2^126 apart.  These must stay correct (the
    boundary factor is a product of two floats, the posterior scale is split between the two recomputed sweeps) and
    whatever fp32 cannot hold must be flagged, never returned."""
    B, T, S, V = 6, 600, 200, 48
    g = torch.Generator().manual_seed(5)
    tl = torch.tensor([200, 15, 22, 40, 17, 60])
    tg = torch.randint(2, V, (B, S), generator=g, dtype=torch.int64) * (torch.arange(S)[None, :] < tl[:, None])
    il = torch.full((B,), T, dtype=torch.int64)
    lp = synth.flat(B, T, 7, V)
    nll, grad, flags = harness.loss(lib, lp.numpy(), il.numpy(), tg.numpy(), tl.numpy(), 1, True, mean_den=B, with_fallbacks="flags")
    n64, _, g64 = ol.torch_ctc(lp, il, tg, tl, 1, True, torch.float64)
    np.testing.assert_allclose(nll, n64.numpy(), rtol=1e-5)
    assert np.abs(grad - g64.numpy()).max() <= 1e-5 * g64.abs().max().item()
    assert flags[0] == 0 and flags[5] == 0 and flags[3] == 0


def test_loss_fast_path_partial_blocks_and_forward_only(lib):
    # input lengths around the 8-frame block size, odd lengths, one label, forward only
    for T in (1, 2, 7, 8, 9, 15, 17, 31):
        lp = synth.flat(2, T, 300 + T)
        S = max(1, min(6, T // 2))
        tg, tl = synth.targets(2, S, 5 + T)
        tl[:] = torch.tensor([S, max(1, S // 2)])
        il = torch.tensor([T, max(1, T - 1)])
        nll, grad, flags = harness.loss(lib, lp.numpy(), il.numpy(), tg.numpy(), tl.numpy(), 1, True, mean_den=2, with_fallbacks="flags")
        n64, _, g64 = ol.torch_ctc(lp, il, tg, tl, 1, True, torch.float64)
        np.testing.assert_allclose(nll, n64.numpy(), rtol=1e-5, atol=1e-6)
        assert np.abs(grad - g64.numpy()).max() <= 1e-5 * g64.abs().max().item() + 1e-12, T
        fwd, none = harness.loss(lib, lp.numpy(), il.numpy(), tg.numpy(), tl.numpy(), 1, True, want_grad=False)
        assert none is None
        np.testing.assert_allclose(fwd, n64.numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("order", ["rev", "rnd:3", "rnd:8"])
def test_schedule_independence(order):
    """Same results when the emulator visits threads in another order (separate process: the
    order is read at launch time)."""
    code = (
        "import numpy as np, os\n"
        "from tests.emu import harness\n"
        "from silent_speech_decoder_b200 import synth\n"
        "lib = harness.library()\n"
        "z = np.load('tests/golden/loss_ref.npz')\n"
        "c = {k.split('.',1)[1]: z[k] for k in z.files if k.startswith('flat_edges.')}\n"
        "nll, g = harness.loss(lib, c['lp'], c['il'], c['tg'], c['tl'], 1, True, mean_den=6)\n"
        "assert np.abs(g - c['grad64']).max() <= 1e-5 * np.abs(c['grad64']).max()\n"
        "x = synth.flat(2, 40, 2205).numpy()\n"
        "ids, sc = harness.beam(lib, x, [40, 33], 1, 0, 3, 50, -10.0)\n"
        "print(ids, [round(float(s), 6) for s in sc])\n"
    )
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for o in ("fwd", order):
        r = subprocess.run([sys.executable, "-c", code], cwd=repo, env={**os.environ, "SSD_EMU_ORDER": o, "PYTHONPATH": repo},
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout)
    assert outs[0] == outs[1]


def _edit_vec(r, h):
    c = oe.levenshtein_counts(list(r), list(h))
    return [c["substitutions"], c["deletions"], c["insertions"], c["hits"]]


def test_edit_counts_golden_and_edges(lib, golden_dir):
    import json

    z = json.load(open(os.path.join(golden_dir, "edit_ref.json")))
    table = {}
    enc = lambda s: [table.setdefault(w, len(table)) for w in s.split()]
    refs = [enc(r["ref"]) for r in z["pairs"]]
    hyps = [enc(r["hyp"]) for r in z["pairs"]]
    got = harness.edit_counts(lib, refs, hyps)
    assert got.tolist() == [r["word"] for r in z["pairs"]]
    crefs = [[ord(c) for c in r["ref"].strip().replace(" ", "_")] for r in z["pairs"]]
    chyps = [[ord(c) for c in r["hyp"].strip().replace(" ", "_")] for r in z["pairs"]]
    got = harness.edit_counts(lib, crefs, chyps)
    assert got.tolist() == [r["char"] for r in z["pairs"]]
    assert harness.edit_counts(lib, [], []).shape == (0, 4)


def test_edit_counts_column_tiles(lib):
    # hypotheses wider than one 256-column tile: the right edge of a tile is the next one's left edge
    rng = np.random.default_rng(7)
    refs, hyps = [], []
    for n, m in [(300, 257), (40, 600), (520, 256), (1, 513), (257, 1), (64, 300), (255, 255), (256, 258)]:
        r = rng.integers(0, 6, n).tolist()
        h = (r[: m // 2] + rng.integers(0, 6, m).tolist())[:m]
        refs.append(r); hyps.append(h)
    got = harness.edit_counts(lib, refs, hyps)
    assert got.tolist() == [_edit_vec(r, h) for r, h in zip(refs, hyps)]


def test_loss_fast_path_several_utterances_per_cta(lib, monkeypatch):
    # 2 CTAs for 7 ragged utterances: scalars, label tables, rings and the late worker's overlay of the sweeps'
    # ring space are rebuilt per utterance
    monkeypatch.setenv("SSD_LOSS_GRID", "2")
    lp = synth.flat(7, 90, 31).numpy()
    tg, tl = synth.targets(7, 20, 32)
    il = synth.ragged_lengths(7, 90, 33).numpy()
    il[3] = 90
    want_nll, _, want_g = ol.torch_ctc(torch.from_numpy(lp), torch.from_numpy(il), tg, tl, 1, True, torch.float64)
    nll, g, flags = harness.loss(lib, lp, il, tg.numpy(), tl.numpy(), 1, mean_den=7, with_fallbacks="flags")
    assert not flags.any()
    np.testing.assert_allclose(nll, want_nll.numpy(), rtol=2e-6)
    assert np.abs(g - want_g.numpy()).max() <= 1e-5 * np.abs(want_g.numpy()).max()
    monkeypatch.setenv("SSD_LOSS_NOLATE", "1")
    nll2, g2, _ = harness.loss(lib, lp, il, tg.numpy(), tl.numpy(), 1, mean_den=7, with_fallbacks="flags")
    np.testing.assert_allclose(nll2, nll, rtol=1e-6)
    assert np.abs(g2 - g).max() <= 2e-6 * np.abs(g).max()


# ---------------------------------------------------------------- ABI v2: logits input, decoder grid, conditional scale
def _logits_case(B, T, S, seed, scale=3.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, V.size, generator=g) * scale + torch.randn(B, T, 1, generator=g) * 5.0  # rows far from normalised
    il = synth.ragged_lengths(B, T, seed + 1)
    tg, tl = synth.targets(B, S, seed + 2)
    return x, il, tg, tl


@pytest.mark.parametrize("exact", [False, True])
def test_loss_from_logits(lib, exact, monkeypatch):
    """Row f-2: the loss from the head's pre-softmax output equals log_softmax (heads.py:36-37) followed by the
    reference loss, and the gradient is the one autograd gives with respect to the LOGITS."""
    if exact:
        monkeypatch.setenv("SSD_LOSS_EXACT", "1")
    x, il, tg, tl = _logits_case(5, 90, 14, 77)
    nll, grad = harness.loss(lib, x.numpy(), il, tg, tl, V.blank_id, True, mean_den=5, from_logits=True)
    x64 = x.double().requires_grad_(True)
    lp64 = torch.log_softmax(x64, dim=-1)
    n64 = torch.nn.functional.ctc_loss(lp64.transpose(0, 1), tg, il, tl, blank=V.blank_id, reduction="none", zero_infinity=True)
    (n64 / tl.clamp_min(1)).mean().backward()
    np.testing.assert_allclose(nll, n64.detach().numpy(), rtol=1e-5, atol=1e-6)
    g64 = x64.grad.numpy()
    assert np.abs(grad - g64).max() <= 1e-5 * np.abs(g64).max()
    # the pre-pass alone
    np.testing.assert_allclose(harness.row_lse(lib, x.numpy()), torch.logsumexp(x.double(), -1).numpy(), rtol=0, atol=2e-6)


def test_beam_grid_one_launch(lib):
    """Row f-3: one launch over several (width, prune, bias) settings == separate launches == the oracle."""
    B, T = 4, 70
    x = torch.cat([synth.flat(2, T, 5), synth.peaked(2, T, 6)]).numpy()
    ln = [T, T - 9, T, 31]
    settings = [(100, -12.0, 0.05, -5.0), (50, -12.0, 0.1, -5.0), (20, -10.0, 0.2, -5.0), (50, -10.0, 0.0, -5.0)]
    got = harness.beam_grid(lib, x, ln, V.blank_id, V.pad_id, SPACE, settings)
    for (w, p, b, t), (ids, score) in zip(settings, got):
        ids1, score1 = harness.beam(lib, x, ln, V.blank_id, V.pad_id, SPACE, w, p, b, t)
        assert ids == ids1 and np.array_equal(score, score1)
        want = ob.BeamDecoder(V.tokens, V.blank_id, V.pad_id, w, p, blank_bias=b)
        texts, scores = zip(*want.decode_with_scores(x, ln))
        assert ["".join(ob.UNK_GLYPH if i == V.unk_id else V.tokens[i] for i in s) for s in ids] == list(texts)
        assert np.abs(score - np.asarray(scores)).max() <= 1e-4


def test_scale_if(lib):
    x = np.random.default_rng(0).standard_normal(1003).astype(np.float32)
    assert np.array_equal(harness.scale_if(lib, x, 0.65, 0.65), x)  # matches what was expected: untouched
    np.testing.assert_allclose(harness.scale_if(lib, x, 1.3, 0.65), x * np.float32(2.0), rtol=1e-6)


def test_loss_extreme_mismatch_is_redone_in_log_space(lib):
    """Found by the round-2 fuzz: confident posteriors laid out for 333 frames, input cut to ~170 frames, 137-164 labels
    (every frame contradicts the forced alignment: nll ~ 900-1150, thousands of binades inside one strip).  The fp32
    kernel flags such a row, the fp64 kernel cannot hold it either (one exponent per strip) and flags it for the
    log-space pass run by the launch's last CTA; what comes back is the reference's value, not NaN."""
    B, T, S, c = 600, 333, 255, 52
    lp, tg, tl = synth.peaked_aligned(B, T, 7000 + c, n_labels=(1, T // 2))
    tg = torch.nn.functional.pad(tg, (0, S - tg.size(1)))
    il = synth.ragged_lengths(B, T, 9000 + c)
    rows = [8, 31, 5]  # il 169 / tl 137, il 175 / tl 164, and an ordinary neighbour
    sel = torch.tensor(rows)
    assert il[sel].tolist() == [169, 175, 199] and tl[sel].tolist() == [137, 164, 38]
    nll, grad = harness.loss(lib, lp[sel].numpy(), il[sel].numpy(), tg[sel].numpy(), tl[sel].numpy(), 1, True, mean_den=3)
    n64, _, g64 = ol.torch_ctc(lp[sel], il[sel], tg[sel], tl[sel], 1, True, torch.float64)
    assert np.isfinite(grad).all()
    np.testing.assert_allclose(nll, n64.numpy(), rtol=1e-5)
    err = np.abs(grad - g64.numpy()).reshape(3, -1).max(1)
    assert (err <= 1e-5 * np.abs(g64.numpy()).reshape(3, -1).max(1)).all(), err


def test_loss_fast_path_long_classes(lib):
    """Skewed label histograms: the most frequent classes are summed by the whole warp (fast_worker, "long" classes),
    the rest by the per-lane range loop.  One dominant label, a few frequent ones, and a vocabulary of 70 (three
    class passes in two trips, long classes in different passes) -- all must stay on the fp32 path and meet the tolerance."""
    g = torch.Generator().manual_seed(11)
    for B, T, S, Vn, hot in ((3, 300, 120, 48, [(5, 0.6)]), (3, 260, 100, 48, [(7, 0.3), (40, 0.2), (33, 0.15)]),
                             (2, 300, 110, 70, [(3, 0.25), (40, 0.25), (69, 0.2), (36, 0.1)])):
        w = torch.ones(Vn)
        w[0] = w[1] = 0.0
        rest = 1.0 - sum(p for _, p in hot)
        w = w / w.sum() * rest
        for v, p in hot:
            w[v] += p
        tg = torch.multinomial(w, B * S, replacement=True, generator=g).reshape(B, S)
        tl = torch.full((B,), S)
        il = torch.full((B,), T)
        lp = synth.flat(B, T, 100 + S, Vn)
        nll, grad, flags = harness.loss(lib, lp.numpy(), il.numpy(), tg.numpy(), tl.numpy(), 1, True, mean_den=B, with_fallbacks="flags")
        assert flags.tolist() == [0] * B
        n64, _, g64 = ol.torch_ctc(lp, il, tg, tl, 1, True, torch.float64)
        np.testing.assert_allclose(nll, n64.numpy(), rtol=1e-5)
        assert np.abs(grad - g64.numpy()).max() <= 1e-5 * g64.abs().max().item()
