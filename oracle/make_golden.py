"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (test infrastructure only).

Run in the build container, where ``/root/reference`` exists:

    python oracle/make_golden.py

Loss and greedy fixtures come from the reference's own code (``src/models/losses.py``,
``src/decoding/ctc.py``, ``src/models/emg_encoder.py``, ``src/models/heads.py``) imported from
``/root/reference`` -- that is what pins ``oracle/ctc_loss.py`` and ``oracle/greedy.py``.
Beam fixtures come from ``oracle/beam.py`` (pyctcdecode is unavailable: PARITY UNPINNED); they
freeze the restatement so that a later change to it cannot silently move the goalposts.
The GPU box has no ``/root/reference``; tests only read the committed ``.npz`` files.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(1, REF)

from silent_speech_decoder_b200 import synth  # noqa: E402
from oracle import beam as obeam  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def _ref_loss(lp, il, tg, tl, blank):
    """The reference module, CTC term only (teacher_repr=None, weights (0,1))."""
    from src.models.losses import DistillationCTCLoss, LossWeights

    fn = DistillationCTCLoss(vocab_size=lp.shape[-1], blank_id=blank, weights=LossWeights(0.0, 1.0))
    x = lp.clone().requires_grad_(True)
    out = fn(x, il, tg, tl, student_repr=x, teacher_repr=None)
    assert set(out) == {"total", "ctc", "distill"}
    out["total"].backward()
    return out["ctc"].detach().numpy(), out["total"].detach().numpy(), x.grad.numpy()


def loss_cases():
    cases = {}
    # (name, B, T, S, generator, ragged, special rows)
    lp = synth.flat(6, 40, 2001)
    tg, tl = synth.targets(6, 12, 2002)
    il = synth.ragged_lengths(6, 40, 2003)
    tg[1, :4] = torch.tensor([7, 7, 7, 9]); tl[1] = 4          # repeats
    tl[2] = 0; tg[2] = 0                                         # empty target
    il[3] = 5; tl[3] = 12                                        # infeasible (S > T)
    il[4] = 40; il[5] = 0; tl[5] = 0; tg[5] = 0                  # zero frames, zero labels
    cases["flat_edges"] = (lp, il, tg, tl, 1)
    lp = synth.peaked(4, 60, 2004)
    tg, tl = synth.targets(4, 10, 2005)
    il = torch.tensor([60, 45, 60, 30])
    cases["peaked"] = (lp, il, tg, tl, 1)
    # reference test shape (tests/test_models.py:26-56): B=2,T=6,V=5, blank 0
    g = torch.Generator().manual_seed(2006)
    lp = torch.log_softmax(torch.randn(2, 6, 5, generator=g), -1)
    cases["test_models_shape"] = (lp, torch.tensor([6, 6]), torch.tensor([[1, 2, 3], [1, 2, 0]]), torch.tensor([3, 2]), 0)
    # tight fit: T == S + repeats
    lp = synth.flat(2, 9, 2007)
    tg = torch.tensor([[5, 5, 6, 6, 7], [2, 3, 4, 5, 6]]); tl = torch.tensor([5, 5]); il = torch.tensor([7, 5])
    cases["tight"] = (lp, il, tg, tl, 1)
    return cases


def edit_pairs():
    """Seeded (reference, hypothesis) sentence pairs: a small word list so hits and ties are common,
    plus the edge cases (empty sides, identical, repeated words, all different)."""
    import random

    rng = random.Random(2301)
    words = ["the", "a", "cat", "sat", "on", "mat", "dog", "ran", "to", "home", "and", "then", "it", "was", "so"]
    pairs = [("", ""), ("", "a cat"), ("the cat sat", ""), ("the cat sat", "the cat sat"), ("a a a a", "a a"),
             ("a b a b a b", "b a b a b a"), ("the  cat   sat ", " the cat sat"), ("x y z", "p q r s")]
    for n in [1, 2, 3, 5, 8, 13, 21, 34, 60]:
        for _ in range(4):
            ref = [rng.choice(words) for _ in range(n)]
            hyp = []
            for w in ref:  # a noisy copy: substitutions, deletions, insertions
                u = rng.random()
                if u < 0.15:
                    hyp.append(rng.choice(words))
                elif u < 0.27:
                    continue
                else:
                    hyp.append(w)
                if rng.random() < 0.12:
                    hyp.append(rng.choice(words))
            pairs.append((" ".join(ref), " ".join(hyp)))
    for _ in range(6):  # unrelated sequences of different lengths
        pairs.append((" ".join(rng.choice(words) for _ in range(rng.randint(1, 40))),
                      " ".join(rng.choice(words) for _ in range(rng.randint(1, 40)))))
    return pairs


def edit_golden() -> None:
    """Word-level alignment counts from the reference's own `compute_error_breakdown`
    (src/evaluation/evaluate.py:41-99).  jiwer is not installed; a stub module without
    `compute_measures` makes the reference take its in-tree Levenshtein branch (:61-99), which is
    the code this fixture pins.  Character-level counts run the same function over character lists
    joined by spaces (one character per "word")."""
    import json
    import types

    sys.modules.setdefault("jiwer", types.ModuleType("jiwer"))
    from src.evaluation.evaluate import compute_error_breakdown

    rows = []
    for ref, hyp in edit_pairs():
        w = compute_error_breakdown([ref], [hyp])
        cr, ch = " ".join(ref.strip().replace(" ", "_")), " ".join(hyp.strip().replace(" ", "_"))
        c = compute_error_breakdown([cr], [ch])
        rows.append({"ref": ref, "hyp": hyp,
                     "word": [int(w[k]) for k in ("substitutions", "deletions", "insertions", "hits")],
                     "char": [int(c[k]) for k in ("substitutions", "deletions", "insertions", "hits")]})
    refs, hyps = [r for r, _ in edit_pairs()], [h for _, h in edit_pairs()]
    total = compute_error_breakdown(refs, hyps)
    with open(os.path.join(OUT, "edit_ref.json"), "w") as f:
        json.dump({"pairs": rows, "corpus_breakdown": total}, f, indent=0)
    print("edit_ref.json", len(rows), "pairs")


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    if sys.argv[1:] == ["edit"]:
        return edit_golden()
    edit_golden()
    from src.decoding.ctc import _greedy_decode, build_greedy_decoder
    from src.data.vocab import Vocab as RefVocab

    rv = RefVocab.from_json(os.path.join(REF, "configs", "vocab.json"))

    # ---- loss: reference module (fp32) + the same ATen op in fp64
    pack = {}
    for name, (lp, il, tg, tl, blank) in loss_cases().items():
        ctc, total, grad = _ref_loss(lp, il, tg, tl, blank)
        x64 = lp.double().clone().requires_grad_(True)
        nll64 = torch.nn.functional.ctc_loss(x64.transpose(0, 1), tg, il, tl, blank=blank, reduction="none", zero_infinity=True)
        (nll64 / tl.clamp_min(1)).mean().backward()
        pack.update({
            f"{name}.lp": lp.numpy(), f"{name}.il": il.numpy(), f"{name}.tg": tg.numpy(), f"{name}.tl": tl.numpy(),
            f"{name}.blank": np.int64(blank), f"{name}.ref_ctc": ctc, f"{name}.ref_total": total, f"{name}.ref_grad": grad,
            f"{name}.nll64": nll64.detach().numpy(), f"{name}.grad64": x64.grad.numpy(),
        })
    np.savez_compressed(os.path.join(OUT, "loss_ref.npz"), **pack)

    # ---- greedy: reference function, KAT of tests/test_evaluation.py:6-14, bias, ragged
    pack = {}
    kat = torch.nn.functional.one_hot(torch.tensor([[1, 1, 2, 2, 2], [0, 1, 1, 0, 2]]), num_classes=3).float()
    got = _greedy_decode(kat, torch.tensor([5, 5]), blank_id=0)
    assert got == [[1, 2], [1, 2]], got
    pack["kat.lp"] = kat.numpy(); pack["kat.len"] = np.array([5, 5]); pack["kat.blank"] = np.int64(0)
    pack["kat.ids"] = np.array(got, dtype=np.int64)

    def ragged(lists):
        n = max((len(x) for x in lists), default=0)
        arr = np.full((len(lists), max(n, 1)), -1, dtype=np.int64)
        for i, x in enumerate(lists):
            arr[i, : len(x)] = x
        return arr

    for name, lp, ln, bias in [
        ("flat", synth.flat(5, 64, 2101), torch.tensor([64, 1, 0, 33, 64]), 0.0),
        ("peaked", synth.peaked(5, 120, 2102), torch.tensor([120, 77, 120, 5, 99]), 0.0),
        ("peaked_bias", synth.peaked(4, 90, 2103), torch.tensor([90, 90, 45, 90]), 2.5),
        ("flat_negbias", synth.flat(3, 50, 2104), torch.tensor([50, 50, 50]), -0.7),
    ]:
        ids = _greedy_decode(lp, ln, blank_id=rv.blank_id, blank_bias=bias)
        txt = build_greedy_decoder(rv, blank_bias=bias)(lp, ln)
        pack[f"{name}.lp"] = lp.numpy(); pack[f"{name}.len"] = ln.numpy(); pack[f"{name}.bias"] = np.float64(bias)
        pack[f"{name}.ids"] = ragged(ids); pack[f"{name}.text"] = np.array(txt, dtype=object).astype("U")
    np.savez_compressed(os.path.join(OUT, "greedy_ref.npz"), **pack)

    # ---- cfg1: the reference's own encoder + head, random init (SURVEY 8d cfg 1)
    from src.models.emg_encoder import EMGConformerEncoder, EncoderConfig
    from src.models.heads import CTCHead

    torch.manual_seed(1001)
    enc = EMGConformerEncoder(EncoderConfig(input_dim=640, d_model=256, num_layers=6, num_heads=4, ffn_dim=512,
                                            depthwise_conv_kernel_size=15, dropout=0.1, subsample_factor=2)).eval()
    head = CTCHead(256, 48).eval()
    x = torch.randn(8, 1000, 640)
    lens = torch.tensor([1000, 900, 800, 1000, 640, 999, 501, 1000])
    with torch.no_grad():
        h, hl = enc(x, lens)
        lp = head(h).contiguous()
    tg, tl = synth.targets(8, 60, 1001)
    ctc, total, grad = _ref_loss(lp, hl, tg, tl, rv.blank_id)
    x64 = lp.double().clone().requires_grad_(True)
    nll64 = torch.nn.functional.ctc_loss(x64.transpose(0, 1), tg, hl, tl, blank=rv.blank_id, reduction="none", zero_infinity=True)
    (nll64 / tl.clamp_min(1)).mean().backward()
    ids = _greedy_decode(lp, hl, blank_id=rv.blank_id)
    txt = build_greedy_decoder(rv)(lp, hl)
    bd = obeam.BeamDecoder(rv.tokens, rv.blank_id, rv.pad_id, beam_width=50, beam_prune_logp=-10.0)
    bout = bd.decode_with_scores(lp.numpy(), hl.tolist())
    np.savez_compressed(
        os.path.join(OUT, "cfg1_ref.npz"),
        lp=lp.numpy(), il=hl.numpy(), tg=tg.numpy(), tl=tl.numpy(), ref_ctc=ctc, ref_grad=grad,
        nll64=nll64.detach().numpy(), greedy_ids=ragged(ids),
        greedy_text=np.array(txt, dtype=object).astype("U"),
        beam_text=np.array([t for t, _ in bout], dtype=object).astype("U"), beam_score=np.array([s for _, s in bout]),
    )

    # ---- beam: frozen outputs of the restatement (UNPINNED against pyctcdecode)
    pack = {}
    for name, lp, ln, width, prune, bias in [
        ("peaked50", synth.peaked(6, 120, 2201), [120, 100, 120, 61, 7, 0], 50, -10.0, 0.0),
        ("flat50", synth.flat(4, 60, 2202), [60, 60, 31, 60], 50, -10.0, 0.0),
        ("flat20_bias", synth.flat(3, 50, 2203), [50, 50, 50], 20, -10.0, 0.12),
        ("peaked100_p12", synth.peaked(3, 100, 2204), [100, 100, 100], 100, -12.0, 0.05),
        ("flat1", synth.flat(2, 40, 2205), [40, 40], 1, -10.0, 0.0),
        ("flat200", synth.flat(2, 40, 2206), [40, 40], 200, -10.0, 0.0),
    ]:
        bd = obeam.BeamDecoder(rv.tokens, rv.blank_id, rv.pad_id, beam_width=width, beam_prune_logp=prune, blank_bias=bias)
        out = bd.decode_with_scores(lp.numpy(), ln)
        pack[f"{name}.lp"] = lp.numpy(); pack[f"{name}.len"] = np.array(ln)
        pack[f"{name}.cfg"] = np.array([width, prune, bias], dtype=np.float64)
        pack[f"{name}.text"] = np.array([t for t, _ in out], dtype=object).astype("U")
        pack[f"{name}.score"] = np.array([s for _, s in out])
    np.savez_compressed(os.path.join(OUT, "beam_restatement.npz"), **pack)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
