"""The reference's call sites, re-enacted under ``install_as_src()``: the imports are the reference's own
(``from src.models.losses import ...``, ``from src.decoding.ctc import ...``, ``from src.data.vocab import ...``) and the
calls have the shapes and keyword arguments of its tests and loops (tests/test_models.py:40-56, tests/test_training.py:60-98,
tests/test_evaluation.py:6-14, src/training/train.py:237-247, src/evaluation/evaluate.py:299).  Values are checked against
the CPU oracle; the GPU box has no reference checkout, so the surrounding reference modules are the package's mirrors."""
import sys

import pytest
import torch

from oracle import ctc_loss as ol

pytestmark = pytest.mark.gpu


@pytest.fixture()
def as_src():
    import silent_speech_decoder_b200 as ssd

    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    ssd.install_as_src()
    yield
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    sys.modules.update(saved)


def test_reference_test_models_call(as_src):
    from src.models.losses import DistillationCTCLoss, LossWeights  # the reference's import line

    torch.manual_seed(0)
    batch, time, vocab_size, dim = 2, 50, 10, 8
    log_probs = torch.log_softmax(torch.randn(batch, time, vocab_size), dim=-1).cuda().requires_grad_(True)
    proj_out = torch.randn(batch, time, dim, device="cuda", requires_grad=True)
    targets = torch.tensor([[1, 2, 3], [1, 2, 0]], dtype=torch.long)
    target_lengths = torch.tensor([3, 2], dtype=torch.long)
    logit_lengths = torch.tensor([time, time], dtype=torch.long)
    teacher = torch.randn(batch, time, dim, device="cuda")
    loss_fn = DistillationCTCLoss(vocab_size=vocab_size, blank_id=0, weights=LossWeights())
    losses = loss_fn(log_probs=log_probs, logit_lengths=logit_lengths, targets=targets, target_lengths=target_lengths,
                     student_repr=proj_out, teacher_repr=teacher)
    assert set(losses.keys()) == {"total", "ctc", "distill"}
    assert losses["total"].item() >= 0
    losses["total"].backward()  # test_training.py:97 / train.py:247
    _, mean64, g64 = ol.torch_ctc(log_probs.detach().cpu(), logit_lengths, targets, target_lengths, 0, True, torch.float64)
    assert abs(losses["ctc"].item() - float(mean64)) <= 1e-5 * float(mean64)
    lam = LossWeights().lambda_ctc
    assert (log_probs.grad.cpu().double() - lam * g64).abs().max() <= 2e-5 * (lam * g64).abs().max()
    assert proj_out.grad is not None


def test_reference_decoder_calls(as_src, golden_dir):
    import os

    import numpy as np
    from src.data.vocab import Vocab
    from src.decoding.ctc import _greedy_decode, build_decoder

    # tests/test_evaluation.py:6-14, through the name the decoding module exports
    log_probs = torch.nn.functional.one_hot(torch.tensor([[1, 1, 2, 2, 2], [0, 1, 1, 0, 2]]), num_classes=3).float()
    assert _greedy_decode(log_probs, torch.tensor([5, 5]), blank_id=0) == [[1, 2], [1, 2]]
    # evaluate.py:212,299: build_decoder(...) then decoder(log_probs.cpu(), enc_lengths.cpu())
    z = np.load(os.path.join(golden_dir, "cfg1_ref.npz"))
    vocab = Vocab.from_json(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                         "silent-speech-decoder_b200", "configs", "vocab.json"))
    for method, want in (("greedy", z["greedy_text"]), ("beam", z["beam_text"])):
        decoder = build_decoder(method=method, vocab=vocab, lm_path=None, beam_width=50, alpha=0.6, beta=0.0,
                                beam_prune_logp=-10.0, blank_bias=0.0)
        hyps = decoder(torch.from_numpy(z["lp"]).cpu(), torch.from_numpy(z["il"]).cpu())
        assert hyps == list(want)
