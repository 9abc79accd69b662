"""The distillation term (scope row f-4, the dense-PyTorch half) against the reference module's own outputs
(tests/golden/distill_ref.npz, oracle/make_golden_r2.py).  CPU: `_distill` is plain PyTorch."""
import os

import numpy as np
import pytest
import torch

from silent_speech_decoder_b200.models.losses import DistillationCTCLoss, LossWeights

CASES = ["same_res", "resampled", "resampled_nolen", "normalised"]


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "distill_ref.npz"))
    return {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith(name + ".")}


@pytest.mark.parametrize("name", CASES)
def test_distill_term_matches_reference(golden_dir, name):
    c = load(golden_dir, name)
    fn = DistillationCTCLoss(48, 1, LossWeights(*c["lam"].tolist()), normalize_distill=bool(c["norm"]))
    s = torch.from_numpy(c["student"]).requires_grad_(True)
    t_len = torch.from_numpy(c["teacher_len"]) if c["teacher_len"].size else None
    d = fn._distill(s, torch.from_numpy(c["teacher"]), torch.from_numpy(c["il"]), t_len)
    np.testing.assert_allclose(d.item(), c["distill"], rtol=1e-6)
    (float(c["lam"][0]) * d).backward()
    np.testing.assert_allclose(s.grad.numpy(), c["grad_student"], rtol=1e-5, atol=1e-9)
