"""Pins oracle/ctc_loss.py against outputs of the reference's own DistillationCTCLoss
(tests/golden/loss_ref.npz, cfg1_ref.npz -- produced by oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import ctc_loss as ol

CASES = ["flat_edges", "peaked", "test_models_shape", "tight"]


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "loss_ref.npz"))
    return {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith(name + ".")}


@pytest.mark.parametrize("name", CASES)
def test_torch_oracle_reproduces_reference_module(golden_dir, name):
    c = _load(golden_dir, name)
    loss, grad = ol.reference_module_ctc(torch.from_numpy(c["lp"]), torch.from_numpy(c["il"]), torch.from_numpy(c["tg"]),
                                         torch.from_numpy(c["tl"]), int(c["blank"]))
    # same ATen kernel, same dtype -> bit-identical
    assert loss.item() == float(c["ref_ctc"])
    assert np.array_equal(grad.numpy(), c["ref_grad"])


@pytest.mark.parametrize("name", CASES)
def test_numpy_restatement_matches_reference(golden_dir, name):
    c = _load(golden_dir, name)
    nll, mean, grad = ol.numpy_ctc(c["lp"], c["il"], c["tg"], c["tl"], int(c["blank"]), zero_infinity=True)
    # vs the fp64 run of the same ATen op: tight
    np.testing.assert_allclose(nll, c["nll64"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(grad, c["grad64"], rtol=0, atol=1e-13)
    # vs the reference module (fp32): its own rounding is the limit
    assert abs(mean - float(c["ref_ctc"])) <= 2e-6 * max(1.0, abs(mean))
    # (already 1e-5..1e-4 normwise at T=40..60: fp32 log-space rounding inside ATen)
    assert np.abs(grad - c["ref_grad"]).max() <= 1e-3 * max(np.abs(grad).max(), 1e-30) + 1e-8


def test_edge_semantics(golden_dir):
    c = _load(golden_dir, "flat_edges")
    nll = c["nll64"]
    lp, il, blank = c["lp"].astype(np.float64), c["il"], int(c["blank"])
    # empty target: only the all-blank path
    assert np.isclose(nll[2], -lp[2, : il[2], blank].sum())
    # infeasible row and the T=0,S=0 row: zero loss, zero gradient, still in the mean
    assert nll[3] == 0.0 and not c["grad64"][3].any()
    assert nll[5] == 0.0 and not c["grad64"][5].any()
    # rows at or beyond the input length carry exactly zero gradient
    for b in range(lp.shape[0]):
        assert not c["ref_grad"][b, il[b]:].any()
    mean = float(np.mean(nll / np.maximum(c["tl"], 1)))
    assert abs(mean - float(c["ref_ctc"])) < 1e-5


def test_cfg1_fixture(golden_dir):
    z = np.load(os.path.join(golden_dir, "cfg1_ref.npz"))
    nll, mean, grad = ol.torch_ctc(torch.from_numpy(z["lp"]), torch.from_numpy(z["il"]), torch.from_numpy(z["tg"]),
                                   torch.from_numpy(z["tl"]), blank=1, dtype=torch.float64)
    np.testing.assert_allclose(nll.numpy(), z["nll64"], rtol=1e-12)
    assert abs(float(mean) - float(z["ref_ctc"])) <= 1e-5 * abs(float(mean))
    # the reference's fp32 gradient is itself ~1e-4 (normwise) away from fp64 at T=500
    gmax = grad.abs().max().item()
    assert np.abs(grad.numpy() - z["ref_grad"]).max() <= 5e-3 * gmax
