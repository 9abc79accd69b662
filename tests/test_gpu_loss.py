"""GPU parity tests for the fused CTC loss kernel through the reference-shaped modules.

Tolerances.  north_star asks for 1e-5 relative on loss and gradients.  The reference's own fp32
ATen path is NOT that accurate against exact arithmetic (gradient off by ~1e-4 normwise at T=60,
~3e-3 at T=1000: fp32 log-space rounding at |alpha| ~ 1e3), so the 1e-5 bound is asserted against
the SAME ATen op run in fp64 (oracle.ctc_loss.torch_ctc), and the native result is additionally
required to be at least as close to fp64 as the reference's fp32 result is."""
import os

import numpy as np
import pytest
import torch

from oracle import ctc_loss as ol
from silent_speech_decoder_b200 import ops, synth
from silent_speech_decoder_b200.models.losses import CTCLoss, DistillationCTCLoss, LossWeights, ctc_loss

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _native_mean(lp, il, tg, tl, blank=1):
    x = lp.cuda().requires_grad_(True)
    fn = DistillationCTCLoss(vocab_size=lp.shape[-1], blank_id=blank, weights=LossWeights(0.0, 1.0))
    out = fn(x, il, tg.cuda(), tl.cuda(), student_repr=x, teacher_repr=None)
    out["total"].backward()
    return out["ctc"].item(), x.grad.cpu()


def _check(lp, il, tg, tl, blank=1):
    nll64, mean64, g64 = ol.torch_ctc(lp, il, tg, tl, blank, True, torch.float64)
    mean, grad = _native_mean(lp, il, tg, tl, blank)
    assert abs(mean - float(mean64)) <= RTOL * max(abs(float(mean64)), 1e-6)
    gmax = g64.abs().max().item()
    err = (grad.double() - g64).abs().max().item()
    assert err <= RTOL * gmax + 1e-12, (err, gmax)
    assert not grad.isnan().any()
    # rows past the input length are exactly zero
    for b in range(lp.shape[0]):
        assert not grad[b, int(il[b]):].any()
    return err / max(gmax, 1e-30)


@pytest.mark.parametrize("name", ["flat_edges", "peaked", "test_models_shape", "tight"])
def test_golden_reference_cases(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "loss_ref.npz"))
    c = {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith(name + ".")}
    lp, il, tg, tl = (torch.from_numpy(c[k]) for k in ("lp", "il", "tg", "tl"))
    mean, grad = _native_mean(lp, il, tg, tl, int(c["blank"]))
    # against the reference module's own (fp32) numbers: its rounding is the limit
    assert abs(mean - float(c["ref_ctc"])) <= 1e-5 * max(abs(mean), 1e-6) + 1e-6
    assert np.abs(grad.numpy() - c["ref_grad"]).max() <= 1e-3 * np.abs(c["ref_grad"]).max() + 1e-9
    # against the same op in fp64: the north_star tolerance
    assert np.abs(grad.numpy() - c["grad64"]).max() <= RTOL * np.abs(c["grad64"]).max() + 1e-12
    nll, _ = ops.ctc_loss_fwd_bwd(lp.cuda(), il, tg, tl, blank=int(c["blank"]), want_grad=False)
    np.testing.assert_allclose(nll.cpu().numpy(), c["nll64"], rtol=RTOL, atol=1e-6)


def test_cfg1_reference_encoder_logits(golden_dir):
    z = np.load(os.path.join(golden_dir, "cfg1_ref.npz"))
    lp, il, tg, tl = (torch.from_numpy(z[k]) for k in ("lp", "il", "tg", "tl"))
    rel = _check(lp, il, tg, tl)
    _, _, g64 = ol.torch_ctc(lp, il, tg, tl, 1, True, torch.float64)
    ref_rel = np.abs(z["ref_grad"] - g64.numpy()).max() / g64.abs().max().item()
    assert rel <= ref_rel  # closer to exact than the reference's fp32 path


@pytest.mark.parametrize("gen,B,T,S", [("flat", 16, 1000, 200), ("peaked", 16, 1000, 200), ("flat", 64, 500, 60),
                                        ("flat", 3, 2000, 500), ("peaked", 5, 700, 300), ("flat", 33, 50, 3)])
def test_vs_fp64_oracle(gen, B, T, S):
    lp = getattr(synth, gen)(B, T, 1002 + B + T)
    tg, tl = synth.targets(B, S, 1002 + S)
    il = synth.ragged_lengths(B, T, 5)
    il[0] = T
    tl[0] = S
    _check(lp, il, tg, tl)


def test_fast_path_and_exact_fallback(monkeypatch):
    """ssd_ctc_loss = fp32 fast kernel + fp64 kernel over what the fast one flags.  Plain batches stay on the fast
    path; confident-but-wrong posteriors lose fp32 mass, are caught by the per-frame mass check and come back
    from the exact kernel; both meet the same tolerance.  SSD_LOSS_EXACT=1 forces the fp64 kernel alone."""
    lp = synth.flat(32, 400, 11)
    tg, tl = synth.targets(32, 100, 12)
    il = synth.ragged_lengths(32, 400, 13)
    _check(lp, il, tg, tl)
    assert ops.loss_fallback_count(32, 400, 48, 100) == 0
    lp = synth.peaked(8, 700, 77, boost=10.0)
    tg, tl = synth.targets(8, 120, 18)
    tl[:] = 120
    il = torch.full((8,), 700)
    _check(lp, il, tg, tl)
    assert ops.loss_fallback_count(8, 700, 48, 120) > 0
    # same inputs, exact kernel only: identical answers within the tolerance of either path
    nll_a, g_a = ops.ctc_loss_fwd_bwd(lp.cuda(), il, tg, tl, blank=1, mean_denominator=8)
    monkeypatch.setenv("SSD_LOSS_EXACT", "1")
    nll_b, g_b = ops.ctc_loss_fwd_bwd(lp.cuda(), il, tg, tl, blank=1, mean_denominator=8)
    assert ops.loss_fallback_count(8, 700, 48, 120) is None
    assert ((nll_a - nll_b).abs() / nll_b.abs()).max().item() < 1e-6
    assert (g_a - g_b).abs().max().item() <= 1e-5 * g_b.abs().max().item()


def test_fast_path_flags_out_of_range_inputs():
    lp = synth.flat(4, 64, 21).clone()
    tg, tl = synth.targets(4, 10, 22)
    il = torch.tensor([64, 64, 0, 7])
    tl[3] = 10                      # infeasible: 10 labels in 7 frames
    lp[1, 5, 3] = -200.0            # exp underflows in fp32
    _check(lp, il, tg, tl)
    assert ops.loss_fallback_count(4, 64, 48, 10) == 3


def test_edge_rows_and_flags():
    lp = synth.flat(6, 30, 77)
    tg, tl = synth.targets(6, 12, 78)
    il = torch.tensor([30, 30, 5, 0, 30, 1])
    tl[1] = 0                      # empty target
    tl[2] = 12                     # infeasible: S > T
    tl[3] = 0                      # no frames, no labels
    tg[4, :3] = 9; tl[4] = 3       # repeats need separating blanks
    tl[5] = 1
    _check(lp, il, tg, tl)
    x = lp.cuda().transpose(0, 1)
    none = ctc_loss(x, tg, il, tl, blank=1, reduction="none", zero_infinity=False).cpu()
    assert torch.isinf(none[2]) and none[3] == 0
    want = torch.nn.functional.ctc_loss(lp.transpose(0, 1).double(), tg, il, tl, blank=1, reduction="none", zero_infinity=True)
    np.testing.assert_allclose(ctc_loss(x, tg, il, tl, 1, "none", True).cpu().numpy(), want.numpy(), rtol=RTOL, atol=1e-6)
    s = ctc_loss(x, tg, il, tl, 1, "sum", True).item()
    assert abs(s - want.sum().item()) <= RTOL * abs(want.sum().item())


def test_nn_ctcloss_call_shape_and_errors():
    lp = synth.flat(4, 40, 5)
    tg, tl = synth.targets(4, 8, 6)
    il = torch.tensor([40, 33, 40, 21])
    x = lp.cuda().transpose(0, 1).contiguous().requires_grad_(True)  # (T,B,V) contiguous, as a bare nn.CTCLoss user has it
    loss = CTCLoss(blank=1, zero_infinity=True)(x, tg, il, tl)
    loss.backward()
    _, mean64, g64 = ol.torch_ctc(lp, il, tg, tl, 1, True, torch.float64)
    assert abs(loss.item() - float(mean64)) <= RTOL * float(mean64)
    assert (x.grad.cpu().transpose(0, 1).double() - g64).abs().max() <= RTOL * g64.abs().max()
    with pytest.raises(RuntimeError):
        CTCLoss(blank=1)(x, tg, torch.tensor([41, 1, 1, 1]), tl)       # input length beyond T
    with pytest.raises(RuntimeError):
        CTCLoss(blank=99)(x, tg, il, tl)                               # blank outside the vocabulary
    with pytest.raises(RuntimeError):
        CTCLoss(blank=1)(x.half(), tg, il, tl)


def test_distill_term_and_weights():
    torch.manual_seed(0)
    B, T, D = 3, 20, 16
    lp = synth.flat(B, T, 9).cuda().requires_grad_(True)
    tg, tl = synth.targets(B, 5, 10)
    il = torch.tensor([20, 15, 20])
    student = torch.randn(B, T, D, device="cuda", requires_grad=True)
    teacher = torch.randn(B, 50, D, device="cuda")
    fn = DistillationCTCLoss(48, 1, LossWeights(0.7, 0.3), normalize_distill=True)
    out = fn(lp, il.cuda(), tg.cuda(), tl.cuda(), student, teacher, teacher_lengths=torch.tensor([50, 30, 44]))
    assert set(out) == {"total", "ctc", "distill"} and out["total"].item() >= 0
    assert abs(out["total"].item() - (0.3 * out["ctc"].item() + 0.7 * out["distill"].item())) < 1e-5
    out["total"].backward()
    _, _, g64 = ol.torch_ctc(lp.detach().cpu(), il, tg, tl, 1, True, torch.float64)
    assert (lp.grad.cpu().double() - 0.3 * g64).abs().max() <= 2 * RTOL * (0.3 * g64).abs().max()
    assert student.grad is not None and student.grad.abs().sum() > 0


def test_full_size_cfg2_properties():
    # config 2: B=256, T=1000, V=48, S up to 200.  Size-independent checks:
    #  * posteriors are a distribution: sum_v (exp(lp) - grad/scale) == 1 for every live frame
    #  * nll agrees with ATen's CUDA kernel (fp32 log-space, so only ~1e-6 relative)
    B, T, S = 256, 1000, 200
    lp = synth.flat(B, T, 1002).cuda()
    tg, tl = synth.targets(B, S, 1002)
    il = torch.full((B,), T)
    nll, grad = ops.ctc_loss_fwd_bwd(lp, il, tg, tl, blank=1)
    assert torch.isfinite(nll).all() and torch.isfinite(grad).all()
    occ = lp.exp() - grad
    assert (occ.sum(-1) - 1).abs().max().item() < 5e-5
    assert occ.min().item() > -1e-5
    ref = torch.nn.functional.ctc_loss(lp.transpose(0, 1), tg.cuda(), il, tl, blank=1, reduction="none", zero_infinity=True)
    assert ((nll - ref).abs() / ref.abs()).max().item() < 5e-6
    # determinism
    nll2, grad2 = ops.ctc_loss_fwd_bwd(lp, il, tg, tl, blank=1)
    assert torch.equal(nll, nll2) and torch.equal(grad, grad2)


@pytest.mark.timeout(180)
def test_more_utterances_than_ctas(monkeypatch):
    """700 ragged utterances on at most 296 CTAs: every CTA walks through several utterances, its late worker
    (one of the sweep warps, once the sweeps are done) re-using the sweeps' ring space each time.  Checked against
    the exact fp64 kernel of the same library, and against itself on fewer CTAs and without the late worker."""
    B, T, S = 700, 160, 40
    lp = synth.flat(B, T, 71).cuda()
    tg, tl = synth.targets(B, S, 72)
    il = synth.ragged_lengths(B, T, 73)
    nll, g = ops.ctc_loss_fwd_bwd(lp, il, tg, tl, blank=1, mean_denominator=B)
    assert ops.loss_fallback_count(B, T, 48, S) == 0
    monkeypatch.setenv("SSD_LOSS_EXACT", "1")
    nll_x, g_x = ops.ctc_loss_fwd_bwd(lp, il, tg, tl, blank=1, mean_denominator=B)
    monkeypatch.delenv("SSD_LOSS_EXACT")
    assert ((nll - nll_x).abs() / nll_x.abs()).max().item() < 2e-6
    assert (g - g_x).abs().max().item() <= 1e-5 * g_x.abs().max().item()
    # fewer CTAs (some alone on their SM, up to 19 utterances each; one CTA for a short prefix) and no late worker:
    # who computes a block, and when, does not change a single bit.  The capped grids are the regression test for
    # the per-utterance mbarrier re-initialisation that used to leave workers asleep in exactly this regime.
    for env, n in (({"SSD_LOSS_GRID": "148"}, B), ({"SSD_LOSS_GRID": "37"}, B), ({"SSD_LOSS_NOLATE": "1"}, B),
                   ({"SSD_LOSS_GRID": "1"}, 9), ({"SSD_LOSS_GRID": "1", "SSD_LOSS_NOLATE": "1"}, 9)):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        nll_b, g_b = ops.ctc_loss_fwd_bwd(lp[:n], il[:n], tg[:n], tl[:n], blank=1, mean_denominator=B)
        for k in env:
            monkeypatch.delenv(k)
        assert torch.equal(nll_b, nll[:n]) and torch.equal(g_b, g[:n]), env


@pytest.mark.timeout(300)
def test_random_shapes_against_exact_kernel():
    """profiles/fuzz_loss.py: random (B, T, S, input kind, grid cap, late worker) against the exact fp64 kernel of the same
    library, each case under a watchdog; exits 2 on a tolerance miss, 3 on a hang."""
    import subprocess
    import sys

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(repo, "profiles", "fuzz_loss.py"), "40", "5"], capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "all 40 cases ok" in r.stdout


@pytest.mark.timeout(300)
def test_full_size_cfg2_gradient_against_fp64_oracle():
    """Config 2 at FULL size (B=256, T=1000, S~U[50,200]): nll and the whole (B,T,V) gradient against the reference's
    ATen op evaluated in fp64 on the host (oracle.ctc_loss.torch_ctc), to the north_star tolerance."""
    B, T, S = 256, 1000, 200
    lp = synth.flat(B, T, 1002)
    tg, tl = synth.targets(B, S, 1002)
    il = torch.full((B,), T)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    nll64, mean64, g64 = ol.torch_ctc(lp, il, tg, tl, 1, True, torch.float64)
    x = lp.cuda().requires_grad_(True)
    out = DistillationCTCLoss(48, 1, LossWeights(0.0, 1.0))(x, il, tg.cuda(), tl.cuda(), x, None)
    out["total"].backward()
    assert abs(out["ctc"].item() - float(mean64)) <= RTOL * float(mean64)
    err = (x.grad.cpu().double() - g64).abs().amax(dim=(1, 2))
    ref = g64.abs().amax(dim=(1, 2))
    assert (err <= RTOL * ref).all(), (err / ref).max().item()   # per utterance, not only over the batch
    assert ops.loss_fallback_count(B, T, 48, S) == 0


@pytest.mark.parametrize("name", ["same_res", "resampled", "resampled_nolen", "normalised"])
def test_module_with_teacher_matches_reference(golden_dir, name):
    """total / ctc / distill and both gradients of the reference's DistillationCTCLoss run WITH a teacher
    (tests/golden/distill_ref.npz).  lambda_ctc != 1 also exercises the up-front gradient scale (expect_grad)."""
    z = np.load(os.path.join(golden_dir, "distill_ref.npz"))
    c = {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith(name + ".")}
    fn = DistillationCTCLoss(48, 1, LossWeights(*c["lam"].tolist()), normalize_distill=bool(c["norm"]))
    x = torch.from_numpy(c["lp"]).cuda().requires_grad_(True)
    s = torch.from_numpy(c["student"]).cuda().requires_grad_(True)
    t_len = torch.from_numpy(c["teacher_len"]).cuda() if c["teacher_len"].size else None
    out = fn(x, torch.from_numpy(c["il"]).cuda(), torch.from_numpy(c["tg"]).cuda(), torch.from_numpy(c["tl"]).cuda(), s,
             torch.from_numpy(c["teacher"]).cuda(), t_len)
    out["total"].backward()
    for k in ("total", "ctc", "distill"):
        np.testing.assert_allclose(out[k].item(), c[k], rtol=2e-6, err_msg=k)
    np.testing.assert_allclose(s.grad.cpu().numpy(), c["grad_student"], rtol=1e-5, atol=1e-9)
    g = c["grad_lp"]
    assert np.abs(x.grad.cpu().numpy() - g).max() <= 2e-5 * np.abs(g).max()   # the reference's fp32 gradient is the looser side


def test_unexpected_upstream_gradient_is_applied():
    """The kernel pre-scales by the scalar it expects from autograd; a different one must still give the right
    gradient (ssd_scale_if's slow branch), also when backward runs twice through the same graph."""
    B, T = 4, 50
    lp = synth.flat(B, T, 31)
    tg, tl = synth.targets(B, 8, 32)
    il = torch.tensor([50, 41, 50, 33])
    _, _, g64 = ol.torch_ctc(lp, il, tg, tl, 1, True, torch.float64)
    x = lp.cuda().requires_grad_(True)
    fn = DistillationCTCLoss(48, 1, LossWeights(0.0, 0.65))
    out = fn(x, il, tg.cuda(), tl.cuda(), x, None)
    (3.0 * out["total"]).backward(retain_graph=True)          # upstream 3 * 0.65, the kernel expected 0.65
    assert (x.grad.cpu().double() - 1.95 * g64).abs().max() <= 2 * RTOL * 1.95 * g64.abs().max()
    x.grad = None
    out["ctc"].backward()                                       # second pass, upstream 1.0
    assert (x.grad.cpu().double() - g64).abs().max() <= 2 * RTOL * g64.abs().max()
    # reductions 'sum' and 'none' through the functional form
    for red in ("sum", "none"):
        y = lp.cuda().transpose(0, 1).detach().requires_grad_(True)
        v = ctc_loss(y, tg.cuda(), il, tl, blank=1, reduction=red, zero_infinity=True)
        w = torch.tensor([1.0, 2.0, 0.5, 3.0], device="cuda")
        (v * w).sum().backward() if red == "none" else v.backward()
        y64 = lp.double().transpose(0, 1).detach().requires_grad_(True)
        r = torch.nn.functional.ctc_loss(y64, tg, il, tl, blank=1, reduction=red, zero_infinity=True)
        (r * w.cpu().double()).sum().backward() if red == "none" else r.backward()
        assert (y.grad.cpu().double() - y64.grad).abs().max() <= 2 * RTOL * y64.grad.abs().max(), red


@pytest.mark.parametrize("exact", [False, True])
def test_loss_from_logits(golden_dir, exact, monkeypatch):
    """Scope row f-2 on the reference's own head output: cfg1's fixture holds CTCHead log-probs of the random-init
    mps_fast_basic encoder; logits = those log-probs plus an arbitrary per-frame constant (what log_softmax removes).
    Loss and d/d(logits) against head + reference loss in fp64."""
    if exact:
        monkeypatch.setenv("SSD_LOSS_EXACT", "1")
    z = np.load(os.path.join(golden_dir, "cfg1_ref.npz"))
    lp, il, tg, tl = (torch.from_numpy(z[k]) for k in ("lp", "il", "tg", "tl"))
    g = torch.Generator().manual_seed(5)
    logits = lp + 7.0 * torch.randn(lp.shape[0], lp.shape[1], 1, generator=g)
    x64 = logits.double().requires_grad_(True)
    ref = torch.nn.functional.ctc_loss(torch.log_softmax(x64, -1).transpose(0, 1), tg, il, tl, blank=1, reduction="mean", zero_infinity=True)
    ref.backward()
    x = logits.cuda().requires_grad_(True)
    out = DistillationCTCLoss(48, 1, LossWeights(0.0, 1.0), from_logits=True)(x, il, tg.cuda(), tl.cuda(), x, None)
    out["total"].backward()
    assert abs(out["ctc"].item() - ref.item()) <= RTOL * ref.item()
    assert (x.grad.cpu().double() - x64.grad).abs().max() <= RTOL * x64.grad.abs().max()
    np.testing.assert_allclose(ops.row_logsumexp(logits.cuda()).cpu().numpy(), torch.logsumexp(logits.double(), -1).numpy(), atol=2e-6, rtol=0)
    # and at config-2 size the logits entry and the log-prob entry agree with each other
    if not exact:
        B, T, S = 256, 1000, 200
        big = torch.randn(B, T, 48, generator=g).cuda()
        tg2, tl2 = synth.targets(B, S, 1002)
        il2 = torch.full((B,), T)
        n1, g1 = ops.ctc_loss_fwd_bwd(big, il2, tg2, tl2, blank=1, from_logits=True)
        n2, g2 = ops.ctc_loss_fwd_bwd(torch.log_softmax(big, -1), il2, tg2, tl2, blank=1)
        assert ((n1 - n2).abs() / n2.abs()).max().item() < 2e-6
        assert (g1 - g2).abs().max().item() <= RTOL * g2.abs().max().item()


def test_two_streams_do_not_share_scratch():
    """Per-(device, stream) workspaces: interleaved launches on two streams give the single-stream results."""
    B, T, S = 64, 300, 40
    a, b = synth.flat(B, T, 91).cuda(), synth.flat(B, T, 92).cuda()
    tg, tl = synth.targets(B, S, 93)
    il = torch.full((B,), T)
    ra, rb = ops.ctc_loss_fwd_bwd(a, il, tg, tl, blank=1), ops.ctc_loss_fwd_bwd(b, il, tg, tl, blank=1)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    tgd, tld, ild = tg.cuda(), tl.cuda(), il.cuda()
    outs = []
    for _ in range(3):
        with torch.cuda.stream(s1):
            o1 = ops.ctc_loss_fwd_bwd(a, ild, tgd, tld, blank=1)
        with torch.cuda.stream(s2):
            o2 = ops.ctc_loss_fwd_bwd(b, ild, tgd, tld, blank=1)
        outs.append((o1, o2))
    torch.cuda.synchronize()
    for o1, o2 in outs:
        assert torch.equal(o1[0], ra[0]) and torch.equal(o1[1], ra[1])
        assert torch.equal(o2[0], rb[0]) and torch.equal(o2[1], rb[1])


@pytest.mark.timeout(240)
def test_per_utterance_strip_widths(monkeypatch):
    """One kernel per strip width (prepass + programmatic dependent launches; the default above 4 utterances per SM,
    forced here for a small batch too): every class of target length in one batch, ragged inputs, rows the fast path
    refuses.  Against the exact fp64 kernel, and the two schemes against each other."""
    B, T, S = 900, 200, 250
    g = torch.Generator().manual_seed(11)
    tl = torch.randint(1, S + 1, (B,), generator=g)
    il = synth.ragged_lengths(B, T, 12)
    tg = torch.randint(2, 48, (B, S), generator=g) * (torch.arange(S)[None, :] < tl[:, None])
    il[3], tl[4] = 0, 0
    lp = synth.flat(B, T, 13).cuda()
    monkeypatch.setenv("SSD_LOSS_EXACT", "1")
    nll_x, g_x = ops.ctc_loss_fwd_bwd(lp, il, tg, tl, blank=1, mean_denominator=B)
    monkeypatch.delenv("SSD_LOSS_EXACT")
    fin = torch.isfinite(nll_x) & (nll_x != 0)
    outs = {}
    for multi, n in (("1", B), ("0", B), ("1", 40), ("0", 40)):
        monkeypatch.setenv("SSD_LOSS_MULTI", multi)
        nll, gr = ops.ctc_loss_fwd_bwd(lp[:n], il[:n], tg[:n], tl[:n], blank=1, mean_denominator=B)
        f = fin[:n]
        assert ((nll[f] - nll_x[:n][f]).abs() / nll_x[:n][f].abs()).max().item() < 5e-6, (multi, n)
        assert (gr - g_x[:n]).abs().max().item() <= RTOL * g_x[:n].abs().max().item(), (multi, n)
        outs[(multi, n)] = (nll, gr)
    monkeypatch.delenv("SSD_LOSS_MULTI")
    # the default for this batch (900 > 4 x 148) is the per-width scheme: same bits as forcing it
    nll, gr = ops.ctc_loss_fwd_bwd(lp, il, tg, tl, blank=1, mean_denominator=B)
    assert torch.equal(nll, outs[("1", B)][0]) and torch.equal(gr, outs[("1", B)][1])


def test_device_resident_bad_lengths(monkeypatch):
    """Lengths that live on the device are not read back by default (no synchronisation on the training step): a bad
    row comes back as NaN.  With validate_lengths=True (or SSD_VALIDATE_LENGTHS=1) the reference's RuntimeError is back."""
    B, T = 3, 30
    lp = synth.flat(B, T, 5).cuda().transpose(0, 1)
    tg, tl = synth.targets(B, 6, 6)
    il = torch.tensor([30, 31, 30]).cuda()   # 31 > T
    out = ctc_loss(lp, tg.cuda(), il, tl.cuda(), blank=1, reduction="none", zero_infinity=True)
    assert torch.isfinite(out[0]) and torch.isfinite(out[2])
    with pytest.raises(RuntimeError):
        ctc_loss(lp, tg.cuda(), il, tl.cuda(), blank=1, reduction="none", zero_infinity=True, validate_lengths=True)
    monkeypatch.setenv("SSD_VALIDATE_LENGTHS", "1")
    with pytest.raises(RuntimeError):
        CTCLoss(blank=1)(lp, tg.cuda(), il, tl.cuda())


@pytest.mark.timeout(120)
def test_extreme_mismatch_is_redone_in_log_space():
    """The fuzz-found rows of tests/test_emu_kernels.py::test_loss_extreme_mismatch_is_redone_in_log_space inside a batch
    of ordinary ones, on the GPU: both self-checking kernels hand them down, the log-space pass returns the reference's
    values; the ordinary rows are untouched by it."""
    B, T, S, c = 600, 333, 255, 52
    lp, tg, tl = synth.peaked_aligned(B, T, 7000 + c, n_labels=(1, T // 2))
    tg = torch.nn.functional.pad(tg, (0, S - tg.size(1)))
    il = synth.ragged_lengths(B, T, 9000 + c)
    sel = torch.arange(0, 64)
    lp, tg, tl, il = lp[sel], tg[sel], tl[sel], il[sel]
    nll, grad = ops.ctc_loss_fwd_bwd(lp.cuda(), il, tg, tl, blank=1, mean_denominator=64)
    n64, _, g64 = ol.torch_ctc(lp, il, tg, tl, 1, True, torch.float64)
    assert torch.isfinite(grad).all()
    fin = torch.isfinite(n64) & (n64 != 0)
    assert ((nll.cpu().double() - n64)[fin].abs() / n64[fin].abs()).max().item() < RTOL
    err = (grad.cpu().double() - g64).abs().amax(dim=(1, 2))
    ref = g64.abs().amax(dim=(1, 2))
    assert (err <= RTOL * ref + 1e-12).all(), (err / ref.clamp_min(1e-30)).max().item()


def test_device_prefetcher_delivers_every_batch_intact():
    """data/prefetch.py: two alternating device buffer sets, copies one batch ahead on a side stream.  Every batch must
    arrive with its own contents while the consumer is still busy with the previous one."""
    from silent_speech_decoder_b200.data.prefetch import DevicePrefetcher

    host = [(torch.full((64, 300, 48), float(i)).pin_memory(), torch.tensor([i, i + 1]).pin_memory()) for i in range(7)]
    sums = []
    for x, n in DevicePrefetcher(iter(host), torch.device("cuda", 0)):
        y = x
        for _ in range(20):  # keep the consumer's stream busy so that the next copy really overlaps
            y = y * 1.0000001
        sums.append((float(x[0, 0, 0]), float(y.mean()), n.tolist()))
    assert [s[0] for s in sums] == [float(i) for i in range(7)]
    assert [s[2] for s in sums] == [[i, i + 1] for i in range(7)]
    assert all(abs(s[1] - s[0]) <= 1e-4 * max(1.0, s[0]) for s in sums)
