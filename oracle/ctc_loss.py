"""CTC loss oracle (test infrastructure only -- see oracle/__init__.py).

Two independent CPU statements of the path ``DistillationCTCLoss.forward``
(reference ``src/models/losses.py:52-56,104-105``) -> ``torch.nn.CTCLoss`` ->
ATen ``ctc_loss_cpu`` (third-party: torch, ``environment.yml:12`` pins
``pytorch>=2.0``; this image has torch 2.11.0):

``torch_ctc``       calls the very implementation the reference calls, on CPU,
                    in fp32 (what the reference runs) or fp64 (ground truth for
                    tolerances: at T=1000 the fp32 run itself is ~3e-3 off).
``numpy_ctc``       a from-scratch numpy fp64 restatement of the published
                    algorithm (Graves et al. 2006, eqs. 6-8, 10-11, 16, in log
                    space) with ATen's conventions, to cross-check ``torch_ctc``.

Conventions restated (verified by running torch 2.11 here):
  * nll_b = -logsumexp(alpha[T_b-1, 2S_b], alpha[T_b-1, 2S_b-1])
  * d nll_b / d lp[t,b,v] = exp(lp) - exp(logsum_{s: l'_s = v}(alpha+beta) + nll - lp)
    for t < T_b, exactly 0 for t >= T_b
  * infeasible sample: nll = inf; with zero_infinity -> nll 0 and zero gradient,
    still counted in the batch mean
  * reduction 'mean' = mean_b( nll_b / max(S_b, 1) )
"""

from __future__ import annotations

from typing import Tuple

import numpy as np
import torch
import torch.nn.functional as F


def torch_ctc(
    log_probs_btv: torch.Tensor,
    input_lengths: torch.Tensor,
    targets: torch.Tensor,
    target_lengths: torch.Tensor,
    blank: int,
    zero_infinity: bool = True,
    dtype: torch.dtype = torch.float64,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Returns (nll[B], mean_loss, d mean_loss / d log_probs as (B,T,V))."""
    lp = log_probs_btv.detach().to("cpu", dtype).clone().requires_grad_(True)
    il = input_lengths.detach().cpu().to(torch.int64)
    tl = target_lengths.detach().cpu().to(torch.int64)
    tg = targets.detach().cpu().to(torch.int64)
    nll = F.ctc_loss(lp.transpose(0, 1), tg, il, tl, blank=blank, reduction="none", zero_infinity=zero_infinity)
    mean = (nll / tl.clamp_min(1).to(dtype)).mean()
    mean.backward()
    return nll.detach(), mean.detach(), lp.grad.detach()


def reference_module_ctc(log_probs, logit_lengths, targets, target_lengths, blank_id: int):
    """Restates the CTC term of ``DistillationCTCLoss.forward`` (losses.py:52-56) in fp32:
    transpose to (T,B,V), lengths to CPU, ``nn.CTCLoss(blank, zero_infinity=True)`` ('mean')."""
    crit = torch.nn.CTCLoss(blank=blank_id, zero_infinity=True)
    lp = log_probs.detach().cpu().float().clone().requires_grad_(True)
    loss = crit(lp.transpose(0, 1), targets.cpu(), logit_lengths.cpu(), target_lengths.cpu())
    loss.backward()
    return loss.detach(), lp.grad.detach()


def _lse(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    m = np.maximum(a, b)
    m_safe = np.where(np.isfinite(m), m, 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = m_safe + np.log(np.exp(a - m_safe) + np.exp(b - m_safe))
    return np.where(np.isneginf(m), -np.inf, out)


def numpy_ctc_single(lp_tv: np.ndarray, labels: np.ndarray, blank: int):
    """One utterance. lp_tv: (T,V) float64 log-probs (already cut to T_b frames),
    labels: (S,) ints. Returns (nll, d nll / d lp (T,V))."""
    lp = np.asarray(lp_tv, dtype=np.float64)
    T, V = lp.shape
    S = int(len(labels))
    L = 2 * S + 1
    ext = np.full(L, blank, dtype=np.int64)
    ext[1::2] = labels
    NEG = -np.inf
    if T == 0:
        nll = 0.0 if S == 0 else np.inf
        return nll, np.zeros((0, V))
    # skip[s]: transition s-2 -> s allowed
    skip = np.zeros(L, dtype=bool)
    if L > 2:
        skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
    alpha = np.full((T, L), NEG)
    alpha[0, 0] = lp[0, blank]
    if L > 1:
        alpha[0, 1] = lp[0, ext[1]]
    for t in range(1, T):
        a = alpha[t - 1]
        a1 = np.full(L, NEG); a1[1:] = a[:-1]
        a2 = np.full(L, NEG); a2[2:] = a[:-2]
        a2 = np.where(skip, a2, NEG)
        alpha[t] = _lse(_lse(a, a1), a2) + lp[t, ext]
    tail = alpha[T - 1, L - 1]
    if L > 1:
        tail = _lse(np.array(tail), np.array(alpha[T - 1, L - 2]))
    nll = -float(tail)
    beta = np.full((T, L), NEG)
    beta[T - 1, L - 1] = lp[T - 1, blank]
    if L > 1:
        beta[T - 1, L - 2] = lp[T - 1, ext[L - 2]]
    skip_f = np.zeros(L, dtype=bool)  # transition s -> s+2 allowed
    if L > 2:
        skip_f[:-2] = skip[2:]
    for t in range(T - 2, -1, -1):
        b = beta[t + 1]
        b1 = np.full(L, NEG); b1[:-1] = b[1:]
        b2 = np.full(L, NEG); b2[:L - 2] = b[2:]
        b2 = np.where(skip_f, b2, NEG)
        beta[t] = _lse(_lse(b, b1), b2) + lp[t, ext]
    grad = np.zeros((T, V))
    if not np.isfinite(nll):
        return nll, np.full((T, V), np.nan)
    ab = alpha + beta  # (T,L)
    occ = np.full((T, V), NEG)
    for s in range(L):
        occ[:, ext[s]] = _lse(occ[:, ext[s]], ab[:, s])
    with np.errstate(over="ignore", invalid="ignore"):
        grad = np.exp(lp) - np.exp(occ + nll - lp)
    return nll, grad


def numpy_ctc(log_probs_btv, input_lengths, targets, target_lengths, blank: int, zero_infinity: bool = True):
    """Batch wrapper: returns (nll[B] float64, mean, grad (B,T,V) float64 of the mean)."""
    lp = np.asarray(log_probs_btv, dtype=np.float64)
    B, T, V = lp.shape
    il = np.asarray(input_lengths, dtype=np.int64)
    tl = np.asarray(target_lengths, dtype=np.int64)
    tg = np.asarray(targets, dtype=np.int64)
    nll = np.zeros(B)
    grad = np.zeros((B, T, V))
    for b in range(B):
        n, g = numpy_ctc_single(lp[b, : il[b]], tg[b, : tl[b]], blank)
        if not np.isfinite(n):
            if zero_infinity:
                n, g = 0.0, np.zeros((il[b], V))
        nll[b] = n
        grad[b, : il[b]] = g / (B * max(int(tl[b]), 1))
    mean = float(np.mean(nll / np.maximum(tl, 1)))
    return nll, mean, grad
