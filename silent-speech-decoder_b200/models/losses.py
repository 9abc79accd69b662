"""Distillation + CTC loss with the reference's interface; the CTC term runs on the
fused sm_100a forward-backward kernel (``ssd_ctc_loss``).

Drop-in for reference ``src/models/losses.py``: ``LossWeights`` and
``DistillationCTCLoss(vocab_size, blank_id, weights, normalize_distill)`` with
``forward(log_probs, logit_lengths, targets, target_lengths, student_repr, teacher_repr,
teacher_lengths=None) -> {"total", "ctc", "distill"}``.  ``self.ctc_loss`` is a module with
``torch.nn.CTCLoss``'s call shape, as the reference invokes it at ``losses.py:56``.
The distillation term is small dense PyTorch work and stays PyTorch (north_star).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


@dataclass
class LossWeights:
    lambda_distill: float = 0.7
    lambda_ctc: float = 0.3


def _check_host_lengths(name: str, t: torch.Tensor, limit: int, what: str) -> None:
    """Lengths that already live on the host are validated for free, with torch's wording."""
    if isinstance(t, torch.Tensor) and not t.is_cuda and t.numel():
        hi, lo = int(t.max()), int(t.min())
        if lo < 0:
            raise RuntimeError(f"Expected {name} to have value at least 0, but got value {lo}")
        if hi > limit:
            raise RuntimeError(f"Expected {name} to have value at most {limit}, but got value {hi} ({what})")


class _CTCNll(torch.autograd.Function):
    """(T,B,V) log-probs -> per-utterance nll (B,).  Forward runs the fused kernel once and keeps
    d nll_b / d log_probs; backward only scales it by the incoming per-utterance gradient."""

    @staticmethod
    def forward(ctx, log_probs_tbv, targets, input_lengths, target_lengths, blank: int, zero_infinity: bool):
        need = ctx.needs_input_grad[0]
        x = log_probs_tbv.transpose(0, 1)  # (B,T,V) view of the same memory, no copy
        nll, grad = ops.ctc_loss_fwd_bwd(x, input_lengths, targets, target_lengths, blank=blank,
                                         zero_infinity=zero_infinity, mean_denominator=0, want_grad=need)
        ctx.in_dtype = log_probs_tbv.dtype
        ctx.in_device = log_probs_tbv.device
        if need:
            ctx.save_for_backward(grad)
        out = nll.to(log_probs_tbv.dtype)
        return out if out.device == log_probs_tbv.device else out.to(log_probs_tbv.device)

    @staticmethod
    def backward(ctx, g_nll):
        (grad_btv,) = ctx.saved_tensors
        g = g_nll.to(device=grad_btv.device, dtype=torch.float32).view(-1, 1, 1)
        out = (grad_btv * g).transpose(0, 1)  # back to (T,B,V)
        if out.dtype != ctx.in_dtype or out.device != ctx.in_device:
            out = out.to(device=ctx.in_device, dtype=ctx.in_dtype)
        return out, None, None, None, None, None


def ctc_loss(log_probs, targets, input_lengths, target_lengths, blank: int = 0, reduction: str = "mean",
             zero_infinity: bool = False) -> torch.Tensor:
    """``torch.nn.functional.ctc_loss`` semantics on the native kernel.  log_probs: (T,B,V)."""
    if log_probs.dim() == 2:  # unbatched (T,V), as torch accepts
        log_probs = log_probs.unsqueeze(1)
        targets = targets.unsqueeze(0) if targets.dim() == 1 else targets
        input_lengths = torch.as_tensor(input_lengths).reshape(1)
        target_lengths = torch.as_tensor(target_lengths).reshape(1)
        unbatched = True
    else:
        unbatched = False
    if log_probs.dim() != 3:
        raise RuntimeError(f"log_probs must be (T, N, C); got {tuple(log_probs.shape)}")
    if log_probs.dtype not in (torch.float32, torch.float64):
        raise RuntimeError(f'"ctc_loss" not implemented for {log_probs.dtype}')
    T, B, V = log_probs.shape
    input_lengths = torch.as_tensor(input_lengths)
    target_lengths = torch.as_tensor(target_lengths)
    if not (0 <= blank < V):
        raise RuntimeError("blank must be in label range")
    if input_lengths.numel() != B or target_lengths.numel() != B:
        raise RuntimeError(f"input_lengths and target_lengths must be of size batch_size ({B})")
    _check_host_lengths("input_lengths", input_lengths, T, "the time dimension of log_probs")
    if targets.dim() == 2:
        _check_host_lengths("target_lengths", target_lengths, targets.size(1), "the width of targets")
    nll = _CTCNll.apply(log_probs, targets, input_lengths, target_lengths, int(blank), bool(zero_infinity))
    if reduction == "none":
        return nll[0] if unbatched else nll
    if reduction == "sum":
        return nll.sum()
    if reduction == "mean":
        tl = target_lengths.to(device=nll.device).clamp_min(1).to(nll.dtype)
        return (nll / tl).mean()
    raise ValueError(f"{reduction} is not a valid value for reduction")


class CTCLoss(nn.Module):
    """``torch.nn.CTCLoss`` call shape: ``(log_probs (T,B,V), targets, input_lengths, target_lengths)``."""

    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = False):
        super().__init__()
        self.blank, self.reduction, self.zero_infinity = blank, reduction, zero_infinity

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        return ctc_loss(log_probs, targets, input_lengths, target_lengths, self.blank, self.reduction, self.zero_infinity)


class DistillationCTCLoss(nn.Module):
    def __init__(self, vocab_size: int, blank_id: int, weights: LossWeights, normalize_distill: bool = False):
        super().__init__()
        del vocab_size  # accepted and unused, as in the reference (losses.py:20-30)
        self.ctc_loss = CTCLoss(blank=blank_id, zero_infinity=True)
        self.weights = weights  # reassigned from outside every epoch (train.py:486-489)
        self.normalize_distill = normalize_distill

    def forward(
        self,
        log_probs: torch.Tensor,
        logit_lengths: torch.Tensor,
        targets: torch.Tensor,
        target_lengths: torch.Tensor,
        student_repr: torch.Tensor,
        teacher_repr: Optional[torch.Tensor],
        teacher_lengths: Optional[torch.Tensor] = None,
    ) -> Dict[str, torch.Tensor]:
        # (B,T,V) -> (T,B,V) view; lengths stay where they are (the reference moves them to the
        # host for ATen, losses.py:55 -- the native kernel reads them on the device)
        ctc = self.ctc_loss(log_probs.transpose(0, 1), targets, logit_lengths, target_lengths)
        distill = torch.tensor(0.0, device=log_probs.device)
        if teacher_repr is not None:
            distill = self._distill(student_repr, teacher_repr, logit_lengths, teacher_lengths)
        total = self.weights.lambda_ctc * ctc + self.weights.lambda_distill * distill
        return {"total": total, "ctc": ctc, "distill": distill}

    def _distill(self, student, teacher, student_lengths, teacher_lengths):
        """Masked mean-squared error between student and (time-resampled) teacher features;
        behaviour of losses.py:58-102: teacher linearly interpolated to the student's length,
        its lengths rescaled and rounded, frames beyond min(student, teacher) length masked
        out, mean over valid frames x feature dim (denominator at least 1)."""
        dev = student.device
        n_s, n_t = student.size(1), teacher.size(1)
        t_len = teacher_lengths.to(dev) if teacher_lengths is not None else None
        if n_t != n_s:
            teacher = F.interpolate(teacher.transpose(1, 2), size=n_s, mode="linear", align_corners=False).transpose(1, 2)
            if t_len is not None and n_t > 0:
                t_len = torch.round(t_len.float() * (float(n_s) / float(n_t))).long().clamp(max=n_s)
        valid = student_lengths.to(dev)
        if t_len is not None:
            valid = torch.minimum(valid, t_len)
        valid = valid.clamp(max=n_s)
        mask = torch.arange(n_s, device=dev)[None, :] < valid[:, None]
        if self.normalize_distill:
            student = F.layer_norm(student, student.shape[-1:])
            teacher = F.layer_norm(teacher, teacher.shape[-1:])
        sq = (student - teacher).pow(2) * mask.unsqueeze(-1)
        return sq.sum() / (mask.sum() * student.size(-1)).clamp_min(1)
