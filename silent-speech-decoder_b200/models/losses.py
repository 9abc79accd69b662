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

import os
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


def _dense(t: torch.Tensor) -> bool:
    """True when the tensor covers its storage range without holes or overlap (any permutation of a contiguous one)."""
    sizes_strides = sorted(((st, sz) for sz, st in zip(t.shape, t.stride()) if sz > 1))
    expect = 1
    for st, sz in sizes_strides:
        if st != expect:
            return False
        expect *= sz
    return True


class _CTCFused(torch.autograd.Function):
    """(T,B,V) log-probs (or logits) -> the REDUCED loss, forward and backward in the one fused launch of forward.

    The kernel writes d loss / d input already scaled for the reduction ('mean': 1 / (B * max(S_b,1)) per utterance,
    in-kernel) and by ``expect`` -- the upstream scalar the caller knows it will receive (``lambda_ctc`` when the
    loss is used as at losses.py:104-105).  backward then hands the buffer out as it is: the incoming scalar is
    compared with ``expect`` on the device (``ssd_scale_if``) and the gradient is rescaled in place only if they
    differ.  No host synchronisation, no extra pass over the (B,T,V) gradient in the common case.
    reduction 'none' keeps a per-utterance incoming gradient and multiplies row-wise."""

    @staticmethod
    def forward(ctx, x_tbv, targets, input_lengths, target_lengths, blank: int, zero_infinity: bool, reduction: str,
                from_logits: bool, expect: float):
        need = ctx.needs_input_grad[0]
        x = x_tbv.transpose(0, 1)  # (B,T,V) view of the same memory, no copy
        B = x.size(0)
        expect = float(expect) if (expect and reduction != "none") else 1.0
        gm = None
        if need and expect != 1.0:
            gm = ops.device_scalar(ops._device(x), expect)
        nll, grad = ops.ctc_loss_fwd_bwd(x, input_lengths, targets, target_lengths, blank=blank, zero_infinity=zero_infinity,
                                         mean_denominator=B if reduction == "mean" else 0, want_grad=need, grad_mul=gm,
                                         from_logits=from_logits)
        ctx.in_dtype, ctx.in_device, ctx.reduction, ctx.expect = x_tbv.dtype, x_tbv.device, reduction, expect
        ctx.grad_btv, ctx.applied = grad, None
        if reduction == "none":
            out = nll
        elif reduction == "sum":
            out = nll.sum()
        else:
            parts = ops.loss_parts(nll, target_lengths)  # [sum_b nll_b / max(S_b,1), B] in one launch
            out = parts[0] / parts[1]
        out = out.to(x_tbv.dtype)
        return out if out.device == x_tbv.device else out.to(x_tbv.device)

    @staticmethod
    def backward(ctx, g):
        grad = ctx.grad_btv
        if ctx.reduction == "none":
            out = grad * g.to(device=grad.device, dtype=torch.float32).view(-1, 1, 1)
        else:
            g = g.to(device=grad.device, dtype=torch.float32).reshape(1)
            if ctx.applied is None and _dense(grad):
                ops.scale_if(grad, g, ctx.expect)  # no-op on the device when g is what forward scaled by
                ctx.applied = g
                out = grad
            else:  # backward through the same graph again (retain_graph), or an input layout with holes
                base = ctx.applied if ctx.applied is not None else torch.full_like(g, ctx.expect)
                out = grad * (g / base)
        out = out.transpose(0, 1)  # back to (T,B,V)
        if out.dtype != ctx.in_dtype or out.device != ctx.in_device:
            out = out.to(device=ctx.in_device, dtype=ctx.in_dtype)
        return out, None, None, None, None, None, None, None, None


def _check_device_lengths(input_lengths, target_lengths, T: int, S_max) -> None:
    """Opt-in (``validate_lengths=True`` or ``SSD_VALIDATE_LENGTHS=1``): torch raises on out-of-range lengths; the
    kernels do not read device-resident lengths back (that would synchronise every step) and turn such a row into
    NaN instead.  This check costs that synchronisation and restores the exception."""
    bad_in = ((input_lengths < 0) | (input_lengths > T)).any()
    bad_tg = (target_lengths < 0).any() if S_max is None else ((target_lengths < 0) | (target_lengths > S_max)).any()
    if bool(bad_in):
        raise RuntimeError(f"Expected input_lengths to have value in [0, {T}] (the time dimension of log_probs)")
    if bool(bad_tg):
        raise RuntimeError("Expected target_lengths to have value in [0, the width of targets]")


def ctc_loss(log_probs, targets, input_lengths, target_lengths, blank: int = 0, reduction: str = "mean",
             zero_infinity: bool = False, from_logits: bool = False, expect_grad: float = 1.0,
             validate_lengths: bool = False) -> torch.Tensor:
    """``torch.nn.functional.ctc_loss`` semantics on the native kernel.  log_probs: (T,B,V).

    ``from_logits`` (scope row f-2): the input is the CTC head's pre-softmax output (``heads.py:36``); the result is
    the loss of ``log_softmax(input)`` and the gradient is taken with respect to the logits, without the log-prob
    tensor ever existing.  ``expect_grad``: the scalar the caller expects autograd to send back (see _CTCFused)."""
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
    if reduction not in ("none", "sum", "mean"):
        raise ValueError(f"{reduction} is not a valid value for reduction")
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
    if (validate_lengths or os.environ.get("SSD_VALIDATE_LENGTHS") == "1") and (input_lengths.is_cuda or target_lengths.is_cuda):
        _check_device_lengths(input_lengths, target_lengths, T, targets.size(1) if targets.dim() == 2 else None)
    out = _CTCFused.apply(log_probs, targets, input_lengths, target_lengths, int(blank), bool(zero_infinity), reduction,
                          bool(from_logits), float(expect_grad))
    if reduction == "none" and unbatched:
        return out[0]
    return out


class CTCLoss(nn.Module):
    """``torch.nn.CTCLoss`` call shape: ``(log_probs (T,B,V), targets, input_lengths, target_lengths)``."""

    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = False, from_logits: bool = False):
        super().__init__()
        self.blank, self.reduction, self.zero_infinity, self.from_logits = blank, reduction, zero_infinity, from_logits
        self.expect_grad = 1.0  # set by the owner when it knows the upstream scalar (DistillationCTCLoss: lambda_ctc)

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        return ctc_loss(log_probs, targets, input_lengths, target_lengths, self.blank, self.reduction, self.zero_infinity,
                        from_logits=self.from_logits, expect_grad=self.expect_grad)


class DistillationCTCLoss(nn.Module):
    def __init__(self, vocab_size: int, blank_id: int, weights: LossWeights, normalize_distill: bool = False,
                 from_logits: bool = False):
        super().__init__()
        del vocab_size  # accepted and unused, as in the reference (losses.py:20-30)
        # from_logits (not in the reference, scope row f-2): `log_probs` is then the head's pre-softmax output
        self.ctc_loss = CTCLoss(blank=blank_id, zero_infinity=True, from_logits=from_logits)
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
        # total = lambda_ctc * ctc + ...: autograd will send lambda_ctc back; the kernel scales by it up front
        self.ctc_loss.expect_grad = float(self.weights.lambda_ctc) or 1.0
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
