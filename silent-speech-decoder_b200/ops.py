"""Torch-facing launchers for the kernels.

PyTorch is used here for device memory, streams and dtype plumbing only; all
arithmetic happens inside libssd_ctc.so.  Inputs may live on the host (the
reference hands its decoders CPU tensors, ``evaluate.py:299``): they are moved to
the current CUDA device.  Without CUDA these functions raise -- there is no
CPU path.
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _native

_WS: Dict[Tuple[int, str], torch.Tensor] = {}
LAUNCHES = {"greedy": 0, "beam": 0, "loss": 0, "edit": 0}  # kernels launched by this process (bench reads it)


def _device(t: torch.Tensor) -> torch.device:
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError(
            "silent_speech_decoder_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback "
            "for the CTC kernels."
        )
    return torch.device("cuda", torch.cuda.current_device())


def _workspace(dev: torch.device, kind: str, nbytes: int) -> torch.Tensor:
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), kind)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
        _WS[key] = buf
    return buf


def _stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _prep_log_probs(log_probs: torch.Tensor, dev: torch.device) -> torch.Tensor:
    if log_probs.dim() != 3:
        raise ValueError(f"log_probs must be (batch, time, vocab); got shape {tuple(log_probs.shape)}")
    x = log_probs.detach()
    if x.dtype != torch.float32:
        x = x.float()
    if x.device != dev:
        x = x.to(dev, non_blocking=True)
    if x.size(-1) > 0 and x.stride(-1) != 1:
        x = x.contiguous()
    return x


def _strides(x: torch.Tensor) -> Tuple[int, int]:
    B, T, V = x.shape
    sb = x.stride(0) if B > 1 else T * V
    st = x.stride(1) if T > 1 else V
    return sb, st


def _prep_lengths(lengths: torch.Tensor, B: int, dev: torch.device) -> torch.Tensor:
    ln = torch.as_tensor(lengths).detach().reshape(-1)
    if ln.numel() != B:
        raise ValueError(f"lengths has {ln.numel()} entries for a batch of {B}")
    return ln.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()


def greedy_ids(log_probs: torch.Tensor, lengths: torch.Tensor, blank: int, blank_bias: float = 0.0):
    """-> (tokens int32 (B,T), counts int32 (B,)) on the device."""
    lib = _native.load()
    dev = _device(log_probs)
    with torch.cuda.device(dev):
        x = _prep_log_probs(log_probs, dev)
        B, T, V = x.shape
        ln = _prep_lengths(lengths, B, dev)
        tok = torch.empty((B, max(T, 1)), dtype=torch.int32, device=dev)
        cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
        if B and T:
            sb, st = _strides(x)
            lib.check(lib.ssd_ctc_greedy(x.data_ptr(), B, T, V, sb, st, ln.data_ptr(), int(blank), float(blank_bias),
                                         tok.data_ptr(), cnt.data_ptr(), _stream_ptr(dev)))
            LAUNCHES["greedy"] += 1
    return tok, cnt


def beam_ids(
    log_probs: torch.Tensor,
    lengths: torch.Tensor,
    blank: int,
    pad: int,
    space: int,
    beam_width: int = 50,
    beam_prune_logp: float = -10.0,
    blank_bias: float = 0.0,
    token_min_logp: float = -5.0,
    sort_by_length: bool = True,
):
    """-> (tokens int32 (B,T), counts int32 (B,), scores float64 (B,)) on the device."""
    lib = _native.load()
    dev = _device(log_probs)
    with torch.cuda.device(dev):
        x = _prep_log_probs(log_probs, dev)
        B, T, V = x.shape
        ln = _prep_lengths(lengths, B, dev)
        tok = torch.empty((B, max(T, 1)), dtype=torch.int32, device=dev)
        cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
        score = torch.zeros((B,), dtype=torch.float64, device=dev)
        if B and T:
            sb, st = _strides(x)
            order = torch.argsort(ln, descending=True, stable=True).to(torch.int32) if (sort_by_length and B > 1) else None
            nbytes = int(lib.ssd_ctc_beam_workspace_bytes(B, T, V, int(beam_width)))
            ws = _workspace(dev, "beam", nbytes)
            lib.check(lib.ssd_ctc_beam(x.data_ptr(), B, T, V, sb, st, ln.data_ptr(),
                                       order.data_ptr() if order is not None else None,
                                       int(blank), int(pad), int(space), float(blank_bias), int(beam_width),
                                       float(beam_prune_logp), float(token_min_logp),
                                       tok.data_ptr(), cnt.data_ptr(), score.data_ptr(), ws.data_ptr(), ws.numel(),
                                       _stream_ptr(dev)))
            LAUNCHES["beam"] += 1
    return tok, cnt, score


def ctc_loss_fwd_bwd(
    log_probs_btv: torch.Tensor,
    input_lengths: torch.Tensor,
    targets: torch.Tensor,
    target_lengths: torch.Tensor,
    blank: int,
    zero_infinity: bool = True,
    mean_denominator: int = 0,
    want_grad: bool = True,
    grad_mul: Optional[torch.Tensor] = None,
):
    """One fused launch. log_probs_btv is indexed (b,t,v) with arbitrary b/t strides.
    -> (nll fp32 (B,), grad with the strides of log_probs_btv or None)."""
    lib = _native.load()
    dev = _device(log_probs_btv)
    with torch.cuda.device(dev):
        x = _prep_log_probs(log_probs_btv, dev)
        B, T, V = x.shape
        il = _prep_lengths(input_lengths, B, dev)
        tl = _prep_lengths(target_lengths, B, dev)
        tg = torch.as_tensor(targets).detach().to(device=dev, dtype=torch.int64, non_blocking=True)
        if tg.dim() == 1:
            # torch also accepts concatenated targets: scatter them into a padded matrix
            S = int(tl.max().item()) if B else 0
            pad = torch.zeros((B, max(S, 1)), dtype=torch.int64, device=dev)
            if S:
                pos = torch.arange(S, device=dev)[None, :]
                mask = pos < tl[:, None]
                pad[:, :S][mask] = tg[: int(mask.sum().item())]
            tg = pad
        elif tg.dim() != 2 or tg.size(0) != B:
            raise ValueError(f"targets must be (batch, max_target_len); got {tuple(tg.shape)}")
        if tg.size(1) > 0 and tg.stride(1) != 1:
            tg = tg.contiguous()
        S = tg.size(1)
        nll = torch.zeros((B,), dtype=torch.float32, device=dev)
        grad = torch.empty_strided(x.shape, x.stride(), dtype=torch.float32, device=dev) if want_grad else None
        if B:
            sb, st = _strides(x)
            nbytes = int(lib.ssd_ctc_loss_workspace_bytes(B, T, V, S))
            ws = _workspace(dev, "loss", nbytes)
            # ragged batches with host-resident lengths (what the reference passes, losses.py:55): longest utterances
            # first, so that the CTAs that start last get the short ones.  Device-resident lengths are not inspected
            # (no sync, no extra launches on the training step).
            order = None
            hl = torch.as_tensor(input_lengths)
            if B > 1 and not hl.is_cuda and int(hl.min()) != int(hl.max()):
                order = torch.argsort(hl.reshape(-1), descending=True, stable=True).to(torch.int32).to(dev)
            gm = None
            if grad_mul is not None:
                gm = grad_mul.detach().to(device=dev, dtype=torch.float32).reshape(1)
            lib.check(lib.ssd_ctc_loss(x.data_ptr(), B, T, V, sb, st, tg.data_ptr() if S else None,
                                       tg.stride(0) if (S and B > 1) else max(S, 1), S, il.data_ptr(), tl.data_ptr(),
                                       order.data_ptr() if order is not None else None,
                                       int(blank), int(bool(zero_infinity)), int(mean_denominator),
                                       gm.data_ptr() if gm is not None else None, nll.data_ptr(),
                                       grad.data_ptr() if grad is not None else None, ws.data_ptr(), ws.numel(),
                                       _stream_ptr(dev)))
            # fp32 fast-path kernel + the fp64 kernel over whatever it flagged, or the fp64 kernel alone
            two = int(lib.ssd_ctc_loss_fallback_counter_offset(B, T, V, S)) != _NO_FAST
            LAUNCHES["loss"] += 2 if two else 1
    return nll, grad


_NO_FAST = 2 ** 64 - 1


def loss_fallback_count(B: int, T: int, V: int, S: int, device=None) -> Optional[int]:
    """How many utterances of the LAST ctc_loss_fwd_bwd call of this shape were recomputed by the exact fp64
    kernel (flagged by the fp32 fast path); None if the fast path does not cover the shape. Synchronises."""
    lib = _native.load()
    off = int(lib.ssd_ctc_loss_fallback_counter_offset(B, T, V, S))
    if off == _NO_FAST:
        return None
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    ws = _WS.get((dev.index if dev.index is not None else torch.cuda.current_device(), "loss"))
    if ws is None or ws.numel() < off + 4:
        return None
    return int(ws[off:off + 4].view(torch.int32).item())


def edit_counts(ref_tok: torch.Tensor, ref_off: torch.Tensor, hyp_tok: torch.Tensor, hyp_off: torch.Tensor,
                max_ref_len: int, max_hyp_len: int) -> torch.Tensor:
    """Levenshtein alignment counts of N token-sequence pairs given CSR style (int32 ids >= 0, int64 offsets
    of N+1 entries). -> int32 (N,4) on the device: substitutions, deletions, insertions, hits."""
    lib = _native.load()
    dev = _device(ref_tok)
    with torch.cuda.device(dev):
        rt = ref_tok.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
        ht = hyp_tok.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
        ro = ref_off.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
        ho = hyp_off.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
        N = ro.numel() - 1
        if N < 0 or ho.numel() != N + 1:
            raise ValueError(f"offset arrays must both have N+1 entries; got {ro.numel()} and {ho.numel()}")
        out = torch.zeros((N, 4), dtype=torch.int32, device=dev)
        if N:
            lib.check(lib.ssd_edit_counts(rt.data_ptr() if rt.numel() else None, ro.data_ptr(),
                                          ht.data_ptr() if ht.numel() else None, ho.data_ptr(), N, int(max_ref_len),
                                          int(max_hyp_len), out.data_ptr(), _stream_ptr(dev)))
            LAUNCHES["edit"] += 1
    return out
