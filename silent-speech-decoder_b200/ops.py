"""Torch-facing launchers for the kernels.

PyTorch is used here for device memory, streams and dtype plumbing only; all
arithmetic happens inside libssd_ctc.so.  Inputs may live on the host (the
reference hands its decoders CPU tensors, ``evaluate.py:299``): they are moved to
the current CUDA device.  Without CUDA these functions raise -- there is no
CPU path.
"""

from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _native

# One scratch buffer per (device, stream, kind): two streams (or two threads, each with its own current stream)
# never share checkpoints, work-queue counters or back-pointer nodes.  Work on ONE stream is ordered, so reuse
# across calls on that stream is safe; a buffer that has to grow is replaced, and the old one is returned to
# torch's caching allocator, which only hands it out again to work queued on the same stream.
_WS: Dict[Tuple[int, int, str], torch.Tensor] = {}
_LAST_LOSS_WS: Dict[int, torch.Tensor] = {}
LAUNCHES = {"greedy": 0, "beam": 0, "loss": 0, "edit": 0, "aux": 0}  # kernels launched by this process (bench reads it)


def _device(t: torch.Tensor) -> torch.device:
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError(
            "silent_speech_decoder_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback "
            "for the CTC kernels."
        )
    return torch.device("cuda", torch.cuda.current_device())


def _dev_index(dev: torch.device) -> int:
    return dev.index if dev.index is not None else torch.cuda.current_device()


def _workspace(dev: torch.device, kind: str, nbytes: int) -> torch.Tensor:
    key = (_dev_index(dev), int(torch.cuda.current_stream(dev).cuda_stream), kind)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)  # allocated on (and owned by) the current stream
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


def beam_grid_ids(
    log_probs: torch.Tensor,
    lengths: torch.Tensor,
    blank: int,
    pad: int,
    space: int,
    settings: Sequence[Tuple[int, float, float, float]],
    sort_by_length: bool = True,
):
    """Several decoder settings ``(beam_width, beam_prune_logp, blank_bias, token_min_logp)`` over the same batch in
    ONE launch (``ssd_ctc_beam_grid``).  -> (tokens int32 (n,B,T), counts int32 (n,B), scores float64 (n,B)) on the
    device, rows in the order of ``settings``."""
    lib = _native.load()
    dev = _device(log_probs)
    n = len(settings)
    if n > 8:
        raise ValueError(f"at most 8 decoder settings per launch; got {n}")
    with torch.cuda.device(dev):
        x = _prep_log_probs(log_probs, dev)
        B, T, V = x.shape
        ln = _prep_lengths(lengths, B, dev)
        tok = torch.empty((n, B, max(T, 1)), dtype=torch.int32, device=dev)
        cnt = torch.zeros((n, B), dtype=torch.int32, device=dev)
        score = torch.zeros((n, B), dtype=torch.float64, device=dev)
        if B and T and n:
            sb, st = _strides(x)
            # the kernel walks settings in array order: widest first keeps the long items at the front of the queue
            perm = sorted(range(n), key=lambda i: -int(settings[i][0]))
            arr = (_native.BeamSetting * n)(*[_native.BeamSetting(int(settings[i][0]), float(settings[i][1]), float(settings[i][2]),
                                                                   float(settings[i][3])) for i in perm])
            order = torch.argsort(ln, descending=True, stable=True).to(torch.int32) if (sort_by_length and B > 1) else None
            wmax = max(int(s[0]) for s in settings)
            ws = _workspace(dev, "beam", int(lib.ssd_ctc_beam_workspace_bytes(B * n, T, V, wmax)))
            import ctypes

            lib.check(lib.ssd_ctc_beam_grid(x.data_ptr(), B, T, V, sb, st, ln.data_ptr(),
                                            order.data_ptr() if order is not None else None, int(blank), int(pad), int(space),
                                            ctypes.cast(arr, ctypes.c_void_p), n, tok.data_ptr(), cnt.data_ptr(), score.data_ptr(),
                                            ws.data_ptr(), ws.numel(), _stream_ptr(dev)))
            LAUNCHES["beam"] += 1
            if perm != list(range(n)):
                inv = torch.tensor([perm.index(i) for i in range(n)], device=dev)
                tok, cnt, score = tok[inv], cnt[inv], score[inv]
    return tok, cnt, score


def beam_lm_ids(
    log_probs: torch.Tensor,
    lengths: torch.Tensor,
    blank: int,
    pad: int,
    space: int,
    lm,
    settings: Sequence[Tuple[int, float, float, float]],
    sort_by_length: bool = True,
):
    """LM-fused prefix beam search (``ssd_ctc_beam_lm``): ``lm`` is a ``decoding.lm.LanguageModelTables``; ``settings``
    as in ``beam_grid_ids``.  -> (tokens int32 (n,B,T), counts int32 (n,B), ranking scores float64 (n,B)) on the device."""
    import ctypes

    lib = _native.load()
    dev = _device(log_probs)
    n = len(settings)
    if n > 8:
        raise ValueError(f"at most 8 decoder settings per launch; got {n}")
    with torch.cuda.device(dev):
        x = _prep_log_probs(log_probs, dev)
        B, T, V = x.shape
        ln = _prep_lengths(lengths, B, dev)
        tok = torch.empty((n, B, max(T, 1)), dtype=torch.int32, device=dev)
        cnt = torch.zeros((n, B), dtype=torch.int32, device=dev)
        score = torch.zeros((n, B), dtype=torch.float64, device=dev)
        if B and T and n:
            sb, st = _strides(x)
            perm = sorted(range(n), key=lambda i: -int(settings[i][0]))
            arr = (_native.BeamSetting * n)(*[_native.BeamSetting(int(settings[i][0]), float(settings[i][1]), float(settings[i][2]),
                                                                   float(settings[i][3])) for i in perm])
            order = torch.argsort(ln, descending=True, stable=True).to(torch.int32) if (sort_by_length and B > 1) else None
            wmax = max(int(s[0]) for s in settings)
            ws = _workspace(dev, "beam_lm", int(lib.ssd_ctc_beam_lm_workspace_bytes(B * n, T, V, wmax)))
            tables = lm.struct_for(dev)
            lib.check(lib.ssd_ctc_beam_lm(x.data_ptr(), B, T, V, sb, st, ln.data_ptr(),
                                          order.data_ptr() if order is not None else None, int(blank), int(pad), int(space),
                                          ctypes.cast(arr, ctypes.c_void_p), n, ctypes.addressof(tables), tok.data_ptr(),
                                          cnt.data_ptr(), score.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev)))
            LAUNCHES["beam"] += 1
            if perm != list(range(n)):
                inv = torch.tensor([perm.index(i) for i in range(n)], device=dev)
                tok, cnt, score = tok[inv], cnt[inv], score[inv]
    return tok, cnt, score


def row_logsumexp(x: torch.Tensor) -> torch.Tensor:
    """(B,T,V) fp32 -> (B,T) fp32 log sum_v exp(x) on the device (``ssd_row_logsumexp``)."""
    lib = _native.load()
    dev = _device(x)
    with torch.cuda.device(dev):
        x = _prep_log_probs(x, dev)
        B, T, V = x.shape
        out = torch.empty((B, T), dtype=torch.float32, device=dev)
        if B and T:
            sb, st = _strides(x)
            lib.check(lib.ssd_row_logsumexp(x.data_ptr(), B, T, V, sb, st, out.data_ptr(), T, _stream_ptr(dev)))
            LAUNCHES["aux"] += 1
    return out


def scale_if(x: torch.Tensor, g: torch.Tensor, expected: float) -> None:
    """In place ``x *= g / expected`` decided ON THE DEVICE: when ``g == expected`` nothing is read or written.
    ``x`` must cover its storage densely (any permutation of a contiguous tensor)."""
    lib = _native.load()
    dev = x.device
    with torch.cuda.device(dev):
        gd = g.detach().to(device=dev, dtype=torch.float32).reshape(1)
        lib.check(lib.ssd_scale_if(x.data_ptr(), x.numel(), gd.data_ptr(), float(expected), _stream_ptr(dev)))
        LAUNCHES["aux"] += 1


def loss_parts(nll: torch.Tensor, target_lengths: torch.Tensor) -> torch.Tensor:
    """(2,) fp64 on the device: [sum_b nll_b / max(S_b, 1), B] -- one launch (``ssd_ctc_loss_parts``) instead of the
    five elementwise/reduction kernels torch needs for the same two numbers."""
    lib = _native.load()
    dev = nll.device
    with torch.cuda.device(dev):
        x = nll.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        tl = _prep_lengths(target_lengths, x.numel(), dev)
        out = torch.empty((2,), dtype=torch.float64, device=dev)
        lib.check(lib.ssd_ctc_loss_parts(x.data_ptr(), tl.data_ptr(), x.numel(), out.data_ptr(), _stream_ptr(dev)))
        LAUNCHES["aux"] += 1
    return out


_SCALARS: Dict[Tuple[int, float], torch.Tensor] = {}


def device_scalar(dev: torch.device, value: float) -> torch.Tensor:
    """A cached one-element fp32 device tensor (the kernels take multipliers by device pointer)."""
    key = (_dev_index(dev), float(value))
    t = _SCALARS.get(key)
    if t is None:
        if len(_SCALARS) > 256:
            _SCALARS.clear()
        t = torch.full((1,), float(value), dtype=torch.float32, device=dev)
        _SCALARS[key] = t
    return t


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
    from_logits: bool = False,
):
    """One fused launch. log_probs_btv is indexed (b,t,v) with arbitrary b/t strides.
    -> (nll fp32 (B,), grad with the strides of log_probs_btv or None).
    ``from_logits``: the input is the head's pre-softmax output (``ssd_ctc_loss_from_logits``); the gradient is then
    the one with respect to the logits."""
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
            nbytes = int((lib.ssd_ctc_loss_from_logits_workspace_bytes if from_logits else lib.ssd_ctc_loss_workspace_bytes)(B, T, V, S))
            ws = _workspace(dev, "loss", nbytes)
            _LAST_LOSS_WS[_dev_index(dev)] = ws
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
            entry = lib.ssd_ctc_loss_from_logits if from_logits else lib.ssd_ctc_loss
            lib.check(entry(x.data_ptr(), B, T, V, sb, st, tg.data_ptr() if S else None,
                                       tg.stride(0) if (S and B > 1) else max(S, 1), S, il.data_ptr(), tl.data_ptr(),
                                       order.data_ptr() if order is not None else None,
                                       int(blank), int(bool(zero_infinity)), int(mean_denominator),
                                       gm.data_ptr() if gm is not None else None, nll.data_ptr(),
                                       grad.data_ptr() if grad is not None else None, ws.data_ptr(), ws.numel(),
                                       _stream_ptr(dev)))
            # fp32 fast-path kernel + the fp64 kernel over whatever it flagged, or the fp64 kernel alone
            two = int(lib.ssd_ctc_loss_fallback_counter_offset(B, T, V, S)) != _NO_FAST
            LAUNCHES["loss"] += 2 if two else 1
            if from_logits:
                LAUNCHES["aux"] += 1
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
    ws = _LAST_LOSS_WS.get(_dev_index(dev))
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
