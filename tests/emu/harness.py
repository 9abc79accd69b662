"""Builds and drives the EMULATED kernel library (tests only).

``libssd_ctc_emu.so`` is the product's kernel sources compiled by g++ with -DSSD_EMU against
tests/emu/ssd_emu.h.  It lets the CPU-only test tier execute kernel logic; it is not part of
the product and nothing under silent-speech-decoder_b200/ refers to it.
"""

from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
CSRC = REPO / "silent-speech-decoder_b200" / "csrc"
# SSD_EMU_FLAGS: extra -D flags for a variant build (kept in its own directory), e.g. "-DSSD_F_DENSE=0"
EXTRA = os.environ.get("SSD_EMU_FLAGS", "").split()
OUT = HERE / ("_build" + ("_" + "".join(c if c.isalnum() else "_" for c in " ".join(EXTRA)) if EXTRA else ""))
LIB = OUT / "libssd_ctc_emu.so"
SOURCES = ["greedy.cu", "beam.cu", "ctc_loss.cu", "ctc_loss_fast.cu", "edit.cu", "aux.cu", "capi.cu"]


def build(force: bool = False) -> Path:
    OUT.mkdir(exist_ok=True)
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE / "ssd_emu.h", REPO / "include" / "ssd_ctc.h"]
    if not force and LIB.exists() and all(d.stat().st_mtime <= LIB.stat().st_mtime for d in deps):
        return LIB

    def cc(src):
        obj = OUT / (Path(src).stem + ".o")
        cmd = ["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-DSSD_EMU", *EXTRA, "-I", str(HERE), "-I", str(CSRC), "-x", "c++", "-c",
               str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"g++ failed on {src}:\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(4) as ex:
        objs = list(ex.map(cc, SOURCES))
    r = subprocess.run(["g++", "-shared", "-o", str(LIB), *map(str, objs)], capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stderr)
    return LIB


def library():
    from silent_speech_decoder_b200._native import Library

    return Library(build())


def _ptr(a):
    return None if a is None else a.ctypes.data


def greedy(lib, lp: np.ndarray, lengths, blank: int, bias: float = 0.0):
    lp = np.ascontiguousarray(lp, dtype=np.float32)
    B, T, V = lp.shape
    ln = np.asarray(lengths, dtype=np.int64)
    tok = np.full((B, max(T, 1)), -7, dtype=np.int32)
    n = np.full((B,), -7, dtype=np.int32)
    lib.check(lib.ssd_ctc_greedy(_ptr(lp), B, T, V, T * V, V, _ptr(ln), blank, bias, _ptr(tok), _ptr(n), None))
    return [[int(x) for x in tok[b, : n[b]]] for b in range(B)]


def beam(lib, lp: np.ndarray, lengths, blank: int, pad: int, space: int, width: int = 50, prune: float = -10.0,
         bias: float = 0.0, tmin: float = -5.0, order=None):
    lp = np.ascontiguousarray(lp, dtype=np.float32)
    B, T, V = lp.shape
    ln = np.asarray(lengths, dtype=np.int64)
    tok = np.full((B, max(T, 1)), -7, dtype=np.int32)
    n = np.full((B,), -7, dtype=np.int32)
    sc = np.full((B,), np.nan, dtype=np.float64)
    nbytes = lib.ssd_ctc_beam_workspace_bytes(B, T, V, width)
    ws = np.zeros(nbytes // 8 + 1, dtype=np.int64)
    od = None if order is None else np.asarray(order, dtype=np.int32)
    lib.check(lib.ssd_ctc_beam(_ptr(lp), B, T, V, T * V, V, _ptr(ln), _ptr(od), blank, pad, space, bias, width, prune, tmin,
                               _ptr(tok), _ptr(n), _ptr(sc), _ptr(ws), ws.nbytes, None))
    return [[int(x) for x in tok[b, : n[b]]] for b in range(B)], sc


def beam_grid(lib, lp: np.ndarray, lengths, blank: int, pad: int, space: int, settings, order=None):
    """settings: list of (width, prune, bias, tmin). -> per setting (ids, scores), one launch for all."""
    from silent_speech_decoder_b200._native import BeamSetting

    lp = np.ascontiguousarray(lp, dtype=np.float32)
    B, T, V = lp.shape
    n = len(settings)
    ln = np.asarray(lengths, dtype=np.int64)
    tok = np.full((n, B, max(T, 1)), -7, dtype=np.int32)
    cnt = np.full((n, B), -7, dtype=np.int32)
    sc = np.full((n, B), np.nan, dtype=np.float64)
    arr = (BeamSetting * max(n, 1))(*[BeamSetting(int(w), float(p), float(b), float(t)) for w, p, b, t in settings])
    wmax = max([int(s[0]) for s in settings], default=1)
    ws = np.zeros(lib.ssd_ctc_beam_workspace_bytes(B * max(n, 1), T, V, wmax) // 8 + 1, dtype=np.int64)
    od = None if order is None else np.asarray(order, dtype=np.int32)
    import ctypes as C

    lib.check(lib.ssd_ctc_beam_grid(_ptr(lp), B, T, V, T * V, V, _ptr(ln), _ptr(od), blank, pad, space, C.cast(arr, C.c_void_p), n,
                                    _ptr(tok), _ptr(cnt), _ptr(sc), _ptr(ws), ws.nbytes, None))
    return [([[int(x) for x in tok[i, b, : cnt[i, b]]] for b in range(B)], sc[i]) for i in range(n)]


def beam_lm(lib, lp: np.ndarray, lengths, blank: int, pad: int, space: int, tables, settings, order=None):
    """LM-fused search on host arrays: ``tables`` = decoding.lm.LanguageModelTables (numpy side only).
    -> per setting (ids, ranking scores)."""
    import ctypes as C

    from silent_speech_decoder_b200._native import BeamSetting, LmTables

    lp = np.ascontiguousarray(lp, dtype=np.float32)
    B, T, V = lp.shape
    n = len(settings)
    ln = np.asarray(lengths, dtype=np.int64)
    tok = np.full((n, B, max(T, 1)), -7, dtype=np.int32)
    cnt = np.full((n, B), -7, dtype=np.int32)
    sc = np.full((n, B), np.nan, dtype=np.float64)
    arr = (BeamSetting * max(n, 1))(*[BeamSetting(int(w), float(p), float(b), float(t)) for w, p, b, t in settings])
    wmax = max([int(s[0]) for s in settings], default=1)
    ws = np.zeros(lib.ssd_ctc_beam_lm_workspace_bytes(B * max(n, 1), T, V, wmax) // 8 + 1, dtype=np.int64)
    od = None if order is None else np.asarray(order, dtype=np.int32)
    held = [np.ascontiguousarray(a) for a in (tables.trie_child, tables.trie_word, tables.ngram_keys, tables.ngram_vals,
                                              tables.ngram_prob, tables.ngram_backoff)]
    st = LmTables(*[_ptr(a) for a in held], len(tables.ngram_keys) - 1, held[0].shape[0], held[0].shape[1], held[4].shape[0],
                  tables.order, tables.bos, tables.eos, tables.unk, tables.eos_rescore, tables.alpha, tables.beta, tables.unk_score_offset)
    lib.check(lib.ssd_ctc_beam_lm(_ptr(lp), B, T, V, T * V, V, _ptr(ln), _ptr(od), blank, pad, space, C.cast(arr, C.c_void_p), n,
                                  C.addressof(st), _ptr(tok), _ptr(cnt), _ptr(sc), _ptr(ws), ws.nbytes, None))
    return [([[int(x) for x in tok[i, b, : cnt[i, b]]] for b in range(B)], sc[i]) for i in range(n)]


def row_lse(lib, x: np.ndarray):
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, T, V = x.shape
    out = np.full((B, T), np.nan, dtype=np.float32)
    lib.check(lib.ssd_row_logsumexp(_ptr(x), B, T, V, T * V, V, _ptr(out), T, None))
    return out


def scale_if(lib, x: np.ndarray, g: float, expected: float):
    x = np.ascontiguousarray(x, dtype=np.float32).copy()
    gv = np.asarray([g], dtype=np.float32)
    lib.check(lib.ssd_scale_if(_ptr(x), x.size, _ptr(gv), float(expected), None))
    return x


def loss(lib, lp: np.ndarray, in_len, targets, tg_len, blank: int, zero_infinity: bool = True, mean_den: int = 0,
         want_grad: bool = True, grad_mul=None, with_fallbacks: bool = False, from_logits: bool = False):
    lp = np.ascontiguousarray(lp, dtype=np.float32)
    B, T, V = lp.shape
    il = np.asarray(in_len, dtype=np.int64)
    tl = np.asarray(tg_len, dtype=np.int64)
    tg = np.ascontiguousarray(targets, dtype=np.int64).reshape(B, -1)
    S = tg.shape[1]
    nll = np.full((B,), np.nan, dtype=np.float32)
    grad = np.full_like(lp, np.nan) if want_grad else None
    nbytes = (lib.ssd_ctc_loss_from_logits_workspace_bytes if from_logits else lib.ssd_ctc_loss_workspace_bytes)(B, T, V, S)
    ws = np.zeros(nbytes // 8 + 1, dtype=np.int64)
    gm = None if grad_mul is None else np.asarray([grad_mul], dtype=np.float32)
    entry = lib.ssd_ctc_loss_from_logits if from_logits else lib.ssd_ctc_loss
    lib.check(entry(_ptr(lp), B, T, V, T * V, V, _ptr(tg), S, S, _ptr(il), _ptr(tl), None, blank, int(zero_infinity),
                               mean_den, _ptr(gm), _ptr(nll), _ptr(grad), _ptr(ws), ws.nbytes, None))
    if with_fallbacks:  # utterances that took the exact fp64 path (None: the fast path does not cover the shape)
        off = lib.ssd_ctc_loss_fallback_counter_offset(B, T, V, S)
        nfb = None if off == 2 ** 64 - 1 else int(ws.view(np.uint32)[off // 4])
        if with_fallbacks == "flags":  # per-utterance reason bits (FastParams::flags follows the counters)
            return nll, grad, None if nfb is None else ws.view(np.int32)[(off - 4 + 256) // 4:(off - 4 + 256) // 4 + B].copy()
        return nll, grad, nfb
    return nll, grad


def edit_counts(lib, refs, hyps):
    """refs / hyps: lists of int sequences. -> (N,4) int32 [sub, del, ins, hits]."""
    N = len(refs)
    ro = np.zeros(N + 1, dtype=np.int64); ho = np.zeros(N + 1, dtype=np.int64)
    ro[1:] = np.cumsum([len(r) for r in refs]); ho[1:] = np.cumsum([len(h) for h in hyps])
    rt = np.array([x for r in refs for x in r] or [0], dtype=np.int32)
    ht = np.array([x for h in hyps for x in h] or [0], dtype=np.int32)
    out = np.full((max(N, 1), 4), -7, dtype=np.int32)
    lib.check(lib.ssd_edit_counts(_ptr(rt), _ptr(ro), _ptr(ht), _ptr(ho), N, max((len(r) for r in refs), default=0),
                                  max((len(h) for h in hyps), default=0), _ptr(out), None))
    return out[:N]
