"""ctypes binding of the C ABI in include/ssd_ctc.h.

The product path loads the nvcc-built ``libssd_ctc.so`` that sits next to this
file and fails loudly if it is missing -- there is no CPU or PyTorch fallback.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

ABI_VERSION = 3
LIB_PATH = Path(__file__).resolve().parent / "libssd_ctc.so"

SSD_OK, SSD_ERR_BAD_ARG, SSD_ERR_UNSUPPORTED, SSD_ERR_WORKSPACE, SSD_ERR_CUDA = range(5)

_P, _I64, _I32, _F32, _SZ = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t

SIGNATURES = {
    "ssd_abi_version": (C.c_int, []),
    "ssd_last_error": (C.c_char_p, []),
    "ssd_ctc_greedy": (C.c_int, [_P, _I64, _I64, _I64, _I64, _I64, _P, _I32, _F32, _P, _P, _P]),
    "ssd_ctc_beam_workspace_bytes": (_SZ, [_I64, _I64, _I64, _I32]),
    "ssd_ctc_beam": (C.c_int, [_P, _I64, _I64, _I64, _I64, _I64, _P, _P, _I32, _I32, _I32, _F32, _I32, _F32, _F32, _P, _P, _P, _P, _SZ, _P]),
    "ssd_ctc_beam_grid": (C.c_int, [_P, _I64, _I64, _I64, _I64, _I64, _P, _P, _I32, _I32, _I32, _P, _I32, _P, _P, _P, _P, _SZ, _P]),
    "ssd_lm_hash": (C.c_uint64, [C.c_uint64]),
    "ssd_ctc_beam_lm_workspace_bytes": (_SZ, [_I64, _I64, _I64, _I32]),
    "ssd_ctc_beam_lm": (C.c_int, [_P, _I64, _I64, _I64, _I64, _I64, _P, _P, _I32, _I32, _I32, _P, _I32, _P, _P, _P, _P, _P, _SZ, _P]),
    "ssd_ctc_loss_from_logits_workspace_bytes": (_SZ, [_I64, _I64, _I64, _I64]),
    "ssd_ctc_loss_from_logits": (C.c_int, [_P, _I64, _I64, _I64, _I64, _I64, _P, _I64, _I64, _P, _P, _P, _I32, _I32, _I64, _P, _P, _P, _P, _SZ, _P]),
    "ssd_row_logsumexp": (C.c_int, [_P, _I64, _I64, _I64, _I64, _I64, _P, _I64, _P]),
    "ssd_scale_if": (C.c_int, [_P, _I64, _P, _F32, _P]),
    "ssd_ctc_loss_parts": (C.c_int, [_P, _P, _I64, _P, _P]),
    "ssd_ctc_loss_workspace_bytes": (_SZ, [_I64, _I64, _I64, _I64]),
    "ssd_ctc_loss_fallback_counter_offset": (_SZ, [_I64, _I64, _I64, _I64]),
    "ssd_edit_counts": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _P, _P]),
    "ssd_ctc_loss": (C.c_int, [_P, _I64, _I64, _I64, _I64, _I64, _P, _I64, _I64, _P, _P, _P, _I32, _I32, _I64, _P, _P, _P, _P, _SZ, _P]),
}


class BeamSetting(C.Structure):
    """``ssd_beam_setting`` of include/ssd_ctc.h."""

    _fields_ = [("beam_width", C.c_int32), ("beam_prune_logp", C.c_float), ("blank_bias", C.c_float), ("token_min_logp", C.c_float)]


class LmTables(C.Structure):
    """``ssd_lm_tables`` of include/ssd_ctc.h (device pointers + scalars)."""

    _fields_ = [
        ("trie_child", C.c_void_p), ("trie_word", C.c_void_p), ("ngram_keys", C.c_void_p), ("ngram_vals", C.c_void_p),
        ("ngram_prob", C.c_void_p), ("ngram_backoff", C.c_void_p), ("ngram_mask", C.c_uint64),
        ("n_trie_nodes", C.c_int32), ("n_labels", C.c_int32), ("n_ngrams", C.c_int32), ("order", C.c_int32),
        ("bos", C.c_int32), ("eos", C.c_int32), ("unk", C.c_int32), ("eos_rescore", C.c_int32),
        ("alpha", C.c_double), ("beta", C.c_double), ("unk_score_offset", C.c_double),
    ]


class NativeError(RuntimeError):
    pass


class Library:
    """A loaded libssd_ctc with typed entry points; raises on non-zero status like the
    reference's dependencies do (ValueError for bad arguments, RuntimeError otherwise)."""

    def __init__(self, path: Path):
        self.path = Path(path)
        self.lib = C.CDLL(str(path))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.lib, name)  # AttributeError if the symbol is missing
            fn.restype, fn.argtypes = res, args
        got = self.lib.ssd_abi_version()
        if got != ABI_VERSION:
            raise NativeError(f"{path}: ABI version {got}, expected {ABI_VERSION}; rebuild with build.py")

    def check(self, status: int) -> None:
        if status == SSD_OK:
            return
        msg = (self.lib.ssd_last_error() or b"").decode("utf-8", "replace")
        if status in (SSD_ERR_BAD_ARG,):
            raise ValueError(msg)
        raise NativeError(f"libssd_ctc error {status}: {msg}")

    def __getattr__(self, name):
        return getattr(self.lib, name)


_LIB: Optional[Library] = None


def load() -> Library:
    """The product library. No fallback: a missing build is an error."""
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            raise NativeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built. "
                "Run `python -m silent_speech_decoder_b200.build` (needs nvcc); there is no CPU fallback."
            )
        _LIB = Library(LIB_PATH)
    return _LIB
