"""B200-native CTC hot path (loss, greedy decode, prefix beam search).

Drop-in for the two reference call sites of ajshul/silent-speech-decoder:

* ``models.losses.DistillationCTCLoss``  (reference ``src/models/losses.py:19-105``)
* ``decoding.ctc.build_decoder``         (reference ``src/decoding/ctc.py:102-122``)

Everything numeric runs in hand-written sm_100a CUDA kernels behind the C ABI
declared in ``include/ssd_ctc.h``; there is no CPU fallback.
"""

__version__ = "0.1.0"

__all__ = ["__version__", "install_as_src"]


def install_as_src() -> None:
    """Alias this package's modules under the reference's ``src.*`` names.

    After this call ``from src.models.losses import DistillationCTCLoss`` and
    ``from src.decoding.ctc import build_decoder`` resolve to the native path,
    which is how an unmodified reference ``train.py`` / ``evaluate.py`` picks it up
    (see INTEGRATION.md).  Modules of the reference that we do not replace are
    left alone if a real ``src`` package is already importable.
    """
    import importlib
    import sys
    import types

    pkg = __name__.split(".")[0]
    pairs = {
        "src.data.vocab": f"{pkg}.data.vocab",
        "src.decoding.ctc": f"{pkg}.decoding.ctc",
        "src.models.losses": f"{pkg}.models.losses",
    }
    for parent in ("src", "src.data", "src.decoding", "src.models"):
        if parent not in sys.modules:
            try:
                importlib.import_module(parent)
            except Exception:
                mod = types.ModuleType(parent)
                mod.__path__ = []  # type: ignore[attr-defined]
                sys.modules[parent] = mod
    for alias, real in pairs.items():
        mod = importlib.import_module(real)
        sys.modules[alias] = mod
        parent, _, leaf = alias.rpartition(".")
        setattr(sys.modules[parent], leaf, mod)
