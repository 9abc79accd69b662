"""Import alias for the on-disk package directory ``silent-speech-decoder_b200/``.

The task names the package directory with a hyphen, which Python cannot import
directly.  This shim makes ``import silent_speech_decoder_b200`` resolve its
submodules from that directory, so the source lives in exactly one place.
"""

from __future__ import annotations

import os as _os

_HERE = _os.path.dirname(_os.path.abspath(__file__))
_REAL = _os.path.join(_os.path.dirname(_HERE), "silent-speech-decoder_b200")
if not _os.path.isdir(_REAL):  # pragma: no cover - broken checkout
    raise ImportError(f"package directory missing: {_REAL}")
__path__.insert(0, _REAL)  # type: ignore[name-defined]

with open(_os.path.join(_REAL, "__init__.py"), "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
