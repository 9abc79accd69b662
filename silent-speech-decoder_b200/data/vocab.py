"""Character vocabulary: id <-> token table with the three special ids.

Mirrors the interface of reference ``src/data/vocab.py:11-57`` (same field
names, same ``from_json``/``to_json``/``encode``/``decode``/``size``) because the
decoders are handed a ``Vocab`` by their callers (``evaluate.py:212``).
"""

from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Iterable, List

_DEFAULT_JSON = Path(__file__).resolve().parent.parent / "configs" / "vocab.json"


@dataclass
class Vocab:
    tokens: List[str]
    token_to_id: Dict[str, int]
    pad_id: int
    blank_id: int
    unk_id: int

    @classmethod
    def from_json(cls, path: Path) -> "Vocab":
        spec = json.loads(Path(path).read_text())
        toks = list(spec["tokens"])
        return cls(
            tokens=toks,
            token_to_id={t: k for k, t in enumerate(toks)},
            pad_id=int(spec.get("pad_id", 0)),
            blank_id=int(spec.get("blank_id", 1)),
            unk_id=int(spec.get("unk_id", 2)),
        )

    @classmethod
    def default(cls) -> "Vocab":
        """The 48-token vocabulary of reference ``configs/vocab.json``."""
        return cls.from_json(_DEFAULT_JSON)

    def to_json(self, path: Path) -> None:
        Path(path).write_text(
            json.dumps(
                {"tokens": self.tokens, "pad_id": self.pad_id, "blank_id": self.blank_id, "unk_id": self.unk_id},
                indent=2,
            )
        )

    def encode(self, text: str) -> List[int]:
        # unknown characters map to unk; never produces pad/blank (vocab.py:41-42)
        lut, unk = self.token_to_id, self.unk_id
        return [lut.get(ch, unk) for ch in text.lower()]

    def decode(self, ids: Iterable[int], skip_blank: bool = True) -> str:
        # drops blank (optionally) and pad, ignores out-of-range ids (vocab.py:44-53)
        n = len(self.tokens)
        out: List[str] = []
        for i in ids:
            if (skip_blank and i == self.blank_id) or i == self.pad_id:
                continue
            if 0 <= i < n:
                out.append(self.tokens[i])
        return "".join(out)

    @property
    def size(self) -> int:
        return len(self.tokens)
