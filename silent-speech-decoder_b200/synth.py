"""Seeded synthetic inputs for the CTC hot path.

``flat``   : log_softmax of white noise -- what a random-init CTC head emits; nearly
             every label passes the beam's per-frame -5 filter (worst case).
``peaked`` : a random transcript laid on a random monotone CTC alignment with
             +8 logit on the aligned label -- what a trained model emits.
All generators are deterministic functions of ``seed`` and run on CPU so the
oracle and the GPU path see bit-identical inputs.
"""

from __future__ import annotations

from typing import Tuple

import torch

V_DEFAULT = 48
BLANK_DEFAULT = 1
SPACE_DEFAULT = 3


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def flat(B: int, T: int, seed: int, V: int = V_DEFAULT) -> torch.Tensor:
    """(B,T,V) fp32 log-probs, near-uniform."""
    x = torch.randn(B, T, V, generator=_gen(seed), dtype=torch.float32)
    return torch.log_softmax(x, dim=-1)


def peaked(
    B: int,
    T: int,
    seed: int,
    V: int = V_DEFAULT,
    blank: int = BLANK_DEFAULT,
    space: int = SPACE_DEFAULT,
    boost: float = 8.0,
) -> torch.Tensor:
    """(B,T,V) fp32 log-probs concentrated on a random valid CTC alignment."""
    g = _gen(seed)
    x = torch.randn(B, T, V, generator=g, dtype=torch.float32)
    n_tok = max(1, T // 5)
    for b in range(B):
        # transcript over ids 3..V-1, space with prob ~0.17, no leading/double space
        ids = torch.randint(4, V, (n_tok,), generator=g)
        is_sp = torch.rand(n_tok, generator=g) < 0.17
        toks = []
        for k in range(n_tok):
            t = space if bool(is_sp[k]) and toks and toks[-1] != space else int(ids[k])
            toks.append(t)
        # extended sequence with blanks where needed between repeats, else random
        ext = []
        for k, t in enumerate(toks):
            if k > 0 and (toks[k - 1] == t or float(torch.rand((), generator=g)) < 0.5):
                ext.append(blank)
            ext.append(t)
        ext = ext[:T]
        # monotone alignment: distribute the remaining frames as repeats
        extra = T - len(ext)
        cuts = torch.sort(torch.randint(0, len(ext), (extra,), generator=g)).values
        reps = torch.bincount(cuts, minlength=len(ext)) + 1
        align = torch.repeat_interleave(torch.tensor(ext), reps)[:T]
        x[b, torch.arange(T), align] += boost
    return torch.log_softmax(x, dim=-1)


def targets(
    B: int, S: int, seed: int, V: int = V_DEFAULT, lo: int = 2, s_min: int | None = None
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Zero-padded (B,S) int64 label matrix with ids in [lo,V) and lengths ~U[S/4,S]."""
    g = _gen(seed)
    s_min = max(1, S // 4) if s_min is None else s_min
    lens = torch.randint(s_min, S + 1, (B,), generator=g, dtype=torch.int64)
    tg = torch.randint(lo, V, (B, S), generator=g, dtype=torch.int64)
    mask = torch.arange(S)[None, :] < lens[:, None]
    return tg * mask, lens


def ragged_lengths(B: int, T: int, seed: int) -> torch.Tensor:
    """(B,) int64 frame counts ~U[T/2,T]."""
    return torch.randint(max(1, T // 2), T + 1, (B,), generator=_gen(seed), dtype=torch.int64)
