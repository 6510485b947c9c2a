"""Pure-Python restatement of Biopython 1.83 ``PairwiseAligner.score()`` for
affine gaps (Gotoh) — TEST INFRASTRUCTURE, small cases only.

Written independently of oracle/gotoh_oracle.c (different recurrence: three
explicit states M / Ix / Iy with Ix<->Iy transitions, floats like Biopython,
dictionary substitution lookup on letters rather than code tables) so that
agreement between the two is evidence, not tautology.  SURVEY.md App. C.
PARITY UNPINNED BY THE REFERENCE (no aligner in jorgecnavarrom/BGClib).
"""
from __future__ import annotations

from .cbind import ALPHABET, BLOSUM62

_SUB62 = {(a, b): float(BLOSUM62[i, j]) for i, a in enumerate(ALPHABET) for j, b in enumerate(ALPHABET)}
_NEG = float("-inf")


def score(a: str, b: str, mode: str = "local", open_gap_score: float = -12.0,
          extend_gap_score: float = -1.0, sub=None) -> float:
    """``sub``: optional 24x24 matrix in ALPHABET order (default BLOSUM62)."""
    _SUB = _SUB62 if sub is None else \
        {(x, y): float(sub[i][j]) for i, x in enumerate(ALPHABET) for j, y in enumerate(ALPHABET)}
    if not a or not b:
        raise ValueError("sequence has zero length")
    for ch in a + b:
        if ch not in ALPHABET:
            raise ValueError(f"sequence contains letters not in the alphabet: {ch!r}")
    local = mode == "local"
    o, e = float(open_gap_score), float(extend_gap_score)
    n, m = len(a), len(b)
    # previous-row states
    if local:
        M0 = [0.0] * (m + 1); X0 = [0.0] * (m + 1); Y0 = [0.0] * (m + 1)
    else:
        M0 = [_NEG] * (m + 1); M0[0] = 0.0
        X0 = [_NEG] + [o + (j - 1) * e for j in range(1, m + 1)]   # gaps along b's axis
        Y0 = [_NEG] * (m + 1)
    best = 0.0
    for i in range(1, n + 1):
        if local:
            M1 = [0.0] * (m + 1); X1 = [0.0] * (m + 1); Y1 = [0.0] * (m + 1)
        else:
            M1 = [_NEG] * (m + 1); X1 = [_NEG] * (m + 1); Y1 = [_NEG] * (m + 1)
            Y1[0] = o + (i - 1) * e
        ai = a[i - 1]
        for j in range(1, m + 1):
            mm = max(M0[j - 1], X0[j - 1], Y0[j - 1]) + _SUB[ai, b[j - 1]]
            xx = max(M1[j - 1] + o, X1[j - 1] + e, Y1[j - 1] + o)
            yy = max(M0[j] + o, Y0[j] + e, X0[j] + o)
            if local:
                mm = max(mm, 0.0); xx = max(xx, 0.0); yy = max(yy, 0.0)
                if mm > best:
                    best = mm
            M1[j], X1[j], Y1[j] = mm, xx, yy
        M0, X0, Y0 = M1, X1, Y1
    return best if local else max(M0[m], X0[m], Y0[m])
