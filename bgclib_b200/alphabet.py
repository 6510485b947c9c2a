"""Residue alphabet and the default substitution matrix of the engine.

BLOSUM62 in the order Biopython's ``substitution_matrices.load("BLOSUM62")``
uses (``ARNDCQEGHILKMFPSTWYVBZX*``) — the matrix Biopython's own ``blastp``
preset pairs with open -12 / extend -1.  The
reference validates no residues (TODO at BGClib/BGClib.py:1905), so ``B Z X *``
can reach the engine and score per this table; any other letter is an error,
as it is in ``Bio.Align.PairwiseAligner``.
"""
from __future__ import annotations

import numpy as np

ALPHABET = "ARNDCQEGHILKMFPSTWYVBZX*"
NSYM = 24
PAD_CODE = 24

_ROWS = (
    "A  4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0 -2 -1  0 -4",
    "R -1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3 -1  0 -1 -4",
    "N -2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3  3  0 -1 -4",
    "D -2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3  4  1 -1 -4",
    "C  0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -3 -3 -2 -4",
    "Q -1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2  0  3 -1 -4",
    "E -1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4",
    "G  0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3 -1 -2 -1 -4",
    "H -2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3  0  0 -1 -4",
    "I -1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3 -3 -3 -1 -4",
    "L -1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1 -4 -3 -1 -4",
    "K -1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2  0  1 -1 -4",
    "M -1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1 -3 -1 -1 -4",
    "F -2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1 -3 -3 -1 -4",
    "P -1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2 -2 -1 -2 -4",
    "S  1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2  0  0  0 -4",
    "T  0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0 -1 -1  0 -4",
    "W -3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3 -4 -3 -2 -4",
    "Y -2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1 -3 -2 -1 -4",
    "V  0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4 -3 -2 -1 -4",
    "B -2 -1  3  4 -3  0  1 -1  0 -3 -4  0 -3 -3 -2  0 -1 -4 -3 -3  4  1 -1 -4",
    "Z -1  0  0  1 -3  3  4 -2  0 -3 -3  1 -1 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4",
    "X  0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2  0  0 -2 -1 -1 -1 -1 -1 -4",
    "*  -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4  1",
)


def _parse():
    m = np.zeros((NSYM, NSYM), dtype=np.int8)
    for i, row in enumerate(_ROWS):
        parts = row.split()
        assert parts[0] == ALPHABET[i]
        m[i] = [int(x) for x in parts[1:]]
    assert (m == m.T).all()
    m.setflags(write=False)
    return m


BLOSUM62 = _parse()

MATRICES = {"BLOSUM62": BLOSUM62}


def resolve_matrix(matrix) -> np.ndarray:
    """Name or 24x24 integer array (alphabet order above) -> int8 matrix."""
    if isinstance(matrix, str):
        try:
            return MATRICES[matrix.upper()]
        except KeyError:
            raise ValueError(f"unknown substitution matrix {matrix!r}; known: {sorted(MATRICES)}") from None
    m = np.asarray(matrix)
    if m.shape != (NSYM, NSYM):
        raise ValueError(f"substitution matrix must be {NSYM}x{NSYM} in order {ALPHABET!r}")
    if not np.issubdtype(m.dtype, np.integer):
        if not np.all(m == np.round(m)):
            raise ValueError("substitution matrix must hold integers (scores are computed exactly)")
    m = m.astype(np.int64)
    if (np.abs(m) > 127).any():
        raise ValueError("substitution scores must fit int8")
    if not (m == m.T).all():
        raise ValueError("substitution matrix must be symmetric")
    return np.ascontiguousarray(m, dtype=np.int8)
