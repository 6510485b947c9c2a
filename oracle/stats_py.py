"""Pure-Python restatements of the alignment-statistics definition (oracle/gotoh_oracle.c,
"Alignment statistics without traceback") — TEST INFRASTRUCTURE, small cases only.

Two formulations, both independent of the C code's packed keys and of its H/E/F recurrence:

* ``stats_3state``: M / Ix / Iy tuple DP on Python tuples (score, identities, -columns,
  q_start, s_start), letters looked up in a dictionary;
* ``stats_bruteforce``: explicit enumeration of EVERY monotone path between every two DP
  corners (tiny inputs), each scored from its own column list.

Both return the lexicographic maximum of
(score, identities, -columns, q_start, s_start, -s_end, -q_end), the empty local alignment
being (0, 0, 0) with all coordinates -1.  The reference (jorgecnavarrom/BGClib) defines no
such quantity: PARITY UNPINNED BY THE REFERENCE.
"""
from __future__ import annotations

from .cbind import ALPHABET, BLOSUM62

_SUB = {(a, b): int(BLOSUM62[i, j]) for i, a in enumerate(ALPHABET) for j, b in enumerate(ALPHABET)}
_NEG = (-(10 ** 9), 0, 0, 0, 0)


def _out(score, ident, cols, qs, ss, qe, se):
    if cols == 0:
        return dict(score=0, identities=0, columns=0, q_start=-1, q_end=-1, s_start=-1, s_end=-1, gaps=0)
    return dict(score=score, identities=ident, columns=cols, q_start=qs, q_end=qe, s_start=ss, s_end=se,
                gaps=2 * cols - (qe - qs) - (se - ss))


def stats_3state(a: str, b: str, mode: str = "local", open_gap_score: int = -12, extend_gap_score: int = -1):
    local = mode == "local"
    o, e = int(open_gap_score), int(extend_gap_score)
    n, m = len(a), len(b)

    def add(t, s, ident):
        return (t[0] + s, t[1] + ident, t[2] - 1, t[3], t[4])

    def origin(i, j):                       # the empty alignment sitting at DP corner (i, j)
        return (0, 0, 0, i, j)

    M = [[_NEG] * (m + 1) for _ in range(n + 1)]
    X = [[_NEG] * (m + 1) for _ in range(n + 1)]     # ends with a gap along b's axis (consumes b_j)
    Y = [[_NEG] * (m + 1) for _ in range(n + 1)]     # ends with a gap along a's axis (consumes a_i)
    Z = [[_NEG] * (m + 1) for _ in range(n + 1)]     # corners an alignment may start from
    for i in range(n + 1):
        for j in range(m + 1):
            if local or (i == 0 and j == 0):
                Z[i][j] = origin(i, j)
    best, best_end = (0, 0, 0, 0, 0), None
    for i in range(n + 1):
        for j in range(m + 1):
            if i > 0 and j > 0:
                prev = max(M[i - 1][j - 1], X[i - 1][j - 1], Y[i - 1][j - 1], Z[i - 1][j - 1])
                M[i][j] = add(prev, _SUB[a[i - 1], b[j - 1]], int(a[i - 1] == b[j - 1]))
            if j > 0:
                X[i][j] = max(add(M[i][j - 1], o, 0), add(X[i][j - 1], e, 0), add(Y[i][j - 1], o, 0),
                              add(Z[i][j - 1], o, 0))
            if i > 0:
                Y[i][j] = max(add(M[i - 1][j], o, 0), add(Y[i - 1][j], e, 0), add(X[i - 1][j], o, 0),
                              add(Z[i - 1][j], o, 0))
            if local:
                for t in (M[i][j], X[i][j], Y[i][j]):
                    if t[2] == 0:
                        continue
                    cand = (t, -j, -i)
                    if best_end is None and t[:3] > (0, 0, 0) or best_end is not None and cand > (best, -best_end[1], -best_end[0]):
                        best, best_end = t, (i, j)
    if not local:
        best, best_end = max(M[n][m], X[n][m], Y[n][m]), (n, m)
    if best_end is None:
        return _out(0, 0, 0, -1, -1, -1, -1)
    return _out(best[0], best[1], -best[2], best[3], best[4], best_end[0], best_end[1])


def stats_bruteforce(a: str, b: str, mode: str = "local", open_gap_score: int = -12, extend_gap_score: int = -1):
    """Every monotone path, scored column by column.  O(exponential): len(a), len(b) <= 5."""
    local = mode == "local"
    o, e = int(open_gap_score), int(extend_gap_score)
    n, m = len(a), len(b)
    best = [None]

    def consider(score, ident, cols, qs, ss, qe, se):
        key = (score, ident, -cols, qs, ss, -se, -qe)
        if best[0] is None or key > best[0]:
            best[0] = key

    def walk(i, j, last, score, ident, cols, qs, ss):
        # an alignment may end at any corner (local) or only at (n, m) (global)
        if cols > 0 and (local or (i == n and j == m)):
            consider(score, ident, cols, qs, ss, i, j)
        if i < n and j < m:
            walk(i + 1, j + 1, "M", score + _SUB[a[i], b[j]], ident + int(a[i] == b[j]), cols + 1, qs, ss)
        if j < m:
            walk(i, j + 1, "X", score + (e if last == "X" else o), ident, cols + 1, qs, ss)
        if i < n:
            walk(i + 1, j, "Y", score + (e if last == "Y" else o), ident, cols + 1, qs, ss)

    starts = [(i, j) for i in range(n + 1) for j in range(m + 1)] if local else [(0, 0)]
    for (i, j) in starts:
        walk(i, j, "", 0, 0, 0, i, j)
    k = best[0]
    if k is None or (local and k[:3] <= (0, 0, 0)):
        return _out(0, 0, 0, -1, -1, -1, -1)
    return _out(k[0], k[1], -k[2], k[3], k[4], -k[6], -k[5])
