"""ctypes binding of oracle/gotoh_oracle.c plus NumPy restatements of the
host-side steps (encoding, BGC-pair aggregation).  TEST INFRASTRUCTURE.

The substitution table below is the oracle's *own* transcription of BLOSUM62
(NCBI table, alphabet/order of Biopython's ``substitution_matrices.load
("BLOSUM62")``: ``ARNDCQEGHILKMFPSTWYVBZX*``).  The product
carries a separate copy; ``tests/test_oracle.py`` checks they agree.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

MODE_LOCAL = 0
MODE_GLOBAL = 1

ALPHABET = "ARNDCQEGHILKMFPSTWYVBZX*"

_B62_TEXT = """
 4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0 -2 -1  0 -4
-1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3 -1  0 -1 -4
-2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3  3  0 -1 -4
-2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3  4  1 -1 -4
 0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -3 -3 -2 -4
-1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2  0  3 -1 -4
-1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
 0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3 -1 -2 -1 -4
-2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3  0  0 -1 -4
-1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3 -3 -3 -1 -4
-1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1 -4 -3 -1 -4
-1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2  0  1 -1 -4
-1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1 -3 -1 -1 -4
-2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1 -3 -3 -1 -4
-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2 -2 -1 -2 -4
 1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2  0  0  0 -4
 0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0 -1 -1  0 -4
-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3 -4 -3 -2 -4
-2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1 -3 -2 -1 -4
 0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4 -3 -2 -1 -4
-2 -1  3  4 -3  0  1 -1  0 -3 -4  0 -3 -3 -2  0 -1 -4 -3 -3  4  1 -1 -4
-1  0  0  1 -3  3  4 -2  0 -3 -3  1 -1 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
 0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2  0  0 -2 -1 -1 -1 -1 -1 -4
-4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4  1
"""
BLOSUM62 = np.array(_B62_TEXT.split(), dtype=np.int8).reshape(24, 24)
BLOSUM62.setflags(write=False)

_HERE = Path(__file__).resolve().parent
_LIB = None


def load():
    """Build (if stale) and load oracle/_build/libbgco.so."""
    global _LIB
    if _LIB is not None:
        return _LIB
    src = _HERE / "gotoh_oracle.c"
    so = _HERE / "_build" / "libbgco.so"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    try:
        lib = ctypes.CDLL(str(so))
    except OSError:
        # a stale .so built on another host: rebuild in place
        subprocess.run(["make", "-C", str(_HERE), "-s", "-B"], check=True)
        lib = ctypes.CDLL(str(so))
    u8p = ctypes.POINTER(ctypes.c_uint8)
    i8p = ctypes.POINTER(ctypes.c_int8)
    i32p = ctypes.POINTER(ctypes.c_int32)
    i64p = ctypes.POINTER(ctypes.c_int64)
    i32 = ctypes.c_int32
    for name in ("bgco_score", "bgco_score_3state"):
        fn = getattr(lib, name)
        fn.argtypes = [u8p, i32, u8p, i32, i8p, i32, i32, i32]
        fn.restype = i32
    lib.bgco_all_pairs.argtypes = [u8p, i64p, i32, i8p, i32, i32, i32, i32p, i32]
    lib.bgco_all_pairs.restype = None
    lib.bgco_pair_list.argtypes = [u8p, i64p, i32, i32p, i32p, ctypes.c_int64, i8p, i32, i32,
                                   i32, i32p, i32]
    lib.bgco_pair_list.restype = None
    lib.bgco_stats_list.argtypes = [u8p, i64p, i32, i32p, i32p, ctypes.c_int64, i8p, i32, i32, i32,
                                    ctypes.c_void_p, i32]
    lib.bgco_stats_list.restype = i32
    lib.bgco_num_threads.argtypes = []
    lib.bgco_num_threads.restype = i32
    _LIB = lib
    return lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


_LUT = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(ALPHABET):
    _LUT[ord(_c)] = _i


def encode(seq: str) -> np.ndarray:
    """Upper-case ASCII -> codes 0..23.  Mirrors Biopython: a zero-length
    sequence or a letter outside the matrix alphabet raises ValueError
    (SURVEY.md App. C)."""
    if len(seq) == 0:
        raise ValueError("sequence has zero length")
    raw = np.frombuffer(seq.encode("ascii"), dtype=np.uint8)
    codes = _LUT[raw]
    if (codes == 255).any():
        bad = sorted({chr(c) for c in raw[codes == 255]})
        raise ValueError(f"sequence contains letters not in the alphabet: {bad}")
    return codes


def _pack(seqs):
    enc = [encode(s) if isinstance(s, str) else np.ascontiguousarray(s, dtype=np.uint8) for s in seqs]
    offsets = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum([len(e) for e in enc], out=offsets[1:])
    codes = np.concatenate(enc) if enc else np.zeros(0, np.uint8)
    return np.ascontiguousarray(codes), offsets


def score(a, b, *, mode=MODE_LOCAL, sub=BLOSUM62, open_gap=-12, extend_gap=-1) -> int:
    lib = load()
    ca, cb = (encode(a) if isinstance(a, str) else a), (encode(b) if isinstance(b, str) else b)
    sub = np.ascontiguousarray(sub, dtype=np.int8)
    return int(lib.bgco_score(_p(ca, ctypes.c_uint8), len(ca), _p(cb, ctypes.c_uint8), len(cb),
                              _p(sub, ctypes.c_int8), open_gap, extend_gap, mode))


def score_3state(a, b, *, mode=MODE_LOCAL, sub=BLOSUM62, open_gap=-12, extend_gap=-1) -> int:
    lib = load()
    ca, cb = (encode(a) if isinstance(a, str) else a), (encode(b) if isinstance(b, str) else b)
    sub = np.ascontiguousarray(sub, dtype=np.int8)
    return int(lib.bgco_score_3state(_p(ca, ctypes.c_uint8), len(ca), _p(cb, ctypes.c_uint8),
                                     len(cb), _p(sub, ctypes.c_int8), open_gap, extend_gap, mode))


def all_pairs(seqs, *, mode=MODE_LOCAL, sub=BLOSUM62, open_gap=-12, extend_gap=-1,
              nthreads=0) -> np.ndarray:
    """Full symmetric n x n int32 score matrix (diagonal = self scores)."""
    lib = load()
    codes, offsets = _pack(seqs)
    n = len(offsets) - 1
    out = np.zeros((n, n), dtype=np.int32)
    sub = np.ascontiguousarray(sub, dtype=np.int8)
    lib.bgco_all_pairs(_p(codes, ctypes.c_uint8), _p(offsets, ctypes.c_int64), n,
                       _p(sub, ctypes.c_int8), open_gap, extend_gap, mode,
                       _p(out, ctypes.c_int32), nthreads)
    return out


def pair_list(seqs, pi, pj, *, mode=MODE_LOCAL, sub=BLOSUM62, open_gap=-12, extend_gap=-1,
              nthreads=0) -> np.ndarray:
    lib = load()
    codes, offsets = _pack(seqs)
    pi = np.ascontiguousarray(pi, dtype=np.int32)
    pj = np.ascontiguousarray(pj, dtype=np.int32)
    out = np.zeros(len(pi), dtype=np.int32)
    sub = np.ascontiguousarray(sub, dtype=np.int8)
    lib.bgco_pair_list(_p(codes, ctypes.c_uint8), _p(offsets, ctypes.c_int64), len(offsets) - 1,
                       _p(pi, ctypes.c_int32), _p(pj, ctypes.c_int32), len(pi),
                       _p(sub, ctypes.c_int8), open_gap, extend_gap, mode,
                       _p(out, ctypes.c_int32), nthreads)
    return out


STATS_DTYPE = np.dtype([(k, np.int32) for k in ("score", "identities", "columns", "q_start", "q_end",
                                                "s_start", "s_end", "gaps")])


def stats_list(seqs, pi, pj, *, mode=MODE_LOCAL, sub=BLOSUM62, open_gap=-12, extend_gap=-1,
               nthreads=0) -> np.ndarray:
    """Alignment statistics without traceback for an explicit pair list (query = seqs[pi[p]],
    subject = seqs[pj[p]]); structured array with the fields of ``bgco_stats``."""
    lib = load()
    codes, offsets = _pack(seqs)
    pi = np.ascontiguousarray(pi, dtype=np.int32)
    pj = np.ascontiguousarray(pj, dtype=np.int32)
    out = np.zeros(len(pi), dtype=STATS_DTYPE)
    sub = np.ascontiguousarray(sub, dtype=np.int8)
    rc = lib.bgco_stats_list(_p(codes, ctypes.c_uint8), _p(offsets, ctypes.c_int64), len(offsets) - 1,
                             _p(pi, ctypes.c_int32), _p(pj, ctypes.c_int32), len(pi),
                             _p(sub, ctypes.c_int8), open_gap, extend_gap, mode,
                             out.ctypes.data_as(ctypes.c_void_p), nthreads)
    if rc != 0:
        raise ValueError("stats: sequence lengths must be in 1..32767")
    return out


def num_threads() -> int:
    return int(load().bgco_num_threads())


def bgc_similarity_np(scores, bgc_of_dom, type_of_dom, n_bgc, same_type_only=True):
    """Domain-pair -> BGC-pair aggregation (this build's
    definition — the reference only states the intent, BGClib/Readme.md:16-18).

    rowbest[d,b] = max{ s(d,e) : e in b, type(e)==type(d) } (0 if none)
    best[a,b]    = sum_{d in a} rowbest[d,b]          (int64)
    self[a]      = sum_{d in a} s(d,d)                (int64)
    sim[a,b]     = (best[a,b]+best[b,a]) / (self[a]+self[b])  (one f64 divide;
                   0/0 -> 0.0; sim[a,a] = 1.0 when self[a] > 0)
    Deliberately written as plain loops over domains.
    """
    scores = np.asarray(scores)
    bgc_of_dom = np.asarray(bgc_of_dom)
    type_of_dom = np.asarray(type_of_dom)
    n = scores.shape[0]
    rowbest = np.zeros((n, n_bgc), dtype=np.int64)
    for d in range(n):
        for e in range(n):
            if same_type_only and type_of_dom[d] != type_of_dom[e]:
                continue
            b = bgc_of_dom[e]
            if scores[d, e] > rowbest[d, b]:
                rowbest[d, b] = scores[d, e]
    best = np.zeros((n_bgc, n_bgc), dtype=np.int64)
    selfs = np.zeros(n_bgc, dtype=np.int64)
    for d in range(n):
        best[bgc_of_dom[d], :] += rowbest[d, :]
        selfs[bgc_of_dom[d]] += int(scores[d, d])
    sim = np.zeros((n_bgc, n_bgc), dtype=np.float64)
    for a in range(n_bgc):
        for b in range(n_bgc):
            den = selfs[a] + selfs[b]
            if den > 0:
                sim[a, b] = float(best[a, b] + best[b, a]) / float(den)
        if selfs[a] > 0:
            sim[a, a] = 1.0
    return sim, best, selfs, rowbest
