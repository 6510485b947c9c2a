"""Known answers from the literature — a pin for the oracle (and through it for the CUDA path)
that neither the surveyor nor this build computed.  Each case: (name, seq a, seq b, mode,
24x24 matrix, open, extend, expected score, source).  Gap convention of the engine: a gap of
length k scores open + (k-1)*extend, so a linear gap penalty d is open = extend = -d and an
affine penalty W_k = u + v*k is open = -(u+v), extend = -v."""
import numpy as np

ALPHABET = "ARNDCQEGHILKMFPSTWYVBZX*"


def _matrix(entries, default=0, diagonal=None):
    m = np.full((24, 24), default, dtype=np.int8)
    if diagonal is not None:
        np.fill_diagonal(m, diagonal)
    for (a, b), v in entries.items():
        m[ALPHABET.index(a), ALPHABET.index(b)] = v
        m[ALPHABET.index(b), ALPHABET.index(a)] = v
    return m


# Durbin, Eddy, Krogh & Mitchison, "Biological sequence analysis" (CUP 1998), section 2.3,
# Figure 2.2: the BLOSUM50 scores of HEAGAWGHEE against PAWHEAE (only these entries of the
# matrix can be met by the two sequences; the rest of the 24x24 table is left 0).
DURBIN_BLOSUM50 = _matrix({
    ("H", "P"): -2, ("E", "P"): -1, ("A", "P"): -1, ("G", "P"): -2, ("P", "W"): -4,
    ("A", "H"): -2, ("A", "E"): -1, ("A", "A"): 5, ("A", "G"): 0, ("A", "W"): -3,
    ("H", "W"): -3, ("E", "W"): -3, ("G", "W"): -3, ("W", "W"): 15,
    ("H", "H"): 10, ("E", "H"): 0, ("G", "H"): -2, ("E", "E"): 6, ("E", "G"): -3})

# Smith & Waterman, J. Mol. Biol. 147 (1981) 195-197, Figure 1: match +1, mismatch -1/3,
# W_k = 1 + k/3; maximum H = 3.33 for GCCAUUG / GCC-UCG.  Everything times 3 to stay integer;
# U written T (U is not a letter of the protein alphabet).
SW1981 = _matrix({}, default=-1, diagonal=3)

# Biopython Tutorial and Cookbook (1.8x), chapter "Pairwise sequence alignment": default
# PairwiseAligner = match 1, mismatch 0, every gap score 0.
IDENTITY = _matrix({}, default=0, diagonal=1)

# Biopython, Bio/pairwise2.py module docstring: globalms("ACCGT", "ACG", 2, -1, -.5, -.1) — match 2,
# mismatch -1, gap open -0.5, extend -0.1, end gaps penalised — prints "Score=5" (ACCGT / A-CG-).
# Everything times 10 to stay integer.
PAIRWISE2_MS = _matrix({}, default=-10, diagonal=20)

CASES = [
    ("durbin-global", "HEAGAWGHEE", "PAWHEAE", "global", DURBIN_BLOSUM50, -8, -8, 1,
     "Durbin et al. 1998, Fig. 2.5: bottom-right cell of the global DP matrix, linear gap d = 8"),
    ("durbin-local", "HEAGAWGHEE", "PAWHEAE", "local", DURBIN_BLOSUM50, -8, -8, 28,
     "Durbin et al. 1998, Fig. 2.6: best local alignment AWGHE / AW-HE"),
    ("smith-waterman-1981", "AATGCCATTGACGG", "CAGCCTCGCTTAG", "local", SW1981, -4, -1, 10,
     "Smith & Waterman 1981, Fig. 1: H max = 3.33 (x3), affine W_k = 1 + k/3"),
    ("biopython-tutorial-global", "GAACT", "GAT", "global", IDENTITY, 0, 0, 3,
     "Biopython Tutorial, PairwiseAligner basic usage: aligner.score('GAACT', 'GAT') == 3.0"),
    ("biopython-tutorial-local", "AGAACTC", "GAACT", "local", IDENTITY, 0, 0, 5,
     "Biopython Tutorial, local mode: aligner.score('AGAACTC', 'GAACT') == 5.0"),
    ("biopython-tutorial-blosum62", "KEVLA", "EVL", "global", "BLOSUM62", 0, 0, 13,
     "Biopython Tutorial, substitution matrices: BLOSUM62, 'KEVLA' vs 'EVL' scores 13.0"),
    ("biopython-pairwise2-doc", "ACCGT", "ACG", "global", PAIRWISE2_MS, -5, -1, 50,
     "Bio.pairwise2 docstring: globalms('ACCGT', 'ACG', 2, -1, -.5, -.1) -> Score=5 (x10), end gaps penalised"),
]

# the 1981 paper also prints the alignment itself: GCCAUUG (query 3..10) over GCC-UCG (subject 2..8)
SW1981_ALIGNMENT = dict(score=10, identities=5, columns=7, q_start=3, q_end=10, s_start=2, s_end=8, gaps=1)
