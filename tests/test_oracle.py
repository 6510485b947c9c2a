"""The oracle pinned against the known answers and against
itself (two independent formulations), plus the algebraic properties the score
must have.  CPU only."""
import hashlib

import numpy as np
import pytest

import oracle
from oracle import gotoh_py

# SURVEY.md App. B — computed by the surveyor with throwaway code, independently of oracle/
APP_B = {
    "local_12_1": ("3324e1d34f95582aab9635cd515908cb7ebf6b1755375f93668eaa086003e15d",
                   [1344, 324, 300, 476, 387, 345, 342, 442, 863, 882, 817, 377, 308, 426, 403], 46757),
    "global_12_1": ("da26b1e3b6539a13dde20ab523cc200ea6574231ddd6bbc44f38013cfa7c8edc",
                    [1344, 274, 287, 464, 367, 332, 341, 441, 863, 882, 805, 351, 307, 425, 383], 44988),
    "local_11_1": ("2e1a8fee7e9a7dc2b8c2effa96ffa4ca00803d1da65d637ef27eab83f1db72d7",
                   [1344, 328, 308, 479, 391, 351, 349, 446, 865, 883, 818, 381, 319, 431, 409], 47213),
    "global_11_1": ("00d6e4b084cbc29c2b773c38bc8b35595fc9454fb501217298d450b4d120fa7d",
                    [1344, 279, 296, 468, 373, 339, 348, 445, 865, 883, 807, 357, 318, 430, 391], 45554),
}
DIAG = [1344, 1100, 1316, 1343, 1295, 1309, 1300, 1307, 1361, 1353, 1347, 1279, 1436, 1321, 1323]


@pytest.mark.parametrize("key", sorted(APP_B))
def test_oracle_reproduces_app_b(ks_records, key):
    sha, row0, upper_sum = APP_B[key]
    mode_name, o, e = key.split("_")
    mode = oracle.MODE_LOCAL if mode_name == "local" else oracle.MODE_GLOBAL
    seqs = [r["sequence"] for r in ks_records]
    m = oracle.all_pairs(seqs, mode=mode, open_gap=-int(o), extend_gap=-int(e))
    assert m[0].tolist() == row0
    assert np.diag(m).tolist() == DIAG
    assert int(np.triu(m, 1).sum()) == upper_sum
    assert hashlib.sha256(m.astype("<i4").tobytes()).hexdigest() == sha


def test_literature_known_answers():
    """Published worked examples (tests/literature.py cites page and figure) through the matrix=
    path of all three oracle formulations: C H/E/F, C 3-state, pure-Python 3-state with floats."""
    import literature
    for name, a, b, mode, matrix, o, e, want, source in literature.CASES:
        sub = oracle.BLOSUM62 if isinstance(matrix, str) else matrix
        cm = oracle.MODE_LOCAL if mode == "local" else oracle.MODE_GLOBAL
        assert oracle.score(a, b, mode=cm, sub=sub, open_gap=o, extend_gap=e) == want, (name, source)
        assert oracle.score(b, a, mode=cm, sub=sub, open_gap=o, extend_gap=e) == want, (name, source)
        assert oracle.score_3state(a, b, mode=cm, sub=sub, open_gap=o, extend_gap=e) == want, (name, source)
        assert gotoh_py.score(a, b, mode, o, e, sub=sub) == float(want), (name, source)
    a, b = literature.CASES[2][1], literature.CASES[2][2]
    st = oracle.stats_list([a, b], [0], [1], mode=oracle.MODE_LOCAL, sub=literature.SW1981, open_gap=-4,
                           extend_gap=-1)[0]
    assert {k: int(st[k]) for k in st.dtype.names} == literature.SW1981_ALIGNMENT


def test_committed_golden_matches_app_b(ks_golden):
    for key, (sha, row0, _) in APP_B.items():
        assert ks_golden[key]["sha256_int32_le"] == sha
        assert ks_golden[key]["matrix"][0] == row0


def test_fixture_shape(ks_records):
    lens = [len(r["sequence"]) for r in ks_records]
    assert len(lens) == 15 and min(lens) == 212 and max(lens) == 274 and sum(lens) == 3725
    assert set("".join(r["sequence"] for r in ks_records)) <= set("ARNDCQEGHILKMFPSTWYV")


def test_blosum62_table():
    m = oracle.BLOSUM62.astype(int)
    assert (m == m.T).all()
    assert m[:20, :20].sum() == -426 and m.min() == -4 and m.max() == 11
    assert np.diag(m).tolist() == [4, 5, 6, 6, 9, 5, 5, 6, 8, 4, 4, 5, 5, 6, 7, 4, 5, 11, 7, 4, 4, 4, -1, 1]
    from bgclib_b200.alphabet import ALPHABET, BLOSUM62
    assert ALPHABET == oracle.ALPHABET
    assert (BLOSUM62 == oracle.BLOSUM62).all()     # product and oracle carry separate transcriptions


def _rand_seq(rng, n, letters=oracle.ALPHABET):
    return "".join(rng.choice(list(letters), size=n))


@pytest.mark.parametrize("mode", [oracle.MODE_LOCAL, oracle.MODE_GLOBAL])
def test_hef_equals_three_state(mode):
    rng = np.random.default_rng(7)
    for _ in range(150):
        a = _rand_seq(rng, int(rng.integers(1, 60)))
        b = _rand_seq(rng, int(rng.integers(1, 60)))
        if rng.random() < 0.4:
            b = (a[: len(a) // 2] + b)[: max(1, len(b))]
        for o, e in ((-12, -1), (-11, -1), (-5, -2), (-3, -3), (-1, 0)):
            assert oracle.score(a, b, mode=mode, open_gap=o, extend_gap=e) == \
                oracle.score_3state(a, b, mode=mode, open_gap=o, extend_gap=e)


@pytest.mark.parametrize("mode", ["local", "global"])
def test_c_equals_pure_python(ks_records, mode):
    rng = np.random.default_rng(11)
    cmode = oracle.MODE_LOCAL if mode == "local" else oracle.MODE_GLOBAL
    for _ in range(25):
        a = _rand_seq(rng, int(rng.integers(1, 40)))
        b = _rand_seq(rng, int(rng.integers(1, 40)))
        assert float(oracle.score(a, b, mode=cmode)) == gotoh_py.score(a, b, mode)
    # two real pairs (the rest of config 1 is covered by the SHA-256 above)
    s = [r["sequence"] for r in ks_records]
    for i, j in ((0, 8), (1, 1)):
        assert float(oracle.score(s[i], s[j], mode=cmode)) == gotoh_py.score(s[i], s[j], mode)


def test_properties():
    rng = np.random.default_rng(3)
    sub = oracle.BLOSUM62.astype(int)
    for _ in range(60):
        a = _rand_seq(rng, int(rng.integers(1, 80)))
        b = _rand_seq(rng, int(rng.integers(1, 80)))
        sl, sg = oracle.score(a, b, mode=0), oracle.score(a, b, mode=1)
        assert sl == oracle.score(b, a, mode=0) and sg == oracle.score(b, a, mode=1)   # symmetry
        assert sl >= 0 and sl >= sg
        ca = oracle.encode(a)
        assert oracle.score(a, a, mode=1) == int(sub[ca, ca].sum())                    # global self = sum diag
    a20 = _rand_seq(rng, 50, "ARNDCQEGHILKMFPSTWYV")
    c = oracle.encode(a20)
    assert oracle.score(a20, a20, mode=0) == int(sub[c, c].sum())


def test_gap_convention():
    # a gap of length k scores open + (k-1)*extend: "AAAA" vs "AA" global = 2*4 + (-12) + (-1)
    assert oracle.score("AAAA", "AA", mode=1) == 8 - 12 - 1
    assert oracle.score("AAAAA", "AA", mode=1) == 8 - 12 - 2
    assert oracle.score("W", "W", mode=0) == 11
    assert oracle.score("W", "P", mode=0) == 0 and oracle.score("W", "P", mode=1) == -4


def test_errors():
    with pytest.raises(ValueError):
        oracle.encode("")
    for bad in ("ACDU", "ACDO", "ACDJ", "acd", "AC-D"):
        with pytest.raises(ValueError):
            oracle.encode(bad)
    with pytest.raises(ValueError):
        gotoh_py.score("", "A")
    assert oracle.encode("BZX*").tolist() == [20, 21, 22, 23]


def test_bgc_similarity_restatement():
    # 3 BGCs, hand-checkable: scores symmetric with diag dominant
    s = np.array([[10, 4, 2, 0], [4, 8, 3, 1], [2, 3, 12, 5], [0, 1, 5, 9]], dtype=np.int32)
    bgc = [0, 0, 1, 2]
    typ = [0, 1, 0, 1]
    sim, best, selfs, rowbest = oracle.bgc_similarity_np(s, bgc, typ, 3, same_type_only=True)
    assert selfs.tolist() == [18, 12, 9]
    assert best[0].tolist() == [18, 2, 1] and best[1].tolist() == [2, 12, 0] and best[2].tolist() == [1, 0, 9]
    assert sim[0, 1] == (2 + 2) / (18 + 12) and sim[0, 0] == 1.0 and sim[1, 2] == 0.0
    sim2, best2, _, _ = oracle.bgc_similarity_np(s, bgc, typ, 3, same_type_only=False)
    assert best2[0].tolist() == [18, 5, 1]


# ---- alignment statistics without traceback (oracle/gotoh_oracle.c bgco_stats_pair) ---------
def _s(rec):
    return {k: int(rec[k]) for k in rec.dtype.names}


def test_stats_three_formulations_agree_on_tiny_inputs():
    """C packed-key H/E/F DP == Python 3-state tuple DP == enumeration of every path."""
    from oracle import stats_py
    rng = np.random.default_rng(5)
    for _ in range(120):
        a = "".join("ACDE"[i] for i in rng.integers(0, 4, int(rng.integers(1, 5))))
        b = "".join("ACDE"[i] for i in rng.integers(0, 4, int(rng.integers(1, 5))))
        for mode, cm in (("local", oracle.MODE_LOCAL), ("global", oracle.MODE_GLOBAL)):
            for o, e in ((-12, -1), (-2, -1), (-1, -1), (-3, 0), (0, 0)):
                c = _s(oracle.stats_list([a, b], [0], [1], mode=cm, open_gap=o, extend_gap=e)[0])
                assert c == stats_py.stats_3state(a, b, mode, o, e) == stats_py.stats_bruteforce(a, b, mode, o, e)


def test_stats_c_vs_python_on_domain_sized_inputs(ks_records):
    from oracle import stats_py
    seqs = [r["sequence"][:120] for r in ks_records[:4]]
    for i, j in ((0, 1), (2, 0), (3, 3)):
        for mode, cm in (("local", oracle.MODE_LOCAL), ("global", oracle.MODE_GLOBAL)):
            assert _s(oracle.stats_list(seqs, [i], [j], mode=cm)[0]) == stats_py.stats_3state(seqs[i], seqs[j], mode)


def test_stats_golden_and_properties(ks_records, ks_golden):
    import json
    from conftest import GOLDEN
    gold = json.loads((GOLDEN / "ks_stats.json").read_text())
    seqs = [r["sequence"] for r in ks_records]
    n = len(seqs)
    pi, pj = np.divmod(np.arange(n * n), n)
    lens = np.array([len(s) for s in seqs])
    for mode, cm in (("local", oracle.MODE_LOCAL), ("global", oracle.MODE_GLOBAL)):
        st = oracle.stats_list(seqs, pi, pj, mode=cm)
        rows = np.stack([st[k] for k in gold["fields"]], axis=1).astype("<i4")
        assert hashlib.sha256(rows.tobytes()).hexdigest() == gold[mode]["sha256_int32_le"]
        assert (rows == np.array(gold[mode]["records"], dtype=np.int32)).all()
        # the score field IS the alignment score of the scoring path
        assert (st["score"].reshape(n, n) == np.array(ks_golden[f"{mode}_12_1"]["matrix"])).all()
        assert (st["identities"] <= st["columns"]).all() and (st["gaps"] >= 0).all()
        assert (st["q_end"] - st["q_start"] + st["s_end"] - st["s_start"] == 2 * st["columns"] - st["gaps"]).all()
        d = st.reshape(n, n)[np.arange(n), np.arange(n)]        # self alignment = the whole sequence
        assert (d["identities"] == lens).all() and (d["columns"] == lens).all() and (d["gaps"] == 0).all()
        if mode == "global":
            assert (st["q_start"] == 0).all() and (st["q_end"] == lens[pi]).all() and (st["s_end"] == lens[pj]).all()
    # swapping query and subject keeps the additive value (score, identities, columns)
    loc = oracle.stats_list(seqs, pi, pj).reshape(n, n)
    for k in ("score", "identities", "columns", "gaps"):
        assert (loc[k] == loc[k].T).all()


def test_stats_empty_local_alignment_and_limits():
    st = oracle.stats_list(["WWWW", "PPPP"], [0], [1])[0]
    assert _s(st) == dict(score=0, identities=0, columns=0, q_start=-1, q_end=-1, s_start=-1, s_end=-1, gaps=0)
    with pytest.raises(ValueError):
        oracle.stats_list(["A" * 32768, "A"], [0], [1])


# ---- property tests (hypothesis) over random sequences, incl. the ambiguity letters -------------
from hypothesis import given, settings, strategies as st  # noqa: E402

_seq = st.text(alphabet=oracle.ALPHABET, min_size=1, max_size=40)
_gaps = st.sampled_from([(-12, -1), (-11, -1), (-5, -2), (-3, -3), (-1, 0), (0, 0)])


@settings(max_examples=150, deadline=None)
@given(a=_seq, b=_seq, gaps=_gaps)
def test_properties_of_the_score(a, b, gaps):
    o, e = gaps
    loc = oracle.score(a, b, mode=oracle.MODE_LOCAL, open_gap=o, extend_gap=e)
    glo = oracle.score(a, b, mode=oracle.MODE_GLOBAL, open_gap=o, extend_gap=e)
    assert loc == oracle.score(b, a, mode=oracle.MODE_LOCAL, open_gap=o, extend_gap=e)      # symmetry
    assert glo == oracle.score(b, a, mode=oracle.MODE_GLOBAL, open_gap=o, extend_gap=e)
    assert loc >= 0 and loc >= glo                                                          # local >= global
    assert loc == oracle.score_3state(a, b, mode=oracle.MODE_LOCAL, open_gap=o, extend_gap=e)
    assert glo == oracle.score_3state(a, b, mode=oracle.MODE_GLOBAL, open_gap=o, extend_gap=e)
    assert float(loc) == gotoh_py.score(a, b, "local", o, e) and float(glo) == gotoh_py.score(a, b, "global", o, e)
    # appending a residue cannot lower the local score; reversing both sequences keeps both scores
    assert oracle.score(a + "W", b, mode=oracle.MODE_LOCAL, open_gap=o, extend_gap=e) >= loc
    assert oracle.score(a[::-1], b[::-1], mode=oracle.MODE_GLOBAL, open_gap=o, extend_gap=e) == glo
    assert oracle.score(a[::-1], b[::-1], mode=oracle.MODE_LOCAL, open_gap=o, extend_gap=e) == loc


@settings(max_examples=60, deadline=None)
@given(a=_seq, b=_seq)
def test_properties_of_the_statistics(a, b):
    for mode, cm in (("local", oracle.MODE_LOCAL), ("global", oracle.MODE_GLOBAL)):
        s = _s(oracle.stats_list([a, b], [0], [1], mode=cm)[0])
        assert s["score"] == oracle.score(a, b, mode=cm)
        assert 0 <= s["identities"] <= s["columns"] - s["gaps"] <= min(len(a), len(b)) or s["columns"] == 0
        if s["columns"]:
            qa, sb = a[s["q_start"]:s["q_end"]], b[s["s_start"]:s["s_end"]]
            # the reported window alone reproduces the score (global alignment of the two windows)
            assert oracle.score(qa, sb, mode=oracle.MODE_GLOBAL) >= s["score"] or mode == "global"
            assert oracle.score(qa, sb, mode=cm) == s["score"]


# ---- opt-in: the real Biopython aligner, if it is ever importable where the tests run -----------
def test_against_biopython_if_available(ks_records):
    Align = pytest.importorskip("Bio.Align", reason="Biopython is not installed in this image (SURVEY.md F5)")
    from Bio.Align import substitution_matrices
    seqs = [r["sequence"] for r in ks_records]
    for mode, cm in (("local", oracle.MODE_LOCAL), ("global", oracle.MODE_GLOBAL)):
        al = Align.PairwiseAligner()
        al.mode = mode
        al.substitution_matrix = substitution_matrices.load("BLOSUM62")
        al.open_gap_score, al.extend_gap_score = -12, -1
        want = oracle.all_pairs(seqs, mode=cm)
        for i in range(len(seqs)):
            for j in range(i, len(seqs)):
                assert al.score(seqs[i], seqs[j]) == float(want[i, j])
