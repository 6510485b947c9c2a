"""Parity of the CUDA path (through the C ABI) against the oracle — bit-exact
integer scores.  Run on the B200 box:  python -m pytest tests -m gpu -x -q"""
import hashlib

import numpy as np
import pytest

import oracle
from bgclib_b200 import bgc_similarity, domain_pair_scores, extract, get_engine, synth
from fakes import build_collection

pytestmark = pytest.mark.gpu

AA = "ARNDCQEGHILKMFPSTWYV"


def _rand(rng, n, letters=AA):
    return "".join(rng.choice(list(letters), size=int(n)))


def _related(rng, n_fam, fam_size, lo, hi, letters=AA):
    out = []
    for _ in range(n_fam):
        anc = list(_rand(rng, rng.integers(lo, hi + 1), letters))
        for _ in range(fam_size):
            m = anc.copy()
            for _ in range(int(len(m) * rng.uniform(0.05, 0.5))):
                m[int(rng.integers(0, len(m)))] = letters[int(rng.integers(0, len(letters)))]
            if rng.random() < 0.7:
                p = int(rng.integers(0, len(m)))
                k = int(rng.integers(1, 9))
                m = m[:p] + list(_rand(rng, k, letters)) + m[p:] if rng.random() < 0.5 else m[:p] + m[p + k:]
            out.append("".join(m) if len(m) > 0 else "A")
    return out


# ---- config 1: the reference's own fixture, four parameter settings ------------------------
@pytest.mark.parametrize("key", ["local_12_1", "global_12_1", "local_11_1", "global_11_1"])
def test_config1_ks_domains_golden(ks_records, ks_golden, key):
    g = ks_golden[key]
    seqs = [r["sequence"] for r in ks_records]
    res = domain_pair_scores(seqs, mode=g["mode"], open_gap_score=g["open"], extend_gap_score=g["extend"])
    assert res.scores.dtype == np.int32 and res.scores.shape == (15, 15)
    assert (res.scores == np.array(g["matrix"], dtype=np.int32)).all()
    assert hashlib.sha256(res.scores.astype("<i4").tobytes()).hexdigest() == g["sha256_int32_le"]
    assert res.plan["total_pairs"] == 120


# ---- random / related sequences over every group width and many row counts -----------------
@pytest.mark.parametrize("mode", ["local", "global"])
@pytest.mark.parametrize("lo,hi", [(1, 16), (20, 70), (60, 140), (200, 300), (380, 460), (500, 640), (900, 1024)])
def test_random_lengths_bit_exact(mode, lo, hi):
    rng = np.random.default_rng(lo * 7 + hi + (mode == "global"))
    seqs = _related(rng, 4, 6, lo, hi) + [_rand(rng, rng.integers(lo, hi + 1)) for _ in range(9)]
    seqs = [s[:hi] for s in seqs]
    res = domain_pair_scores(seqs, mode=mode)
    want = oracle.all_pairs(seqs, mode=oracle.MODE_LOCAL if mode == "local" else oracle.MODE_GLOBAL)
    assert (res.scores == want).all()
    assert res.plan["n_fallback32"] == 0


@pytest.mark.parametrize("force_g", ["1", "2", "4", "8", "16", "32"])
def test_every_group_width(monkeypatch, force_g):
    monkeypatch.setenv("BGCA_FORCE_G", force_g)
    rng = np.random.default_rng(int(force_g))
    seqs = _related(rng, 5, 5, 30, 250) + [_rand(rng, n) for n in (1, 2, 3, 31, 32, 33, 64, 255, 256)]
    for mode, cm in (("local", 0), ("global", 1)):
        res = domain_pair_scores(seqs, mode=mode)
        assert (res.scores == oracle.all_pairs(seqs, mode=cm)).all()


def test_mixed_lengths_and_ambiguity_letters():
    rng = np.random.default_rng(99)
    letters = "ARNDCQEGHILKMFPSTWYVBZX*"
    seqs = _related(rng, 6, 5, 40, 600, letters) + ["*", "X", "BZ", "W" * 100]
    for mode, cm in (("local", 0), ("global", 1)):
        for o, e in ((-12, -1), (-11, -1), (-5, -2), (-3, -3)):
            res = domain_pair_scores(seqs, mode=mode, open_gap_score=o, extend_gap_score=e)
            want = oracle.all_pairs(seqs, mode=cm, open_gap=o, extend_gap=e)
            assert (res.scores == want).all(), (mode, o, e)


def test_long_sequences_take_the_32bit_kernel():
    rng = np.random.default_rng(5)
    base = _rand(rng, 3300)
    seqs = [base[:1500], base[200:3300], _rand(rng, 1100), base[1000:2100], _rand(rng, 300), base[:257],
            "W" * 3200, "WY" * 1550]
    for mode, cm in (("local", 0), ("global", 1)):
        res = domain_pair_scores(seqs, mode=mode)
        assert res.plan["n_fallback32"] > 0
        assert (res.scores == oracle.all_pairs(seqs, mode=cm)).all()
    # the poly-W self score (3200 * 11) exceeds int16: only the 32-bit path can hold it
    loc = domain_pair_scores(seqs, mode="local").scores
    assert loc[6, 6] == 35200


@pytest.mark.parametrize("mode", ["local", "global"])
def test_protein_length_queries_stay_on_the_packed_kernel(mode):
    """Queries of more than 1024 residues run as a multi-pass sweep of the whole-warp packed
    kernel (bottom row of a pass parked for the next one) — no 32-bit fallback while the scores
    fit int16.  Pass boundaries 1024/1025, 2048/2049 and the two halves of a query pair ending in
    different passes are all in here."""
    rng = np.random.default_rng(17 + (mode == "global"))
    seqs = _related(rng, 2, 4, 1025, 2048) + [_rand(rng, k) for k in (1024, 1025, 1280, 1281, 1536, 1793, 2047, 2048,
                                                                     2049, 2500, 90)]
    res = domain_pair_scores(seqs, mode=mode)
    assert res.plan["n_fallback32"] == 0
    cm = oracle.MODE_LOCAL if mode == "local" else oracle.MODE_GLOBAL
    assert (res.scores == oracle.all_pairs(seqs, mode=cm)).all()


@pytest.mark.parametrize("mode", ["local", "global"])
def test_multi_pass_many_subjects_and_long_queries(mode):
    """The multi-pass sweep with full tiles (several subjects per group, so the parked line is
    indexed across subject boundaries), three- and six-pass queries against short database
    domains (cross plan), other gap scores (run-time gap kernels), and the fused rowbest epilogue."""
    rng = np.random.default_rng(23 + (mode == "global"))
    base = _rand(rng, 6000)
    long_q = [base[:1100], base[300:1700], base[:2100], base[100:3000], base[:6000], _rand(rng, 1300)]
    short = _related(rng, 3, 8, 60, 500) + [base[k:k + 300] for k in range(0, 5400, 450)]
    cm = oracle.MODE_LOCAL if mode == "local" else oracle.MODE_GLOBAL
    res = domain_pair_scores(long_q, database=short, mode=mode)
    assert res.plan["n_fallback32"] <= 1       # only the self tile of the 6000-residue query (score > int16)
    full = oracle.all_pairs(long_q + short, mode=cm)
    assert (res.scores == full[: len(long_q), len(long_q):]).all()
    # square plan: 40 sequences of 1025..1500 residues -> tiles with several subjects per group
    many = _related(rng, 4, 10, 1025, 1500)
    for o, e in ((-12, -1), (-7, -2)):
        res = domain_pair_scores(many, mode=mode, open_gap_score=o, extend_gap_score=e)
        assert res.plan["n_fallback32"] == 0
        assert (res.scores == oracle.all_pairs(many, mode=cm, open_gap=o, extend_gap=e)).all(), (o, e)


def test_literature_known_answers():
    """The published worked examples of tests/literature.py through the C ABI (custom matrices,
    linear and affine gaps, both modes), scores and — for Smith & Waterman's 1981 figure — the
    alignment coordinates."""
    import literature
    from bgclib_b200 import domain_pair_stats
    for name, a, b, mode, matrix, o, e, want, source in literature.CASES:
        res = domain_pair_scores([a, b], mode=mode, matrix=matrix, open_gap_score=o, extend_gap_score=e)
        assert res.scores[0, 1] == want and res.scores[1, 0] == want, (name, source)
    a, b = literature.CASES[2][1], literature.CASES[2][2]
    st = domain_pair_stats([a, b], np.array([[0, 1]]), mode="local", matrix=literature.SW1981,
                           open_gap_score=-4, extend_gap_score=-1).stats[0]
    assert {k: int(st[k]) for k in st.dtype.names} == literature.SW1981_ALIGNMENT


def test_long_queries_against_a_persistent_database():
    """Queries longer than one pass against a PackedDatabase — on the multi-pass packed kernel,
    and, with gap scores that leave int16 in global mode, on the 32-bit kernel, whose pair list
    must hold every query x database pair although the database sorts in FRONT of its queries."""
    from bgclib_b200 import PackedDatabase
    rng = np.random.default_rng(41)
    base = _rand(rng, 2700)
    db_seqs = [base[:100], base[50:170], base[:150], _rand(rng, 200), base[400:620], base[:300]]
    q_seqs = [base[:2500], base[60:2660], base[100:400]]
    db = PackedDatabase(db_seqs)
    for mode, cm in (("local", oracle.MODE_LOCAL), ("global", oracle.MODE_GLOBAL)):
        for (o, e), fallback in (((-12, -1), False), ((-40, -40), mode == "global")):
            dps = domain_pair_scores(q_seqs, database=db, mode=mode, open_gap_score=o, extend_gap_score=e)
            assert (dps.plan["n_fallback32"] >= 1) == fallback, (mode, o, e)
            full = oracle.all_pairs(q_seqs + db_seqs, mode=cm, open_gap=o, extend_gap=e)
            assert (dps.scores == full[:3, 3:]).all(), (mode, o, e)
    # local scores that leave int16 (W-rich, > 2727 residues on both sides): 32-bit kernel again
    wq, wdb = ["W" * 2800 + base[:50], base[:90]], ["WY" * 1400, "W" * 2750, base[20:200]]
    dps = domain_pair_scores(wq, database=PackedDatabase(wdb))
    assert dps.plan["n_fallback32"] >= 1
    assert (dps.scores == oracle.all_pairs(wq + wdb)[:2, 2:]).all()
    # the fused similarity on the same plan shape (rowbest of the long queries must not stay 0)
    from bgclib_b200.extract import DomainBatch, DomainRecord

    def as_batch(seqs, tag):
        recs = [DomainRecord(i, f"{tag}{i // 2}", f"p{i}", "d", "KS", 0, len(s), len(s)) for i, s in enumerate(seqs)]
        ids = sorted({r.bgc_id for r in recs})
        return DomainBatch(list(seqs), recs, ids, np.asarray([ids.index(r.bgc_id) for r in recs], np.int32),
                           ["KS"], np.zeros(len(seqs), np.int32))
    qb, dbb = as_batch(q_seqs, "q"), as_batch(db_seqs, "d")
    got = bgc_similarity(qb, database=PackedDatabase(dbb))
    want = bgc_similarity(qb, database=dbb)
    assert (got.best == want.best).all() and (got.best_db == want.best_db).all() and (got.sim == want.sim).all()
    assert (got.best > 0).all()
    # oracle statement of the query -> database direction
    sc = oracle.all_pairs(q_seqs + db_seqs)
    for a in range(len(qb.bgc_ids)):
        for b in range(len(dbb.bgc_ids)):
            rows = [i for i in range(3) if qb.bgc_of_seq[i] == a]
            cols = [3 + j for j in range(6) if dbb.bgc_of_seq[j] == b]
            assert got.best[a, b] == sum(max(sc[i, j] for j in cols) for i in rows)


def test_int16_guard_sees_the_longest_subject_of_a_mixed_type_tile():
    """Types packed, every pair planned (domain_pair_scores against a PackedDatabase): a tile
    crosses type runs, so its last subject is not its longest.  A 30000-residue subject in global
    mode leaves int16 and has to take the 32-bit kernel."""
    from bgclib_b200 import PackedDatabase
    from bgclib_b200.extract import DomainBatch, DomainRecord
    rng = np.random.default_rng(43)
    seqs = [_rand(rng, 100), _rand(rng, 110), _rand(rng, 30000), _rand(rng, 50)]
    types = ["A", "A", "A", "T"]
    recs = [DomainRecord(i, f"b{i}", f"p{i}", "d", t, 0, len(s), len(s)) for i, (s, t) in enumerate(zip(seqs, types))]
    dbb = DomainBatch(seqs, recs, [f"b{i}" for i in range(4)], np.arange(4, dtype=np.int32), ["A", "T"],
                      np.asarray([0, 0, 0, 1], np.int32))
    q = [_rand(rng, 105), _rand(rng, 95)]
    for mode, cm in (("local", oracle.MODE_LOCAL), ("global", oracle.MODE_GLOBAL)):
        dps = domain_pair_scores(q, database=PackedDatabase(dbb), mode=mode)
        want = oracle.pair_list(q + seqs, np.repeat([0, 1], 4), np.tile([2, 3, 4, 5], 2), mode=cm).reshape(2, 4)
        assert (dps.scores == want).all(), mode


def test_rejected_query_sets_leave_the_database_usable():
    """Every way a query set can be refused (empty sequence, too long, illegal letter) leaves the
    PackedDatabase ready for the next call."""
    from bgclib_b200 import PackedDatabase
    rng = np.random.default_rng(44)
    db_seqs = [_rand(rng, n) for n in (80, 120, 200)]
    db = PackedDatabase(db_seqs)
    q = [_rand(rng, 90), _rand(rng, 150)]
    want = oracle.all_pairs(q + db_seqs)[:2, 2:]
    for bad in ([q[0], ""], [q[0], "A" * 70000], [q[0], "ACDUFG"]):
        with pytest.raises(ValueError):      # straight at the C ABI (extract() would drop the empty one)
            db.engine.pack_append(bad)
        assert (domain_pair_scores(q, database=db).scores == want).all()


def test_pair_list_32bit_kernel():
    rng = np.random.default_rng(6)
    seqs = _related(rng, 3, 4, 100, 700)
    eng = get_engine()
    for mode, cm in (("local", 0), ("global", 1)):
        eng.set_scoring(mode=mode)
        eng.pack(seqs)
        pi = rng.integers(0, len(seqs), 64)
        pj = rng.integers(0, len(seqs), 64)
        got = eng.score_pairs32(pi, pj).cpu().numpy()
        assert (got == oracle.pair_list(seqs, pi, pj, mode=cm)).all()


def test_packer_matches_oracle_encoding():
    rng = np.random.default_rng(8)
    seqs = [_rand(rng, n, "ARNDCQEGHILKMFPSTWYVBZX*") for n in (1, 15, 16, 17, 33, 250, 64, 5, 480)]
    eng = get_engine()
    eng.pack(seqs)
    codes, offs, order = eng.get_packed()
    lens = np.array([len(s) for s in seqs])
    assert (lens[order] == np.sort(lens)).all() and sorted(order.tolist()) == list(range(len(seqs)))
    for r, o in enumerate(order):
        seg = codes[offs[r]:offs[r + 1]]
        assert offs[r] % 16 == 0 and len(seg) == (lens[o] + 16) // 16 * 16
        assert (seg[: lens[o]] == oracle.encode(seqs[o])).all()
        assert (seg[lens[o]:] == 24).all()


def test_packer_alignment_long_sequences_and_device_input():
    """K1's aligned-word loads: every source misalignment 0..15, lengths around the 16- and
    512-byte boundaries, ASCII handed over as an unaligned device view."""
    import torch
    rng = np.random.default_rng(9)
    lens = [1, 2, 15, 16, 17, 31, 32, 33, 255, 256, 257, 495, 496, 497, 511, 512, 513, 527, 528, 529, 1023, 1024,
            1025, 2000, 3, 7, 11, 13, 5, 9]
    seqs = [_rand(rng, n, "ARNDCQEGHILKMFPSTWYVBZX*") for n in lens]
    eng = get_engine()
    from bgclib_b200.engine import concat_ascii
    ascii_, offsets = concat_ascii(seqs)
    for shift in (0, 1, 5, 8, 15):
        buf = torch.zeros(len(ascii_) + 64, dtype=torch.uint8, device="cuda")
        view = buf[shift: shift + len(ascii_)]
        view.copy_(torch.from_numpy(ascii_.copy()))
        eng.pack_ascii(view, offsets)
        codes, offs, order = eng.get_packed()
        for r, o in enumerate(order):
            seg = codes[offs[r]:offs[r + 1]]
            assert len(seg) == (lens[o] + 16) // 16 * 16
            assert (seg[: lens[o]] == oracle.encode(seqs[o])).all() and (seg[lens[o]:] == 24).all()
    # the smallest (sequence, position) of an illegal letter is what the error names
    bad = list(seqs)
    bad[21] = bad[21][:700] + "U" + bad[21][701:900] + "J" + bad[21][901:]
    bad[23] = bad[23][:1999] + "O"
    with pytest.raises(ValueError, match="sequence 21 .* position 700"):
        eng.pack(bad)


def test_errors_mirror_biopython():
    with pytest.raises(ValueError, match="alphabet"):
        domain_pair_scores(["ACDEFG", "ACDUFG"])
    with pytest.raises(ValueError, match="alphabet"):
        get_engine().pack(["ACDEFG", "acdefg"])       # the engine itself does not upper-case
    with pytest.raises(ValueError, match="zero length"):
        get_engine().pack(["ACD", ""])
    with pytest.raises(ValueError):
        domain_pair_scores([])
    with pytest.raises(ValueError):
        domain_pair_scores(["ACD"], open_gap_score=-1, extend_gap_score=-5)
    with pytest.raises(ValueError):
        domain_pair_scores(["ACD"], open_gap_score=-11.5)
    # the engine is still usable after a failed call
    assert domain_pair_scores(["WW", "WW"]).scores.tolist() == [[22, 22], [22, 22]]


@pytest.mark.parametrize("mode", ["local", "global"])
@pytest.mark.parametrize("o,e", [(-10, -2), (-5, -5), (-1, 0), (0, 0), (-40, -3), (-12, -2)])
def test_run_time_gap_scores_and_custom_matrices(mode, o, e):
    """Gap scores outside the compiled presets (run-time operands) and substitution matrices
    other than BLOSUM62: a random symmetric small-integer matrix and one large enough that the
    planner must route every tile to the 32-bit kernel."""
    rng = np.random.default_rng(abs(o) * 31 + abs(e) + (mode == "global"))
    seqs = _related(rng, 3, 4, 30, 330, AA + "BZX*") + [_rand(rng, k, AA) for k in (1, 2, 17, 333)]
    cm = oracle.MODE_LOCAL if mode == "local" else oracle.MODE_GLOBAL
    res = domain_pair_scores(seqs, mode=mode, open_gap_score=o, extend_gap_score=e)
    assert (res.scores == oracle.all_pairs(seqs, mode=cm, open_gap=o, extend_gap=e)).all()
    m = rng.integers(-6, 9, (24, 24))
    m = np.triu(m) + np.triu(m, 1).T
    res = domain_pair_scores(seqs, mode=mode, matrix=m, open_gap_score=o, extend_gap_score=e)
    assert (res.scores == oracle.all_pairs(seqs, mode=cm, sub=m.astype(np.int8), open_gap=o, extend_gap=e)).all()
    assert res.plan["n_fallback32"] == 0
    big = m * 14                                   # |s| up to 112: 112 * 333 leaves int16
    res = domain_pair_scores(seqs, mode=mode, matrix=big, open_gap_score=o, extend_gap_score=e)
    assert (res.scores == oracle.all_pairs(seqs, mode=cm, sub=big.astype(np.int8), open_gap=o, extend_gap=e)).all()
    assert res.plan["n_fallback32"] > 0


@pytest.mark.parametrize("seed", range(6))
def test_random_collections_stress(seed):
    """Many small random collections: odd sequence counts (single query pairs, ragged subject
    chunks), lengths straddling the G*K row boundaries of the kernel variants and the 16-byte
    packing boundaries, families and unrelated sequences mixed, both modes, matrix and fused
    BGC reduction — everything against the oracle."""
    import torch
    rng = np.random.default_rng(1000 + seed)
    eng = get_engine()
    for _ in range(10):
        n = int(rng.integers(1, 70))
        style = int(rng.integers(0, 4))
        if style == 0:      # around a variant boundary G*K
            g, k = int(rng.choice([8, 16, 32])), int(rng.integers(2, 33))
            lens = np.clip(g * k + rng.integers(-3, 4, n), 1, 1024)
        elif style == 1:    # wide spread
            lens = rng.integers(1, 700, n)
        elif style == 2:    # short domains (T/ACP-like)
            lens = rng.integers(40, 100, n)
        else:               # multiples of 16 +- 1
            lens = np.clip(16 * rng.integers(1, 40, n) + rng.integers(-1, 2, n), 1, 1024)
        anc = _rand(rng, 1100)
        seqs = []
        for L in lens:
            if rng.random() < 0.6:
                m = list(anc[: int(L)])
                for _ in range(int(L * rng.uniform(0.0, 0.5))):
                    m[int(rng.integers(0, L))] = AA[int(rng.integers(0, 20))]
                seqs.append("".join(m))
            else:
                seqs.append(_rand(rng, L))
        mode = "local" if rng.random() < 0.6 else "global"
        cm = oracle.MODE_LOCAL if mode == "local" else oracle.MODE_GLOBAL
        want = oracle.all_pairs(seqs, mode=cm)
        got = domain_pair_scores(seqs, mode=mode).scores
        assert (got == want).all(), (seed, n, style, mode)
        if mode == "local":
            B = int(rng.integers(1, 9))
            T = int(rng.integers(1, 4))
            bgc_of = rng.integers(0, B, n).astype(np.int32)
            type_of = rng.integers(0, T, n).astype(np.int32)
            eng.set_scoring()
            eng.pack(seqs, type_of_seq=type_of, bgc_of_seq=bgc_of, n_bgc=B)
            eng.plan(same_type_only=True)
            rowbest = torch.zeros((n, B), dtype=torch.int32, device="cuda")
            selfsc = torch.zeros(n, dtype=torch.int32, device="cuda")
            eng.score(rowbest=rowbest, selfsc=selfsc)
            sim, best, selfsum = eng.reduce_bgc(rowbest, selfsc)
            o_sim, o_best, o_self, o_rowbest = oracle.bgc_similarity_np(want, bgc_of, type_of, B)
            assert (rowbest.cpu().numpy() == o_rowbest).all() and (best.cpu().numpy() == o_best).all()
            assert (selfsum.cpu().numpy() == o_self).all() and (sim.cpu().numpy() == o_sim).all()


# ---- config 2: 37 Af293 BGCs -> BGC-pair similarity ----------------------------------------
def _af293_collection(af293, ks_records, seed=20260102):
    """Config 2b: the 282-domain architecture of the 57 CBPs with the real
    KS sequences where they exist and seeded synthetic sequences for every other domain."""
    rng = np.random.default_rng(seed)
    ks_by_protein = {}
    for r in ks_records:
        ks_by_protein.setdefault(r["protein_identifier"], []).append(r["sequence"])
    length = {"KS": 253, "KS_C": 119, "KS_Ce": 112, "AT": 319, "C": 457, "A": 424, "A_C": 76, "DH": 298,
              "KR": 180, "SAT": 240, "ER": 313, "PT": 329, "DMAT": 363, "T/ACP": 70}
    anc = {}
    spec = {b: [] for b in af293["bgcs"]}
    for cbp in af293["cbps"]:
        seq, doms, pos = "M", [], 1
        real = list(ks_by_protein.get(cbp["protein_identifier"], []))
        for t in cbp["domains"]:
            if t == "KS" and real:
                s = real.pop(0)
            else:
                L = length.get(t, 150)
                if t not in anc:
                    anc[t] = _rand(rng, L)
                m = list(anc[t])
                for _ in range(int(L * rng.uniform(0.1, 0.6))):
                    m[int(rng.integers(0, L))] = AA[int(rng.integers(0, 20))]
                s = "".join(m)
            linker = _rand(rng, rng.integers(3, 12))
            doms.append((t, t, pos, pos + len(s)))
            seq += s + linker
            pos += len(s) + len(linker)
        spec[cbp["bgc"]].append((cbp["protein_identifier"], seq, doms))
    return build_collection(spec)


@pytest.mark.parametrize("same_type_only", [True, False])
def test_config2_af293_bgc_similarity(af293, ks_records, same_type_only):
    col = _af293_collection(af293, ks_records)
    batch = extract(col)
    assert len(batch) == 282 and len(batch.bgc_ids) == 37
    res = bgc_similarity(col, same_type_only=same_type_only, bgc_ids=af293["bgcs"])
    scores = oracle.all_pairs(batch.sequences)
    sim, best, selfs, _ = oracle.bgc_similarity_np(scores, batch.bgc_of_seq, batch.type_of_seq, 37,
                                                   same_type_only=same_type_only)
    assert res.bgc_ids == af293["bgcs"]
    assert (res.best == best).all() and (res.self_score == selfs).all()
    assert (res.sim == sim).all()                      # one correctly-rounded f64 divide: bit-exact
    assert res.sim.min() >= 0 and res.sim.max() <= 1.0 and (np.diag(res.sim)[selfs > 0] == 1.0).all()
    assert (res.sim[selfs == 0] == 0).all()            # BGCs without CBP domains: zero rows
    # the fused path equals the unfused one built from the materialised matrix
    dps = domain_pair_scores(col, same_type_only=same_type_only)
    if same_type_only:
        same = batch.type_of_seq[:, None] == batch.type_of_seq[None, :]
        assert (dps.computed == same).all()
        assert (dps.scores[same] == scores[same]).all() and (dps.scores[~same] == 0).all()
    else:
        assert (dps.scores == scores).all()


def test_config2a_ks_only(af293, ks_records):
    spec = {b: [] for b in af293["bgcs"]}
    for r in ks_records:
        spec[r["bgc"]].append((r["identifier"], r["sequence"], [("ketoacyl-synt", "KS", 0, len(r["sequence"]))]))
    col = build_collection(spec)
    res = bgc_similarity(col, bgc_ids=af293["bgcs"])
    nonempty = res.self_score > 0
    assert int(nonempty.sum()) == 13 and (np.diag(res.sim)[nonempty] == 1.0).all()
    assert (res.sim[~nonempty] == 0).all() and (res.sim[:, ~nonempty] == 0).all()
    seqs = [r["sequence"] for r in ks_records]
    bgc_of = np.array([af293["bgcs"].index(r["bgc"]) for r in ks_records])
    sim, best, selfs, _ = oracle.bgc_similarity_np(oracle.all_pairs(seqs), bgc_of, np.zeros(15, int), 37)
    assert (res.sim == sim).all() and (res.best == best).all()


# ---- multi-rank logic on one GPU: ranks' tile shares run one after another -------------------
def test_rank_shares_combine_to_single_rank_result():
    import torch
    col = synth.ks_like(96, lo=60, hi=200, mean=120, sd=30)
    seqs = col.sequences()
    eng = get_engine()
    eng.set_scoring()
    eng.pack(seqs)
    n = len(seqs)
    parts, pairs = [], 0
    for r in range(4):
        info = eng.plan(rank=r, world=4)
        pairs += info["n_pairs"]
        m = torch.full((n, n), -2**31, dtype=torch.int32, device=eng.tdev)
        eng.score(scores=m)
        parts.append(m)
    assert pairs == n * (n + 1) // 2
    combined = torch.stack(parts).max(dim=0).values.cpu().numpy()     # what all_reduce(MAX) yields
    stacked = torch.stack(parts)
    assert ((stacked != -2**31).sum(dim=0) == 1).all()                # every entry owned by exactly one rank
    assert (combined == oracle.all_pairs(seqs)).all()


# ---- full-size properties (config 3 scale is too big for the oracle: sample + invariants) ---
def test_config3_scale_sample_and_invariants():
    col = synth.ks_like(2000)
    seqs = col.sequences()
    res = domain_pair_scores(seqs)
    m = res.scores
    assert (m == m.T).all() and (m >= 0).all()
    sub = oracle.BLOSUM62.astype(int)
    codes = [oracle.encode(s) for s in seqs]
    assert (np.diag(m) == np.array([int(sub[c, c].sum()) for c in codes])).all()   # self score = sum diag
    assert (m <= np.minimum.outer(np.diag(m), np.diag(m))).all()                   # s(d,e) <= min self
    rng = np.random.default_rng(0)
    pi, pj = rng.integers(0, 2000, 3000), rng.integers(0, 2000, 3000)
    assert (m[pi, pj] == oracle.pair_list(seqs, pi, pj)).all()
    assert res.plan["cells"] == col.pair_stats()["cells"]


def test_e2e_host_call_pipelined_blocks(monkeypatch):
    """bgca_score_all_host with several kernel runs: row blocks are permuted, copied and scattered
    to the host matrix behind each run while later runs compute.  Mixed lengths (many runs),
    repeated calls of different sizes on one engine, and the unpipelined path for comparison."""
    rng = np.random.default_rng(77)
    eng = get_engine()
    eng.set_scoring()
    for n, lo, hi in ((300, 20, 700), (150, 200, 1400), (40, 30, 60), (301, 20, 700)):
        seqs = [_rand(rng, int(k)) for k in rng.integers(lo, hi, n)]
        from bgclib_b200.engine import concat_ascii
        a, off = concat_ascii(seqs)
        want = oracle.all_pairs(seqs)
        out = eng.score_all_host(a, off)
        assert (out == want).all(), (n, lo, hi)
        monkeypatch.setenv("BGCA_NO_PIPELINE", "1")
        assert (eng.score_all_host(a, off) == want).all()
        monkeypatch.delenv("BGCA_NO_PIPELINE")
    # global mode, and a W-rich sequence that sends tiles to the 32-bit kernels (no pipelining then)
    eng.set_scoring("BLOSUM62", -12, -1, "global")
    seqs = [_rand(rng, int(k)) for k in rng.integers(50, 600, 120)]
    a, off = concat_ascii(seqs)
    assert (eng.score_all_host(a, off) == oracle.all_pairs(seqs, mode=oracle.MODE_GLOBAL)).all()
    eng.set_scoring()
    seqs = [_rand(rng, int(k)) for k in rng.integers(50, 600, 60)] + ["W" * 3000, "WY" * 1500]
    a, off = concat_ascii(seqs)
    assert (eng.score_all_host(a, off) == oracle.all_pairs(seqs)).all()


def test_e2e_host_call_matches():
    col = synth.ks_like(64, lo=100, hi=180, mean=140, sd=20)
    eng = get_engine()
    eng.set_scoring()
    out = eng.score_all_host(col.ascii, col.offsets)
    assert (out == oracle.all_pairs(col.sequences())).all()
    assert eng.last_launches() >= 1 and eng.last_dp_ms() > 0


@pytest.mark.parametrize("n,B", [(40, 5), (64, 8), (300, 36), (1000, 64), (777, 129)])
def test_k5_reduction_kernels_vs_numpy(n, B):
    """K5 second stage in isolation (vector and scalar row paths, transposed read)."""
    import torch
    rng = np.random.default_rng(n + B)
    seqs = ["ACDEFGHIKL"[: int(rng.integers(3, 10))] for _ in range(n)]
    bgc_of = rng.integers(0, B, n).astype(np.int32)
    bgc_of[: min(n, B - 1)] = np.arange(min(n, B - 1))          # most BGCs non-empty, the last one empty
    eng = get_engine()
    eng.set_scoring()
    eng.pack(seqs, type_of_seq=np.zeros(n, np.int32), bgc_of_seq=bgc_of, n_bgc=B)
    rowbest = rng.integers(0, 2000, (n, B)).astype(np.int32)
    selfsc = rng.integers(0, 3000, n).astype(np.int32)
    sim, best, selfsum = eng.reduce_bgc(torch.from_numpy(rowbest).cuda(), torch.from_numpy(selfsc).cuda())
    want_best = np.zeros((B, B), np.int64)
    np.add.at(want_best, bgc_of, rowbest.astype(np.int64))
    want_self = np.zeros(B, np.int64)
    np.add.at(want_self, bgc_of, selfsc.astype(np.int64))
    den = want_self[:, None] + want_self[None, :]
    num = (want_best + want_best.T).astype(np.float64)
    want_sim = np.where(den > 0, num / np.where(den > 0, den, 1), 0.0)
    want_sim[np.arange(B), np.arange(B)] = np.where(want_self > 0, 1.0, 0.0)
    assert (best.cpu().numpy() == want_best).all() and (selfsum.cpu().numpy() == want_self).all()
    assert (sim.cpu().numpy() == want_sim).all()


# ---- query-vs-database (rectangular) -------------------------------------------------------
@pytest.mark.parametrize("mode", ["local", "global"])
def test_query_vs_database_scores(mode):
    rng = np.random.default_rng(31)
    fam = _related(rng, 4, 8, 50, 400)
    rng.shuffle(fam)
    q, d = fam[:7], fam[7:] + [_rand(rng, n) for n in (1, 33, 700)]
    res = domain_pair_scores(q, database=d, mode=mode)
    cm = 0 if mode == "local" else 1
    full = oracle.all_pairs(q + d, mode=cm)
    assert res.scores.shape == (len(q), len(d))
    assert (res.scores == full[: len(q), len(q):]).all()
    assert res.plan["total_pairs"] == len(q) * len(d) + len(q) + len(d)
    assert len(res.db_index) == len(d)


@pytest.mark.parametrize("same_type_only", [True, False])
def test_query_vs_database_bgc_similarity(af293, ks_records, same_type_only):
    col = _af293_collection(af293, ks_records)
    ids = list(col.bgcs)
    qcol, dcol = type(col)(), type(col)()
    for k, b in enumerate(ids):
        (qcol if k % 5 == 0 else dcol).bgcs[b] = col.bgcs[b]
    res = bgc_similarity(qcol, database=dcol, same_type_only=same_type_only)
    qb, db = extract(qcol), extract(dcol)
    seqs = qb.sequences + db.sequences
    names = list(dict.fromkeys(qb.type_names + db.type_names))
    types = np.array([names.index(r.type_key) for r in qb.records + db.records])
    Bq, Bd = len(qb.bgc_ids), len(db.bgc_ids)
    bgcs = np.concatenate([qb.bgc_of_seq, db.bgc_of_seq + Bq])
    scores = oracle.all_pairs(seqs)
    # cross semantics: same-side pairs do not exist (except self)
    side = np.arange(len(seqs)) >= len(qb)
    masked = np.where((side[:, None] != side[None, :]) | np.eye(len(seqs), dtype=bool), scores, 0)
    sim, best, selfs, _ = oracle.bgc_similarity_np(masked, bgcs, types, Bq + Bd, same_type_only=same_type_only)
    assert res.bgc_ids == qb.bgc_ids and res.db_bgc_ids == db.bgc_ids
    assert (res.best == best[:Bq, Bq:]).all() and (res.best_db == best[Bq:, :Bq]).all()
    assert (res.self_score == selfs[:Bq]).all() and (res.db_self_score == selfs[Bq:]).all()
    assert (res.sim == sim[:Bq, Bq:]).all()


def test_cli_end_to_end(tmp_path, ks_records, af293):
    """The --comparative-style entry point on the reference's own KS fixture."""
    from bgclib_b200 import cli
    fa = tmp_path / "all_KS_domains.fasta"
    with open(fa, "w") as f:
        for r in ks_records:
            f.write(">" + r["identifier"] + "\n")
            s = r["sequence"]
            f.write("\n".join(s[i:i + 80] for i in range(0, len(s), 80)) + "\n")
    assert cli.main(["--fasta", str(fa), "--out", str(tmp_path / "af293"), "--domain-scores"]) == 0
    rows = (tmp_path / "af293.bgc_similarity.tsv").read_text().splitlines()
    ids = rows[0].split("\t")[1:]
    assert ids == list(dict.fromkeys(r["bgc"] for r in ks_records))
    sim = np.array([[float(x) for x in r.split("\t")[1:]] for r in rows[1:]])
    seqs = [r["sequence"] for r in ks_records]
    bgc_of = np.array([ids.index(r["bgc"]) for r in ks_records])
    want, _, _, _ = oracle.bgc_similarity_np(oracle.all_pairs(seqs), bgc_of, np.zeros(15, int), len(ids))
    assert (sim == want).all()
    assert (np.load(tmp_path / "af293.domain_scores.npy") == oracle.all_pairs(seqs)).all()
    dist_rows = (tmp_path / "af293.bgc_distance.tsv").read_text().splitlines()
    dmat = np.array([[float(x) for x in r.split("\t")[1:]] for r in dist_rows[1:]])
    assert (dmat == 1.0 - want).all() and (np.diag(dmat) == 0).all()
    # query-vs-database through the CLI
    assert cli.main(["--fasta", str(fa), "--db-fasta", str(fa), "--out", str(tmp_path / "x")]) == 0
    # alignment statistics of the pairs above a score, selected on the device
    assert cli.main(["--fasta", str(fa), "--out", str(tmp_path / "s"), "--domain-stats", "400"]) == 0
    lines = (tmp_path / "s.domain_stats.tsv").read_text().splitlines()
    sc = oracle.all_pairs(seqs)
    want_pairs = np.argwhere(np.triu(sc >= 400, k=1))
    want = oracle.stats_list(seqs, want_pairs[:, 0], want_pairs[:, 1])
    assert len(lines) == 1 + len(want_pairs) > 1
    cols = lines[0].split("\t")
    for line, w in zip(lines[1:], want):
        row = dict(zip(cols, line.split("\t")))
        assert all(int(row[k]) == int(w[k]) for k in want.dtype.names)
        assert abs(float(row["identity"]) - int(w["identities"]) / int(w["columns"])) < 1e-6


def test_cli_bgccase_bgclist_comparative(tmp_path, af293, ks_records, monkeypatch):
    """--comparative on .bgccase pickles (BGClib/BGClib.py:758-777) with a --bgclist order, on the
    GPU.  The reference package does not exist on the GPU box, so a stand-in module named BGClib
    with the same class surface (tests/fakes.py) is what the pickles load through;
    tests/test_reference_objects.py runs the loader on pickles of the REAL classes."""
    import pickle
    import sys
    import types
    import fakes
    from bgclib_b200 import cli
    mod = types.ModuleType("BGClib")
    mod.BGCCollection = fakes.FakeCollection
    monkeypatch.setitem(sys.modules, "BGClib", mod)
    col = _af293_collection(af293, ks_records)
    ids = list(col.bgcs)
    q, d = fakes.FakeCollection(), fakes.FakeCollection()
    for k, b in enumerate(ids):
        (q if k % 3 == 0 else d).bgcs[b] = col.bgcs[b]
    for name, c in (("q", q), ("d", d), ("all", col)):
        with open(tmp_path / f"{name}.bgccase", "wb") as f:
            pickle.dump(c, f)
    order = [b for b in reversed(ids) if len(extract(col.bgcs[b])) > 0][:9]
    (tmp_path / "order.tsv").write_text("".join(f"{b}\tprot\tnote\n" for b in order))
    assert cli.main(["--comparative", "--bgccase", str(tmp_path / "all.bgccase"), "--bgclist",
                     str(tmp_path / "order.tsv"), "--out", str(tmp_path / "o")]) == 0
    rows = (tmp_path / "o.bgc_similarity.tsv").read_text().splitlines()
    assert rows[0].split("\t")[1:] == order and [r.split("\t")[0] for r in rows[1:]] == order
    sim = np.array([[float(x) for x in r.split("\t")[1:]] for r in rows[1:]])
    sub = extract(col).select_bgcs(order)
    want, _, _, _ = oracle.bgc_similarity_np(oracle.all_pairs(sub.sequences), sub.bgc_of_seq, sub.type_of_seq, len(order))
    assert (sim == want).all()
    # query pickle against database pickle
    assert cli.main(["--comparative", "--bgccase", str(tmp_path / "q.bgccase"), "--db-bgccase",
                     str(tmp_path / "d.bgccase"), "--out", str(tmp_path / "x")]) == 0
    rows = (tmp_path / "x.bgc_similarity.tsv").read_text().splitlines()
    got = np.array([[float(x) for x in r.split("\t")[1:]] for r in rows[1:]])
    ref_res = bgc_similarity(q, database=d)
    assert rows[0].split("\t")[1:] == ref_res.db_bgc_ids and (got == ref_res.sim).all()


# ---- K7: alignment statistics without traceback --------------------------------------------
def _stat_rows(st):
    return np.stack([st[k] for k in oracle.STATS_DTYPE.names], axis=1)


@pytest.mark.parametrize("mode", ["local", "global"])
def test_stats_ks_golden(ks_records, mode):
    import hashlib as _h
    import json
    from conftest import GOLDEN
    from bgclib_b200 import domain_pair_stats
    gold = json.loads((GOLDEN / "ks_stats.json").read_text())
    seqs = [r["sequence"] for r in ks_records]
    n = len(seqs)
    pairs = np.stack(np.divmod(np.arange(n * n), n), axis=1)
    res = domain_pair_stats(seqs, pairs, mode=mode)
    rows = _stat_rows(res.stats).astype("<i4")
    assert (rows == np.array(gold[mode]["records"], dtype=np.int32)).all()
    assert _h.sha256(rows.tobytes()).hexdigest() == gold[mode]["sha256_int32_le"]
    assert (res.identity.reshape(n, n).diagonal() == 1.0).all()


@pytest.mark.parametrize("mode", ["local", "global"])
@pytest.mark.parametrize("lo,hi", [(1, 12), (30, 90), (200, 300), (250, 530), (700, 1100), (1023, 1023), (1020, 1026)])
def test_stats_random_bit_exact(mode, lo, hi):
    """Single-pass and multi-pass (query > 256 rows) sweeps, ties, B/Z/X letters, empty alignments."""
    from bgclib_b200 import domain_pair_stats
    rng = np.random.default_rng(lo + 3 * hi + (mode == "global"))
    letters = AA + "BZX" if lo < 100 else AA
    seqs = _related(rng, 3, 5, lo, hi, letters) + [_rand(rng, rng.integers(lo, hi + 1), "WP"[k % 2]) for k in range(4)]
    n = len(seqs)
    pairs = np.stack(np.divmod(np.arange(n * n), n), axis=1)
    res = domain_pair_stats(seqs, pairs, mode=mode)
    cm = oracle.MODE_LOCAL if mode == "local" else oracle.MODE_GLOBAL
    want = oracle.stats_list(seqs, pairs[:, 0], pairs[:, 1], mode=cm)
    assert (_stat_rows(res.stats) == _stat_rows(want)).all()
    assert (res.stats["score"] == domain_pair_scores(seqs, mode=mode).scores.reshape(-1)).all()


def test_stats_other_gap_scores_and_default_pairs():
    from bgclib_b200 import domain_pair_stats
    rng = np.random.default_rng(77)
    seqs = _related(rng, 2, 6, 40, 160)
    for o, e in ((-11, -1), (-5, -2), (-3, 0), (-1, -1)):
        res = domain_pair_stats(seqs, open_gap_score=o, extend_gap_score=e)
        assert len(res.pairs) == len(seqs) * (len(seqs) - 1) // 2 and (res.pairs[:, 0] < res.pairs[:, 1]).all()
        want = oracle.stats_list(seqs, res.pairs[:, 0], res.pairs[:, 1], open_gap=o, extend_gap=e)
        assert (_stat_rows(res.stats) == _stat_rows(want)).all()


def test_stats_rejects_sequences_beyond_32767():
    from bgclib_b200 import domain_pair_stats
    with pytest.raises(ValueError, match="32767"):
        domain_pair_stats(["A" * 40000, "ACD"], [[0, 1]])


def test_shard_rows_under_a_process_group():
    """The reduce-scatter path (row-sharded result) on a one-rank NCCL group."""
    import socket
    import torch.distributed as dist
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    rng = np.random.default_rng(12)
    seqs = _related(rng, 3, 5, 40, 200)
    want = oracle.all_pairs(seqs)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        full = domain_pair_scores(seqs)
        part = domain_pair_scores(seqs, shard_rows=True)
    finally:
        dist.destroy_process_group()
    assert (full.scores == want).all() and full.rows is None
    assert part.rows == (0, len(seqs)) and (part.scores == want).all()


def test_resumable_rounds_with_checkpoint(tmp_path, af293, ks_records):
    """rounds + checkpoint: an interrupted run continues after the last finished round and ends
    bit-identical to a single pass; a changed parameter invalidates the file."""
    col = _af293_collection(af293, ks_records)
    whole = bgc_similarity(col)
    ck = tmp_path / "run"
    with pytest.raises(KeyboardInterrupt):
        bgc_similarity(col, rounds=5, checkpoint=ck, _stop_after_rounds=2)
    z = np.load(f"{ck}.rank0.npz")
    assert sorted(z["done"].tolist()) == [0, 1]
    part = bgc_similarity(col, rounds=5, checkpoint=ck)
    assert (part.best == whole.best).all() and (part.sim == whole.sim).all()
    assert part.plan["n_pairs"] == whole.plan["n_pairs"] and part.plan["cells"] == whole.plan["cells"]
    assert sorted(np.load(f"{ck}.rank0.npz")["done"].tolist()) == [0, 1, 2, 3, 4]
    # other gap scores: the old file is ignored, not continued
    other = bgc_similarity(col, rounds=5, checkpoint=ck, open_gap_score=-11)
    assert (other.best == bgc_similarity(col, open_gap_score=-11).best).all()


def test_stats_for_pairs_above_a_score():
    """Two-stage use: all-vs-all scores, then statistics only where the score reaches a threshold."""
    from bgclib_b200 import domain_pair_stats
    rng = np.random.default_rng(21)
    seqs = _related(rng, 4, 5, 60, 220) + [_rand(rng, 100) for _ in range(6)]
    want_scores = oracle.all_pairs(seqs)
    thr = 120
    res = domain_pair_stats(seqs, min_score=thr)
    iu = np.triu_indices(len(seqs), k=1)
    sel = want_scores[iu] >= thr
    want_pairs = np.stack(iu, axis=1)[sel]
    assert 0 < len(want_pairs) < len(sel)                      # the threshold separates the families
    assert (res.pairs == want_pairs).all()
    want = oracle.stats_list(seqs, want_pairs[:, 0], want_pairs[:, 1])
    assert (_stat_rows(res.stats) == _stat_rows(want)).all()
    assert (res.stats["score"] >= thr).all()


def test_stats_with_known_scores_right_low_and_high():
    """bgca_pair_stats_hinted: with the true scores as hints, with hints that are too low (more
    candidate cells, same answer) and with hints that are too high (no candidate -> detected, run
    again without hints) the records equal the oracle's; global mode ignores hints."""
    rng = np.random.default_rng(31)
    seqs = _related(rng, 5, 6, 40, 520) + [_rand(rng, k) for k in (1, 2, 33, 129, 257, 385, 449, 513, 1023)]
    n = len(seqs)
    iu = np.triu_indices(n, k=0)
    pi, pj = iu[0].astype(np.int32), iu[1].astype(np.int32)
    eng = get_engine()
    for mode, cm in (("local", oracle.MODE_LOCAL), ("global", oracle.MODE_GLOBAL)):
        eng.set_scoring("BLOSUM62", -12, -1, mode)
        eng.pack(seqs)
        want = oracle.stats_list(seqs, pi, pj, mode=cm)
        true = want["score"].astype(np.int32)
        for hints in (true, np.maximum(true - rng.integers(0, 40, len(true)).astype(np.int32), 0),
                      true + (rng.random(len(true)) < 0.05) * 7, np.zeros_like(true)):
            got = eng.pair_stats(pi, pj, known_scores=hints.astype(np.int32)).cpu().numpy()
            got = np.ascontiguousarray(got).view(oracle.STATS_DTYPE).reshape(-1)
            assert (_stat_rows(got) == _stat_rows(want)).all(), mode


@pytest.mark.parametrize("same_type_only", [False, True])
def test_stats_selected_on_the_device(same_type_only):
    """bgca_select_pairs + bgca_pair_stats_selected: pairs i < j above a score chosen on the device
    from the score matrix (row-major order, optional same-type restriction), statistics with the
    selected scores as hints; pairs and records equal the oracle's, in both modes."""
    import torch
    rng = np.random.default_rng(33 + same_type_only)
    seqs = _related(rng, 6, 7, 50, 500) + [_rand(rng, k) for k in (30, 90, 300, 700)]
    n = len(seqs)
    types = rng.integers(0, 3, n).astype(np.int32)
    eng = get_engine()
    for mode, cm, thr in (("local", oracle.MODE_LOCAL, 90), ("global", oracle.MODE_GLOBAL, -400)):
        eng.set_scoring("BLOSUM62", -12, -1, mode)
        eng.pack(seqs, type_of_seq=types if same_type_only else None)
        eng.plan(same_type_only=same_type_only)
        scores = torch.full((n, n), -2**31, dtype=torch.int32, device="cuda")
        eng.score(scores=scores)
        pairs, out = eng.pair_stats_above(scores, thr, same_type_only)
        full = oracle.all_pairs(seqs, mode=cm)
        keep = np.triu(full >= thr, k=1)
        if same_type_only:
            keep &= types[:, None] == types[None, :]
        want_pairs = np.argwhere(keep)
        assert 0 < len(want_pairs) < n * (n - 1) // 2
        assert (pairs.cpu().numpy() == want_pairs).all()
        want = oracle.stats_list(seqs, want_pairs[:, 0], want_pairs[:, 1], mode=cm)
        got = np.ascontiguousarray(out.cpu().numpy()).view(oracle.STATS_DTYPE).reshape(-1)
        assert (_stat_rows(got) == _stat_rows(want)).all(), mode
    # nothing above the threshold: empty result, and the selection is spent
    pairs, out = eng.pair_stats_above(scores, 10**6, same_type_only)
    assert pairs.shape == (0, 2) and out.shape[0] == 0


def test_outputs_stay_inside_their_buffers():
    """Guard rows around every caller-owned output (compute-sanitizer is not available on the
    pool): nothing outside the declared shapes may be written, whatever the kernel variant."""
    import torch
    rng = np.random.default_rng(5)
    seqs = _related(rng, 4, 6, 20, 400) + [_rand(rng, k) for k in (1, 5, 700, 1500)]
    n, B = len(seqs), 7
    bgc_of = rng.integers(0, B, n).astype(np.int32)
    eng = get_engine()
    SENT = -123456789
    for mode in ("local", "global"):
        eng.set_scoring(mode=mode)
        eng.pack(seqs, type_of_seq=np.zeros(n, np.int32), bgc_of_seq=bgc_of, n_bgc=B)
        eng.plan()
        big = torch.full((n + 2, n), SENT, dtype=torch.int32, device="cuda")
        rb = torch.full((n + 2, B), SENT, dtype=torch.int32, device="cuda")
        ss = torch.full((n + 2,), SENT, dtype=torch.int32, device="cuda")
        big[1:n + 1].fill_(-2**31)
        rb[1:n + 1].zero_()
        ss[1:n + 1].zero_()
        eng.score(scores=big[1:n + 1], rowbest=rb[1:n + 1], selfsc=ss[1:n + 1])
        torch.cuda.synchronize()
        for t in (big, rb, ss):
            assert (t[0] == SENT).all() and (t[-1] == SENT).all()
        cm = oracle.MODE_LOCAL if mode == "local" else oracle.MODE_GLOBAL
        assert (big[1:n + 1].cpu().numpy() == oracle.all_pairs(seqs, mode=cm)).all()


@pytest.mark.parametrize("same_type_only", [True, False])
def test_persistent_packed_database(af293, ks_records, same_type_only):
    """PackedDatabase: the database stays packed on the device, query sets are appended behind it.
    Results equal the one-shot query-vs-database call (itself checked against the oracle), for two
    different query sets in a row; a bad query leaves the database usable."""
    from bgclib_b200 import PackedDatabase
    col = _af293_collection(af293, ks_records)
    ids = list(col.bgcs)
    dcol, q1, q2 = type(col)(), type(col)(), type(col)()
    for k, b in enumerate(ids):
        (q1 if k % 5 == 0 else q2 if k % 5 == 1 else dcol).bgcs[b] = col.bgcs[b]
    db = PackedDatabase(dcol)
    assert len(db) == len(extract(dcol))
    for qcol in (q1, q2, q1):
        got = bgc_similarity(qcol, database=db, same_type_only=same_type_only)
        want = bgc_similarity(qcol, database=dcol, same_type_only=same_type_only)
        assert got.bgc_ids == want.bgc_ids and got.db_bgc_ids == want.db_bgc_ids
        assert (got.best == want.best).all() and (got.best_db == want.best_db).all()
        assert (got.self_score == want.self_score).all() and (got.db_self_score == want.db_self_score).all()
        assert (got.sim == want.sim).all()
    # the rectangular domain-pair matrix from the same resident database, in both modes
    qb, dbb = extract(q1), extract(dcol)
    for mode, cm in (("local", oracle.MODE_LOCAL), ("global", oracle.MODE_GLOBAL)):
        dps = domain_pair_scores(q1, database=db, mode=mode)
        full = oracle.all_pairs(qb.sequences + dbb.sequences, mode=cm)
        assert (dps.scores == full[: len(qb), len(qb):]).all()
    with pytest.raises(ValueError, match="alphabet"):
        bgc_similarity(["ACDEFGHIKL", "ACDUFG"], database=db)
    again = bgc_similarity(q2, database=db, same_type_only=same_type_only)
    assert (again.sim == bgc_similarity(q2, database=dcol, same_type_only=same_type_only).sim).all()
    with pytest.raises(ValueError, match="scoring"):
        bgc_similarity(q1, database=db, open_gap_score=-11)
