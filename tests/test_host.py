"""Host-side logic that needs no GPU: the C ABI loads and exports what
include/bgca.h declares, the K6 planner's partition properties, extraction from
(fake) reference objects, the synthetic generators, and the pure-Python model of
the kernel's systolic data flow against the oracle."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

import oracle
from bgclib_b200 import engine as eng
from bgclib_b200 import extract, plan_host, synth
from bgclib_b200.alphabet import resolve_matrix
from fakes import FakeProteinCollection, FakeProtein, build_collection
from systolic_model import systolic_score, systolic_stream

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "bgca.h").read_text()
    declared = set(re.findall(r"\b(bgca_[a-z0-9_]+)\s*\(", header))
    assert declared == set(eng.EXPORTED_SYMBOLS)
    lib = eng.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bgca_version() == 105


def test_packed_batch_stays_inside_one_4gib_window():
    """The DP kernels add to the low half of their token pointer only (gotoh_pair16.cuh,
    advance_lo32): whatever cudaMalloc returns, the packed batch must start where it crosses no
    4 GiB boundary, and that start must lie inside an allocation of twice its size."""
    lib = eng.load_library()
    G = 1 << 32
    rng = np.random.default_rng(7)
    cases = [(0x7F0000000, 1), (0x7FFFFFFFF, 1), (0x7FFFFFFFF, 2), (0x7FFFF0000, 0x10000), (0x7FFFF0000, 0x10001),
             (5 * G, G), (5 * G + 512, G), (5 * G - 512, 42 << 20), (3 * G - 1, G)]
    cases += [(int(rng.integers(1, 1 << 46)) & ~0xFF, int(rng.integers(1, 3 << 30))) for _ in range(2000)]
    for raw, nbytes in cases:
        out = ctypes.c_uint64()
        assert lib.bgca_window4g_start(raw, nbytes, ctypes.byref(out)) == 0
        start = out.value
        assert start >> 32 == (start + nbytes - 1) >> 32, (raw, nbytes)      # one window
        assert raw <= start and start + nbytes <= raw + 2 * nbytes, (raw, nbytes)   # inside the 2x allocation
        if raw >> 32 == (raw + nbytes - 1) >> 32:
            assert start == raw                                               # nothing to do: nothing moved
    out = ctypes.c_uint64()
    assert lib.bgca_window4g_start(G, G + 1, ctypes.byref(out)) == eng.BGCA_ERR_INVALID


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = eng.load_library()
    h = ctypes.c_void_p()
    rc = lib.bgca_create(0, ctypes.byref(h))
    assert rc == eng.BGCA_ERR_CUDA and b"no CPU fallback" in lib.bgca_last_error()
    with pytest.raises(RuntimeError):
        eng.Engine(0)
    out = (ctypes.c_double * 10)()
    assert lib.bgca_dpx_probe(0, out, 10) == eng.BGCA_ERR_CUDA


def _expand(tiles):
    """tiles -> list of unique (i<=j) pairs exactly as the kernel epilogue emits them."""
    pairs = []
    for t in tiles:
        for s in range(t["s_begin"], t["s_end"]):
            for q in (t["qa"], t["qb"]):
                if q >= 0 and s >= q:
                    pairs.append((q, s))
    return pairs


@pytest.mark.parametrize("n,world", [(1, 1), (2, 1), (7, 1), (64, 1), (301, 2), (500, 8)])
def test_planner_every_pair_exactly_once(n, world):
    rng = np.random.default_rng(n)
    lens = np.sort(rng.integers(30, 700, size=n)).astype(np.int32)
    seen = []
    total = None
    for r in range(world):
        tiles, info = plan_host(lens, rank=r, world=world)
        pairs = _expand(tiles)
        assert info["n_pairs"] == len(pairs)
        assert info["cells"] == sum(int(lens[i]) * int(lens[j]) for i, j in pairs)
        total = info["total_pairs"]
        seen += pairs
    assert total == n * (n + 1) // 2
    assert len(seen) == total and len(set(seen)) == total
    assert set(seen) == {(i, j) for i in range(n) for j in range(i, n)}


def test_planner_same_type_blocks():
    rng = np.random.default_rng(5)
    types = np.sort(rng.integers(0, 4, size=200)).astype(np.int32)
    lens = np.zeros(200, np.int32)
    for t in range(4):
        m = types == t
        lens[m] = np.sort(rng.integers(60, 500, size=int(m.sum())))
    tiles, info = plan_host(lens, types, same_type_only=True)
    pairs = set(_expand(tiles))
    want = {(i, j) for i in range(200) for j in range(i, 200) if types[i] == types[j]}
    assert pairs == want and info["total_pairs"] == len(want)


def test_planner_balance_and_variants():
    lens = np.sort(np.random.default_rng(0).integers(360, 481, size=4000)).astype(np.int32)
    loads = []
    for r in range(8):
        tiles, info = plan_host(lens, rank=r, world=8)
        loads.append(info["cells_computed"])
        for v in set(tiles["variant"].tolist()):
            G, K = (v >> 8) & 0xFF, v & 0xFF
            sel = tiles[tiles["variant"] == v]
            lq = np.maximum(lens[sel["qa"]], np.where(sel["qb"] >= 0, lens[np.maximum(sel["qb"], 0)], 0))
            assert G in (8, 16, 32) and (G * K >= lq).all()
    assert max(loads) / (sum(loads) / 8) < 1.01          # LPT on exact cell counts
    # pairs whose scores can leave int16 go to the 32-bit form of the multi-pass kernel, one query
    # per tile (0x20000), every pair still planned exactly once
    tiles, info = plan_host(np.array([100, 200, 1500, 3200], np.int32))
    assert info["total_pairs"] == 10 and info["n_fallback32"] >= 1
    w32 = [t for t in tiles if int(t["variant"]) & 0x20000]
    assert w32 and all(int(t["qb"]) == -1 and int(t["variant"]) & 0x10000 for t in w32)
    # the int16 decision is taken per query PAIR (1500, 3200) and subject chunk: both go, each alone
    assert {(int(t["qa"]), int(t["s_begin"]), int(t["s_end"])) for t in w32} == {(2, 2, 4), (3, 3, 4)}
    # short queries with scores beyond int16 stay on the pair-list kernel (variant 0)
    tiles, info = plan_host(np.array([250, 260], np.int32), max_abs_sub=127, open_gap=-12, extend_gap=-1)
    assert all(int(t["variant"]) == 0 for t in tiles) and info["n_fallback32"] == info["n_tiles"]
    # global mode: a huge gap-extension penalty leaves int16 -> fallback
    _, info_g = plan_host(np.array([400, 420], np.int32), mode=1, open_gap=-40, extend_gap=-40)
    assert info_g["n_fallback32"] == info_g["n_tiles"]
    # the int16 guard looks at the LONGEST subject of a tile, not at its last one: with types packed
    # but all pairs planned, a tile crosses type runs and its last subject can be short
    lens4, types4 = np.array([100, 110, 30000, 50], np.int32), np.array([0, 0, 0, 1], np.int32)
    tiles, _ = plan_host(lens4, types4, same_type_only=False, mode=1)
    for t in tiles:
        if t["s_begin"] <= 2 < t["s_end"]:
            assert t["variant"] == 0 or int(t["variant"]) & 0x20000, \
                "a 30000-residue subject in global mode must take a 32-bit kernel"
    # a persistent database in front of two queries longer than one pass of the packed kernel: the
    # cross tiles take the multi-pass variants and account for every query x database pair
    lens_db = np.array([100, 120, 150, 200, 220, 300, 2500, 2600], np.int32)
    tiles, info = plan_host(lens_db, cross_only=2, n_db_first=6)
    assert info["n_fallback32"] == 0 and info["total_pairs"] == 2 * 6 + 2
    assert all(int(t["variant"]) & 0x10000 for t in tiles)              # BGCA_VARIANT_MP
    assert {int(t["variant"]) & 0xFFFF for t in tiles} == {(32 << 8) | 28}   # ceil(2600 / (32 * 3)) rows per lane
    # ... and the 32-bit kernel when the scores can leave int16 (11 * min length > 30000)
    lens_big = np.array([2800, 2900, 3000, 3100, 3200, 3300, 3400, 3500], np.int32)
    tiles, info = plan_host(lens_big, cross_only=2, n_db_first=6)
    assert info["n_fallback32"] == info["n_tiles"] and info["total_pairs"] == 2 * 6 + 2


def test_multi_rank_large_plan_cuts_shorter_tiles_and_still_covers_everything():
    """Large plans shared by several ranks cut shorter tiles for kernel runs that would leave a
    rank only a few tiles per CTA slot; the shares still partition the pairs and cells exactly."""
    col = synth.mibig_scale(250, seed=11)                 # ~6 000 mixed domains, 6 types: a "large" plan
    order = np.lexsort((col.lengths, col.type_of_seq))
    L, T = col.lengths[order].astype(np.int32), col.type_of_seq[order].astype(np.int32)
    one, info1 = plan_host(L, T, same_type_only=True)
    pairs = cells = 0
    n_tiles = 0
    seen = np.zeros(len(L), dtype=np.int64)               # subjects visited per query row, summed over ranks
    for r in range(8):
        tiles, info = plan_host(L, T, same_type_only=True, rank=r, world=8)
        pairs += info["n_pairs"]; cells += info["cells"]; n_tiles += info["n_tiles"]
        assert info["total_pairs"] == info1["total_pairs"] and info["total_cells"] == info1["total_cells"]
        np.add.at(seen, tiles["qa"], tiles["s_end"] - tiles["s_begin"])
    assert pairs == info1["total_pairs"] == col.pair_stats(same_type_only=True)["pairs"]
    assert cells == info1["total_cells"] == col.pair_stats(same_type_only=True)["cells"]
    assert n_tiles > info1["n_tiles"]                     # shorter tiles than the single-rank plan
    seen1 = np.zeros(len(L), dtype=np.int64)
    np.add.at(seen1, one["qa"], one["s_end"] - one["s_begin"])
    assert (seen == seen1).all()                          # every query row meets the same subjects


def test_extract_from_collection():
    p1 = "M" + "ACDEFGHIKL" * 6
    p2 = "MNPQRSTVWY" * 5
    spec = {
        "bgcA": [("bgcA~L0+CDS1", p1, [("ketoacyl-synt", "KS", 1, 31), ("Acyl_transf_1", "AT", 31, 61)])],
        "bgcB": [("bgcB~L0+CDS1", p2, [("ketoacyl-synt", "", 0, 25)]),
                 ("bgcB~L0+CDS2", p1, [("PP-binding", "T/ACP", 5, 5)])],      # empty slice
        "bgcC": [],
    }
    col = build_collection(spec)
    b = extract(col, alias={"ketoacyl-synt": "KS"})
    assert b.bgc_ids == ["bgcA", "bgcB", "bgcC"]
    assert [r.type_key for r in b.records] == ["KS", "AT", "KS"]
    assert b.sequences == [p1[1:31], p1[31:61], p2[0:25]]
    assert b.bgc_of_seq.tolist() == [0, 0, 1]
    assert b.type_of_seq.tolist() == [0, 1, 0] and b.type_names == ["KS", "AT"]
    assert len(b.skipped) == 1
    # label precedence without external alias: own alias, else ID (BGClib.py:1991-2002)
    b2 = extract(col)
    assert [r.type_key for r in b2.records] == ["KS", "AT", "ketoacyl-synt"]
    # cbp_only honours CBPcontent
    col2 = build_collection(spec, cbp=False)
    assert len(extract(col2, cbp_only=True)) == 0 and len(extract(col2, cbp_only=False)) == 3
    # domain type filter
    assert len(extract(col, alias={"ketoacyl-synt": "KS"}, domain_types={"KS"})) == 2


def test_synthetic_object_collection_round_trips_through_extract(tmp_path):
    """synth.as_collection lays a synthetic collection out as BGC / protein / domain objects;
    extract() must give back exactly its sequences, BGC and type assignment, in order — in span
    form (each protein's residues once) — and DomainBatch.save/load keep all of it."""
    from bgclib_b200.extract import DomainBatch
    col = synth.mibig_scale(30, seed=9)
    objs = synth.as_collection(col)
    b = extract(objs)
    assert b.sequences == col.sequences() and (b.bgc_of_seq == col.bgc_of_seq).all()
    assert [b.type_names[t] for t in b.type_of_seq] == [col.type_names[t] for t in col.type_of_seq]
    assert b.sequences == [d.get_sequence() for g in objs.bgcs.values() for p in g.protein_list for d in p.domain_list]
    blob, starts, lengths = b.spans()
    assert len(blob) == sum(len(p.sequence) for g in objs.bgcs.values() for p in g.protein_list)
    assert (lengths == col.lengths).all()
    b.save(tmp_path / "b.npz")
    again = DomainBatch.load(tmp_path / "b.npz")
    assert again.sequences == b.sequences and again.bgc_ids == b.bgc_ids and again.type_names == b.type_names
    assert [r.protein_identifier for r in again.records] == [r.protein_identifier for r in b.records]
    sub = b.select_bgcs(b.bgc_ids[3:6][::-1])
    assert sub.bgc_ids == b.bgc_ids[3:6][::-1] and len(sub) == int(np.isin(b.bgc_of_seq, [3, 4, 5]).sum())


def test_extract_protein_units_and_strings(ks_records):
    pc = FakeProteinCollection()
    for r in ks_records:
        pc.proteins[r["identifier"]] = FakeProtein(r["identifier"], r["sequence"])
    b = extract(pc, unit="protein")
    assert len(b) == 15 and b.sequences[0] == ks_records[0]["sequence"]
    b2 = extract(["acd efg\n", "WWW"])
    assert b2.sequences == ["ACDEFG", "WWW"]


def test_resolve_matrix_and_concat():
    assert (resolve_matrix("blosum62") == oracle.BLOSUM62).all()
    with pytest.raises(ValueError):
        resolve_matrix("PAM999")
    bad = np.array(oracle.BLOSUM62, dtype=np.int64)
    bad[0, 1] = 3
    with pytest.raises(ValueError):
        resolve_matrix(bad)
    a, off = eng.concat_ascii(["AC", "D", "EFG"])
    assert a.tobytes() == b"ACDEFG" and off.tolist() == [0, 2, 3, 6]


def test_synth_is_deterministic_and_in_range():
    a = synth.ks_like(300)
    b = synth.ks_like(300)
    assert (a.ascii == b.ascii).all() and (a.offsets == b.offsets).all()
    L = a.lengths
    assert L.min() >= 360 and L.max() <= 480 and abs(L.mean() - 420) < 10
    assert set(a.ascii.tobytes().decode()) <= set(synth.AA20)
    st = a.pair_stats()
    assert st["pairs"] == 300 * 301 // 2
    assert st["cells"] == sum(int(L[i]) * int(L[j]) for i in range(300) for j in range(i, 300))
    m = synth.mixed_collection(20, 1, {"A": 1, "T/ACP": 2})
    assert m.lengths.min() >= 70 and m.lengths.max() <= 600 and m.n_bgc == 20
    # family structure: related members score far above unrelated background
    seqs = synth.ks_like(120, family_size=60).sequences()
    sc = oracle.all_pairs(seqs[:40])
    assert np.triu(sc, 1).max() > 400


@pytest.mark.parametrize("local", [True, False])
def test_systolic_model_equals_oracle(local):
    """The kernel's recurrence / boundary / capture logic, modelled in Python."""
    rng = np.random.default_rng(42)
    sub = oracle.BLOSUM62.astype(int).tolist()
    for _ in range(40):
        la, ls = int(rng.integers(1, 36)), int(rng.integers(1, 50))
        q = rng.integers(0, 24, la).astype(np.uint8)
        s = rng.integers(0, 24, ls).astype(np.uint8)
        if rng.random() < 0.5:
            s = np.concatenate([q, s])[:ls].astype(np.uint8)
        for G, K in ((8, 5), (4, 9), (16, 3)):
            if G * K < la:
                continue
            for o, e in ((-12, -1), (-4, -2)):
                want = oracle.score(q, s, mode=0 if local else 1, open_gap=o, extend_gap=e)
                assert systolic_score(q.tolist(), s.tolist(), sub, o, e, local, G, K) == want


@pytest.mark.parametrize("local", [True, False])
def test_systolic_stream_of_subjects(local):
    """Back-to-back subjects through one group: boundary tokens, state reset, wind-up lag,
    score collection — including subjects shorter than the group is wide."""
    rng = np.random.default_rng(7)
    sub = oracle.BLOSUM62.astype(int).tolist()
    for G, K in ((4, 6), (8, 3), (16, 2)):
        for _ in range(6):
            la = int(rng.integers(1, G * K + 1))
            q = rng.integers(0, 24, la).astype(np.uint8)
            subjects = []
            for n in rng.integers(1, 40, size=int(rng.integers(1, 9))):
                s = rng.integers(0, 24, int(n)).astype(np.uint8)
                if rng.random() < 0.5:
                    s = np.concatenate([q, s])[: int(n)].astype(np.uint8)
                subjects.append(s)
            subjects += [np.array([3], np.uint8), np.array([5, 5], np.uint8)]     # shorter than G
            got = systolic_stream(q.tolist(), [s.tolist() for s in subjects], sub, -12, -1, local, G, K)
            want = [oracle.score(q, s, mode=0 if local else 1) for s in subjects]
            assert got == want


def test_planner_query_vs_database():
    """cross_only: every query x database pair exactly once (+ one self pair per sequence)."""
    rng = np.random.default_rng(3)
    # sorted order = (type, role, length)
    types, roles, lens = [], [], []
    for t in range(3):
        for role, cnt in ((0, int(rng.integers(0, 9))), (1, int(rng.integers(0, 40)))):
            types += [t] * cnt
            roles += [role] * cnt
            lens += sorted(rng.integers(40, 500, size=cnt).tolist())
    types, roles, lens = (np.asarray(x, np.int32) for x in (types, roles, lens))
    n = len(lens)
    for same_type in (True, False):
        if not same_type:       # one block: re-sort by (role, length)
            o = np.lexsort((lens, roles))
            ty, ro, le = types[o], roles[o], lens[o]
        else:
            ty, ro, le = types, roles, lens
        seen = []
        for r in range(2):
            tiles, info = plan_host(le, ty, ro, same_type_only=same_type, cross_only=True, rank=r, world=2)
            for t in tiles:
                for s in range(t["s_begin"], t["s_end"]):
                    for q in (t["qa"], t["qb"]):
                        if q < 0 or s < q:
                            continue
                        if q == s or ro[q] != ro[s]:          # what emit_pair keeps in cross mode
                            seen.append((q, s))
        want = {(i, i) for i in range(n)}
        want |= {(i, j) for i in range(n) for j in range(i + 1, n)
                 if ro[i] != ro[j] and (not same_type or ty[i] == ty[j])}
        assert len(seen) == len(set(seen)) and set(seen) == want
        assert info["total_pairs"] == len(want)
    # a (type, role, length) order planned as ONE block would silently drop queries: refused
    with pytest.raises(ValueError, match="queries-first"):
        plan_host(lens, types, roles, same_type_only=False, cross_only=True)


def test_cli_fasta_reader_and_headers(tmp_path, ks_records):
    from bgclib_b200 import cli
    fa = tmp_path / "d.fasta"
    lines = []
    for r in ks_records[:3]:
        lines.append(">" + r["identifier"])
        s = r["sequence"]
        lines += [s[i:i + 80].lower() if i == 0 else s[i:i + 80] for i in range(0, len(s), 80)]
        lines.append("")
    lines += [">" + ks_records[0]["identifier"], "AAAA"]          # repeated header: dropped
    fa.write_text("\n".join(lines) + "\n")
    recs = cli.read_fasta(fa)
    assert [h for h, _ in recs] == [r["identifier"] for r in ks_records[:3]]
    assert [s for _, s in recs] == [r["sequence"] for r in ks_records[:3]]
    assert cli.parse_domain_header(ks_records[0]["identifier"]) == \
        ("CM000170.1.region002", "CM000170.1.region002~L0+CDS9", "KS")
    assert cli.parse_domain_header("x~L0+CDS1_T/ACP12 ProteinId:a GeneId:")[2] == "T/ACP"
    b = cli.batch_from_fasta(fa)
    assert b.bgc_ids == [r["bgc"] for r in ks_records[:3]] and b.type_names == ["KS"]
    alias = tmp_path / "alias.tsv"
    alias.write_text("# c\nketoacyl-synt\tKS\tPfam\n\nAMP-binding\tA\n")
    assert cli.read_alias_tsv(alias) == {"ketoacyl-synt": "KS", "AMP-binding": "A"}


def test_domain_batch_round_trips_through_disk(tmp_path):
    spec = {"b1": [("b1~L0+CDS1", "ACDEFGHIKLMNPQRSTVWY" * 3, [("ketoacyl-synt", "KS", 0, 30), ("x", "", 30, 55)])],
            "b2": [("b2~L0+CDS1", "WWWWYYYY" * 5, [("PP-binding", "T/ACP", 2, 38)])]}
    b = extract(build_collection(spec))
    b.save(tmp_path / "batch.npz")
    from bgclib_b200 import DomainBatch
    c = DomainBatch.load(tmp_path / "batch.npz")
    assert c.sequences == b.sequences and c.records == b.records and c.bgc_ids == b.bgc_ids
    assert (c.type_of_seq == b.type_of_seq).all() and (c.bgc_of_seq == b.bgc_of_seq).all()


def test_split_domain_merge_matches_reference_fix_core_split_domains():
    """merged_domain_view against the REAL fix_core_split_domains (BGCtoolkit.py:575-687) run on
    real BGCProtein/BGCDomain objects by tests/golden/make_split_golden.py."""
    import json
    from types import SimpleNamespace
    from conftest import GOLDEN
    from bgclib_b200.extract import merged_domain_view, split_domain_groups
    cases = json.loads((GOLDEN / "split_domains.json").read_text())["cases"]
    assert len(cases) == 400
    merged = 0
    for c in cases:
        p = SimpleNamespace(identifier=c["identifier"], sequence="A" * 3000, domain_list=[])
        p.domain_list = [SimpleNamespace(protein=p, ID=d[0], alias="", ali_from=d[1], ali_to=d[2], hmm_from=d[3],
                                         hmm_to=d[4], score=d[5]) for d in c["domains"]]
        before = list(p.domain_list)
        view = merged_domain_view(p)
        got = [[v.ID, v.ali_from, v.ali_to, v.hmm_from, v.hmm_to, v.score] for v in view]
        assert got == c["after"], c["identifier"]
        assert p.domain_list == before                                   # the protein is untouched
        assert sum(v.n_fragments for v in view) == len(before)
        merged += any(v.n_fragments > 1 for v in view)
        for run in split_domain_groups(before):
            assert len(run) >= 2 and len({before[k].ID for k in run}) == 1
    assert merged > 100


def test_extract_with_merged_split_domains():
    from types import SimpleNamespace
    from bgclib_b200 import extract
    p = SimpleNamespace(identifier="P1", sequence="ACDEFGHIKLMNPQRSTVWY" * 5, domain_list=[], parent_cluster_id="B1",
                        protein_type="PKS")
    mk = lambda ID, a0, a1, h0, h1: SimpleNamespace(protein=p, ID=ID, alias="", ali_from=a0, ali_to=a1,  # noqa: E731
                                                    hmm_from=h0, hmm_to=h1, score=1.0,
                                                    get_sequence=lambda: p.sequence[a0:a1])
    p.domain_list = [mk("KS", 0, 20, 1, 100), mk("KS", 25, 50, 120, 250), mk("AT", 60, 90, 1, 300)]
    plain = extract([p])
    fixed = extract([p], merge_split_domains=True)
    assert [r.length for r in plain.records] == [20, 25, 30]
    assert [(r.domain_id, r.ali_from, r.ali_to) for r in fixed.records] == [("KS", 0, 50), ("AT", 60, 90)]
    assert fixed.sequences[0] == p.sequence[0:50]


def test_row_block_covers_every_row_once():
    from bgclib_b200.api import row_block
    for n in (1, 7, 8, 9, 100, 28284):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                blk, r0, r1 = row_block(n, r, world)
                assert 0 <= r0 <= r1 <= n and r1 - r0 <= blk and blk * world >= n
                seen += list(range(r0, r1))
            assert seen == list(range(n))


def test_planner_persistent_database_layout():
    """n_db_first: database in front (sorted by type, length), queries appended behind it.  Every
    query x database pair of the same type exactly once, plus the self pairs (all of them with
    cross_only=1, the queries' only with cross_only=2)."""
    rng = np.random.default_rng(11)
    d_types = np.sort(rng.integers(0, 4, 90)).astype(np.int32)
    d_lens = np.concatenate([np.sort(rng.integers(40, 500, int((d_types == t).sum()))) for t in range(4)]).astype(np.int32)
    q_types = np.sort(rng.choice([0, 2, 3, 5], 17)).astype(np.int32)          # type 5 is not in the database
    q_lens = np.concatenate([np.sort(rng.integers(40, 500, int((q_types == t).sum()))) for t in np.unique(q_types)]).astype(np.int32)
    lens, types = np.concatenate([d_lens, q_lens]), np.concatenate([d_types, q_types])
    roles = np.concatenate([np.ones(90, np.int32), np.zeros(17, np.int32)])
    nd, n = 90, 107
    for same_type in (True, False):
        for cross in (1, 2):
            seen = []
            for r in range(3):
                tiles, info = plan_host(lens, types, roles, same_type_only=same_type, cross_only=cross,
                                        n_db_first=nd, rank=r, world=3)
                for t in tiles:
                    for s in range(t["s_begin"], t["s_end"]):
                        for q in (t["qa"], t["qb"]):
                            if q < 0:
                                continue
                            if q == s or (q >= nd) != (s >= nd):            # what emit_pair keeps in cross mode
                                seen.append((q, s))
            want = {(i, i) for i in range(nd if cross == 2 else 0, n)}
            want |= {(q, d) for q in range(nd, n) for d in range(nd) if (not same_type or types[q] == types[d])}
            assert len(seen) == len(set(seen)) and set(seen) == want
            assert info["total_pairs"] == len(want)
    with pytest.raises(ValueError):
        plan_host(lens, types, roles, cross_only=False, n_db_first=nd)


def test_hand_built_batches_need_unique_type_keys_and_bgc_ids():
    from bgclib_b200.api import _as_batch
    from bgclib_b200.extract import DomainBatch
    ok = DomainBatch(["ACD", "ACE"], [], ["b0", "b1"], np.array([0, 1], np.int32), ["KS", "AT"], np.array([0, 1], np.int32))
    assert _as_batch(ok, "domain", True, None, None) is ok
    for bad in (DomainBatch(["ACD", "ACE"], [], ["b0", "b1"], np.array([0, 1], np.int32), ["KS", "KS"], np.array([0, 1], np.int32)),
                DomainBatch(["ACD", "ACE"], [], ["b0", "b0"], np.array([0, 1], np.int32), ["KS", "AT"], np.array([0, 1], np.int32))):
        with pytest.raises(ValueError, match="unique"):
            _as_batch(bad, "domain", True, None, None)


def test_bench_reference_arm_line_contract():
    """`bench.py --impl reference` (the oracle port on the host cores — the one place besides the
    tests where bench.py may execute oracle/): one JSON line with the keys the driver divides by,
    the same `config` keys and units as the GPU arm's line, no GPU needed."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--n-seqs", "400", "--cpu-seconds", "0.5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "GCUPS" and line["unit"] == "GCUPS"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sequences" in cb["sample"]
    assert set(line["config"]) == {"workload", "mode", "matrix", "open", "extend"}
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_for_test", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert set(bench.bench_config("x")) == set(line["config"])     # the GPU arm prints the same keys
