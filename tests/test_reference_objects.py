"""Boundary tests against the REAL reference classes (BGClib/BGClib.py), imported
from /root/reference with its three absent third-party imports stubbed.  Only
runs in the build container; skipped where the reference does not exist (GPU
box) — tests/test_host.py covers the same surface with stand-ins there."""
import pickle
import sys
from pathlib import Path

import pytest

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REF / "BGClib" / "BGClib.py").exists(),
                                reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
    from make_golden import import_reference
    return import_reference()


def test_fasta_load_matches_committed_fixture(ref, ks_records):
    pc = ref.ProteinCollection(REF / "examples" / "all_KS_domains.fasta")
    assert list(pc.proteins) == [r["identifier"] for r in ks_records]
    assert [p.sequence for p in pc.proteins.values()] == [r["sequence"] for r in ks_records]


def test_extract_walks_real_objects(ref, ks_records):
    from bgclib_b200 import extract
    col = ref.BGCCollection()
    alias = {"ketoacyl-synt": "KS"}
    for k, r in enumerate(ks_records[:4]):
        bgc = ref.BGC()
        bgc.identifier = r["bgc"] + f"_{k}"
        p = ref.BGCProtein()
        p.identifier = r["protein_identifier"]
        p.parent_cluster_id = bgc.identifier
        p.sequence = "mm" + r["sequence"].lower() + "kk"          # setter upper-cases (BGClib.py:1906)
        d = ref.BGCDomain(p, "ketoacyl-synt", "PF00109.27", "", "", 2, 2 + len(r["sequence"]),
                          0, 250, 253, 300.0, 1e-90, "")
        p.domain_list.append(d)
        bgc.protein_list.append(p)
        bgc.proteins[p.identifier] = p
        bgc.CBPcontent["rPKS"] = [p]
        col.bgcs[bgc.identifier] = bgc
    b = extract(col, alias=alias)
    assert b.sequences == [r["sequence"] for r in ks_records[:4]]
    assert [r.type_key for r in b.records] == ["KS"] * 4
    assert b.bgc_ids == list(col.bgcs)
    # nothing was attached: the collection still pickles and round-trips
    blob = pickle.dumps(col)
    again = pickle.loads(blob)
    assert extract(again, alias=alias).sequences == b.sequences
    assert not any(k.startswith("_bgca") for k in vars(col))


def test_sequence80_and_get_sequence_agree(ref, ks_records):
    p = ref.BGCProtein()
    p.sequence = ks_records[0]["sequence"]
    d = ref.BGCDomain(p, "x", "", "", "", 10, 200, 0, 0, 0, 0.0, 0.0, "")
    assert d.get_sequence() == p.sequence80(10, 200).replace("\n", "")
    assert p.sequence80(5, 5) == "" and p.sequence80(-1, 5) == ""


def test_attach_adds_methods_without_state(ref):
    from bgclib_b200 import attach
    cls = attach(ref.BGCCollection)
    assert callable(cls.domain_pair_scores) and callable(cls.similarity_matrix)
    col = cls()
    pickle.dumps(col)



def _real_collection(ref, ks_records, lo, hi, tag=""):
    """A BGCCollection of the REAL reference classes: one rPKS protein per KS record, the KS
    domain embedded between linkers, BGC = the record's BGC."""
    col = ref.BGCCollection()
    for r in ks_records[lo:hi]:
        bid = r["bgc"] + tag
        bgc = col.bgcs.get(bid)
        if bgc is None:
            bgc = ref.BGC()
            bgc.identifier = bid
            col.bgcs[bid] = bgc
        p = ref.BGCProtein()
        p.identifier = r["protein_identifier"] + tag + f"#{len(bgc.protein_list)}"
        p.parent_cluster_id = bid
        p.sequence = "MA" + r["sequence"] + "GK"
        p.domain_list.append(ref.BGCDomain(p, "ketoacyl-synt", "PF00109.27", "", "", 2, 2 + len(r["sequence"]),
                                           0, 250, 253, 300.0, 1e-90, ""))
        bgc.protein_list.append(p)
        bgc.proteins[p.identifier] = p
        bgc.CBPcontent.setdefault("rPKS", []).append(p)
    return col


def test_cli_reads_bgccase_pickles_of_the_real_classes(ref, ks_records, tmp_path, monkeypatch):
    """--bgccase / --db-bgccase / --bgclist / --comparative on pickles written from the real
    BGCCollection (BGClib/BGClib.py:758-777).  No GPU here: the similarity call behind the CLI is
    replaced by the oracle's statement of it, so what is checked is the loader, the BGC list
    filter and order, and the tables written."""
    import numpy as np
    import oracle
    from bgclib_b200 import api, cli, extract
    col, dbc = _real_collection(ref, ks_records, 0, 8), _real_collection(ref, ks_records, 8, 15, tag="_db")
    case, dbcase = tmp_path / "q.bgccase", tmp_path / "db.bgccase"
    for path, c in ((case, col), (dbcase, dbc)):
        with open(path, "wb") as f:
            pickle.dump(c, f)
    alias_tsv = tmp_path / "alias.tsv"
    alias_tsv.write_text("# comment\nketoacyl-synt\tKS\n")
    alias = cli.read_alias_tsv(alias_tsv)
    got = cli.batch_from_bgccase(case, alias)
    want = extract(col, alias=alias)
    assert got.sequences == want.sequences == [r["sequence"] for r in ks_records[:8]]
    assert got.bgc_ids == want.bgc_ids and [r.type_key for r in got.records] == ["KS"] * 8

    seen = {}

    def oracle_similarity(batch, database=None, same_type_only=True, **kw):
        seen["batch"], seen["db"] = batch, database
        seqs = batch.sequences + (database.sequences if database is not None else [])
        Bq = len(batch.bgc_ids)
        bgcs = np.concatenate([batch.bgc_of_seq, database.bgc_of_seq + Bq]) if database is not None else batch.bgc_of_seq
        B = Bq + (len(database.bgc_ids) if database is not None else 0)
        sim, best, selfs, _ = oracle.bgc_similarity_np(oracle.all_pairs(seqs), bgcs, np.zeros(len(seqs), int), B)
        if database is None:
            return api.BGCSimilarity(sim, best, selfs, list(batch.bgc_ids), batch.records, {"total_pairs": 0})
        return api.BGCSimilarity(sim[:Bq, Bq:], best[:Bq, Bq:], selfs[:Bq], list(batch.bgc_ids), batch.records,
                                 {"total_pairs": 0}, list(database.bgc_ids), selfs[Bq:], best[Bq:, :Bq])
    monkeypatch.setattr(api, "bgc_similarity", oracle_similarity)

    ids = list(col.bgcs)
    order = [ids[2], ids[0], "not_a_bgc", ids[1]]
    lst = tmp_path / "order.tsv"
    lst.write_text("# bgc\tprotein\tcomment\n" + "".join(f"{b}\t\tx\n" for b in order) + "\n")
    out = tmp_path / "res"
    assert cli.main(["--comparative", "--bgccase", str(case), "--alias-tsv", str(alias_tsv), "--bgclist", str(lst),
                     "--out", str(out)]) == 0
    rows = (tmp_path / "res.bgc_similarity.tsv").read_text().splitlines()
    assert rows[0].split("\t")[1:] == [ids[2], ids[0], ids[1]] == [r.split("\t")[0] for r in rows[1:]]
    assert seen["batch"].bgc_ids == [ids[2], ids[0], ids[1]]
    assert all(r.bgc_id in (ids[2], ids[0], ids[1]) for r in seen["batch"].records)
    assert float(rows[1].split("\t")[1]) == 1.0
    # query pickle against database pickle
    assert cli.main(["--comparative", "--bgccase", str(case), "--db-bgccase", str(dbcase), "--alias-tsv",
                     str(alias_tsv), "--out", str(tmp_path / "x")]) == 0
    rows = (tmp_path / "x.bgc_similarity.tsv").read_text().splitlines()
    assert rows[0].split("\t")[1:] == list(dbc.bgcs) and [r.split("\t")[0] for r in rows[1:]] == ids
    assert seen["db"].sequences == [r["sequence"] for r in ks_records[8:15]]


def test_array_form_extraction_equals_get_sequence_on_real_objects(ref, ks_records):
    """extract() never calls get_sequence(): it collects (protein, ali_from, ali_to) and evaluates
    all slices at once.  On the real classes — with coordinates Python's slice rules have to
    handle (negative, beyond the end, reversed) — the result equals BGCDomain.get_sequence()
    (BGClib/BGClib.py:3240-3241) domain by domain, and K1's input spans cut exactly those strings."""
    import numpy as np
    from bgclib_b200 import extract
    rng = np.random.default_rng(5)
    col = ref.BGCCollection()
    want, want_ids = [], []
    for k, r in enumerate(ks_records):
        bgc = ref.BGC()
        bgc.identifier = f"bgc{k}"
        col.bgcs[bgc.identifier] = bgc
        for c in range(2):
            p = ref.BGCProtein()
            p.identifier = f"bgc{k}~L0+CDS{c}"
            p.parent_cluster_id = bgc.identifier
            p.sequence = r["sequence"][: 100 + 40 * c] if c else r["sequence"]
            L = len(p.sequence)
            coords = [(0, L), (5, 60), (-30, L), (L - 10, L + 25), (50, 50), (70, 20), (-L - 5, 12), (L, L + 3)]
            coords += [tuple(sorted(rng.integers(0, L, 2).tolist())) for _ in range(4)]
            for a0, a1 in coords:
                d = ref.BGCDomain(p, "ketoacyl-synt", "PF00109.27", "", "", int(a0), int(a1), 0, 250, 253, 1.0, 1e-9, "")
                p.domain_list.append(d)
                if d.get_sequence():
                    want.append(d.get_sequence())
                    want_ids.append((bgc.identifier, p.identifier, int(a0), int(a1)))
            bgc.protein_list.append(p)
            bgc.proteins[p.identifier] = p
            bgc.CBPcontent.setdefault("rPKS", []).append(p)
    b = extract(col)
    assert b.sequences == want
    assert [(r.bgc_id, r.protein_identifier, r.ali_from, r.ali_to) for r in b.records] == want_ids
    assert [r.length for r in b.records] == [len(s) for s in want]
    blob, starts, lengths = b.spans()
    text = blob.tobytes().decode()
    assert [text[a:a + l] for a, l in zip(starts.tolist(), lengths.tolist())] == want
    assert len(blob) == sum(len(p.sequence) for g in col.bgcs.values() for p in g.protein_list)   # each protein once
    assert len(b.skipped) == sum(1 for g in col.bgcs.values() for p in g.protein_list for d in p.domain_list
                                 if not d.get_sequence())
