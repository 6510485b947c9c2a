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
