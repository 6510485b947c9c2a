#!/usr/bin/env python
"""Golden vectors for the split-domain merge (bgclib_b200.extract.merged_domain_view).

Runs ONLY in the build container (needs /root/reference).  It imports the REAL reference —
BGClib/BGClib.py and BGCtoolkit.py with the absent third-party imports stubbed — builds
seeded random proteins out of real BGCProtein/BGCDomain objects, runs the reference's own
``fix_core_split_domains`` (BGCtoolkit.py:575-687) on them, and writes the inputs and the
resulting domain lists to tests/golden/split_domains.json.  Both code paths of the reference
are exercised: proteins reached through ``BGC.CBPcontent`` and proteins of a
ProteinCollection with role "biosynthetic".
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import import_reference, REF  # noqa: E402


def main():
    ref = import_reference()
    import types
    sys.modules["lxml"].etree = sys.modules["lxml.etree"]
    sys.path.insert(0, str(REF))
    import importlib.util
    spec = importlib.util.spec_from_file_location("BGCtoolkit_ref", REF / "BGCtoolkit.py")
    tk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tk)           # guarded by __main__: only defines functions

    rng = np.random.default_rng(20260107)
    ids = ["ketoacyl-synt", "Ketoacyl-synt_C", "AMP-binding", "Condensation", "PP-binding"]
    cases = []
    bgc_col = ref.BGCCollection()
    prot_col = ref.ProteinCollection()
    proteins = []
    for c in range(400):
        p = ref.BGCProtein()
        p.identifier = f"case{c}"
        p.sequence = "A" * 3000
        n = int(rng.integers(0, 9)) if c >= 8 else c      # includes empty and single-domain lists
        doms, pos, prev_h1, prev_id = [], 0, 0, ids[0]
        for _ in range(n):
            ID = prev_id if rng.random() < 0.55 else ids[int(rng.integers(0, len(ids)))]
            a0 = pos + int(rng.integers(0, 60))
            a1 = a0 + int(rng.integers(20, 300))
            pos = a1 if rng.random() < 0.8 else a0          # occasional overlap / equal ali_from
            h0 = prev_h1 + int(rng.integers(0, 30)) if rng.random() < 0.6 else int(rng.integers(1, 200))
            h1 = h0 + int(rng.integers(1, 120))
            prev_h1, prev_id = h1, ID
            doms.append(ref.BGCDomain(p, ID, "PF0", "desc", "", a0, a1, h0, h1, 400, float(rng.integers(1, 500)), 0.0, ""))
        p.domain_list = doms
        proteins.append(p)
        before = [[d.ID, d.ali_from, d.ali_to, d.hmm_from, d.hmm_to, d.score] for d in doms]
        cases.append({"identifier": p.identifier, "via": "bgc" if c % 2 == 0 else "protein_collection",
                      "domains": before})
        if c % 2 == 0:
            bgc = ref.BGC()
            bgc.identifier = f"bgc{c}"
            bgc.protein_list.append(p)
            bgc.CBPcontent["x"] = [p]
            bgc_col.bgcs[bgc.identifier] = bgc
        else:
            p.role = "biosynthetic"
            prot_col.proteins[p.identifier] = p
    # the reference indexes domain_list[-1] unguarded in its ProteinCollection loop (:650): keep
    # empty domain lists out of that path, as its callers do (proteins come from hmmscan hits)
    for k in [k for k, p in prot_col.proteins.items() if not p.domain_list]:
        del prot_col.proteins[k]
    tk.fix_core_split_domains(bgc_col, prot_col)
    for case, p in zip(cases, proteins):
        case["after"] = [[d.ID, d.ali_from, d.ali_to, d.hmm_from, d.hmm_to, d.score] for d in p.domain_list]
    n_merged = sum(len(c["after"]) < len(c["domains"]) for c in cases)
    (HERE / "split_domains.json").write_text(json.dumps({"cases": cases}) + "\n")
    print(f"wrote split_domains.json: {len(cases)} proteins, {n_merged} with merged fragments")


if __name__ == "__main__":
    main()
