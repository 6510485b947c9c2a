#!/usr/bin/env python
"""Regenerates the committed fixtures under tests/golden/.

Runs ONLY in the build container (needs /root/reference).  It imports the
*real* reference object model (BGClib/BGClib.py) with its three absent
third-party imports (Bio, pyhmmer, lxml — BGClib/BGClib.py:16,20-22,25-26)
stubbed in sys.modules, loads examples/all_KS_domains.fasta through the
reference's own ProteinCollection.fasta_load (BGClib/BGClib.py:1569-1621), and
writes

  ks_domains.json   the 15 KS domain records exactly as the reference parsed
                    them (identifier = whole header line, sequence = upper-cased
                    concatenation), with the BGC / CBT each came from
                    (examples/all_KS_domains.fasta.metadata.tsv)
  af293_arch.json   the 37 BGC ids (examples/AfumigatusAf293.metadata.BGCs.tsv,
                    row order) and the 57 CBP domain architectures
                    (examples/AfumigatusAf293.metadata.CBPs.tsv)
  ks_scores.json    oracle score matrices for the four parameter settings of
                    SURVEY.md App. B with their SHA-256 digests.
  ks_stats.json     oracle alignment statistics (score, identities, columns, coordinates, gaps)
                    of all 225 ordered pairs, local and global, -12/-1.

The score vectors come from oracle/ (CPU restatement), not from the reference:
the reference has no aligner (PARITY UNPINNED).
"""
import hashlib
import json
import sys
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")


def import_reference():
    for name in ("Bio", "Bio.SearchIO", "Bio.SeqIO", "Bio.SeqFeature", "pyhmmer", "pyhmmer.hmmer",
                 "pyhmmer.plan7", "pyhmmer.easel", "lxml", "lxml.etree"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["Bio"].SearchIO = sys.modules["Bio.SearchIO"]
    sys.modules["Bio"].SeqIO = sys.modules["Bio.SeqIO"]
    sys.modules["Bio.SeqFeature"].FeatureLocation = object
    sys.modules["lxml"].etree = sys.modules["lxml.etree"]
    sys.modules["pyhmmer.hmmer"].hmmpress = lambda *a, **k: None
    sys.modules["pyhmmer.plan7"].HMMFile = object
    sys.path.insert(0, str(REF))
    import BGClib  # noqa: E402
    return BGClib


def main():
    sys.path.insert(0, str(ROOT))
    import numpy as np
    import oracle

    ref = import_reference()
    pc = ref.ProteinCollection(REF / "examples" / "all_KS_domains.fasta")
    meta = {}
    with open(REF / "examples" / "all_KS_domains.fasta.metadata.tsv") as f:
        next(f)
        for line in f:
            pid, prot, dom, cbt = line.rstrip("\n").split("\t")
            meta[pid] = {"protein_id": prot, "domain_number": int(dom), "cbt": cbt}
    records = []
    for header, protein in pc.proteins.items():
        pid = header.split(" ")[0].rsplit("_KS", 1)[0]
        records.append({
            "identifier": header,
            "sequence": protein.sequence,
            "protein_identifier": pid,
            "bgc": pid.split("~")[0],
            "cbt": meta[pid]["cbt"],
            "protein_id": meta[pid]["protein_id"],
        })
    (HERE / "ks_domains.json").write_text(json.dumps(records, indent=1) + "\n")

    bgcs = []
    with open(REF / "examples" / "AfumigatusAf293.metadata.BGCs.tsv") as f:
        next(f)
        for line in f:
            bgcs.append(line.split("\t")[0])
    cbps = []
    with open(REF / "examples" / "AfumigatusAf293.metadata.CBPs.tsv") as f:
        next(f)
        for line in f:
            bgc, cbt, pid, prot, gene, arch = line.rstrip("\n").split("\t")
            arch = arch.strip()
            fwd = arch.endswith(">")
            doms = [d.strip() for d in arch.strip("<> ").split("|")]
            doms = [d for d in doms if d]
            cbps.append({"bgc": bgc, "cbt": cbt, "protein_identifier": pid, "protein_id": prot,
                         "forward": fwd, "domains": doms})
    (HERE / "af293_arch.json").write_text(json.dumps({"bgcs": bgcs, "cbps": cbps}, indent=1) + "\n")

    seqs = [r["sequence"] for r in records]
    out = {}
    for (o, e) in ((-12, -1), (-11, -1)):
        for mode_name, mode in (("local", oracle.MODE_LOCAL), ("global", oracle.MODE_GLOBAL)):
            m = oracle.all_pairs(seqs, mode=mode, open_gap=o, extend_gap=e)
            key = f"{mode_name}_{-o}_{-e}"
            out[key] = {
                "open": o, "extend": e, "mode": mode_name,
                "sha256_int32_le": hashlib.sha256(m.astype("<i4").tobytes()).hexdigest(),
                "matrix": m.tolist(),
            }
    (HERE / "ks_scores.json").write_text(json.dumps(out) + "\n")

    # alignment statistics without traceback (oracle/gotoh_oracle.c bgco_stats_pair) for every
    # ORDERED pair (query i, subject j) of the 15 KS domains, -12/-1
    n = len(seqs)
    pi, pj = np.divmod(np.arange(n * n), n)
    stats = {"fields": list(oracle.STATS_DTYPE.names), "open": -12, "extend": -1}
    for mode_name, mode in (("local", oracle.MODE_LOCAL), ("global", oracle.MODE_GLOBAL)):
        st = oracle.stats_list(seqs, pi, pj, mode=mode)
        rows = np.stack([st[k] for k in oracle.STATS_DTYPE.names], axis=1).astype("<i4")
        stats[mode_name] = {"sha256_int32_le": hashlib.sha256(rows.tobytes()).hexdigest(), "records": rows.tolist()}
    (HERE / "ks_stats.json").write_text(json.dumps(stats) + "\n")
    print("wrote", [p.name for p in HERE.glob("*.json")])
    for k, v in out.items():
        print(k, v["sha256_int32_le"], np.array(v["matrix"])[0].tolist())


if __name__ == "__main__":
    main()
