"""Command-line wiring in the reference's idiom — the ``--comparative`` switch BGCtoolkit
sketches and leaves commented out (BGCtoolkit.py:157-161: "calculate protein similarity
between each pair of BGCs"), as a stand-alone entry point:

    python -m bgclib_b200.cli --comparative --fasta all_KS_domains.fasta --out results/af293
    python -m bgclib_b200.cli --comparative --fasta new.fasta --db-fasta mibig_domains.fasta --out hits
    python -m bgclib_b200.cli --comparative --bgccase AfumigatusAf293.bgccase --bgclist order.tsv \
        --alias-tsv domains_alias.tsv --out af293

``--comparative`` is spelled as the reference sketches it; it is also what the tool does when
neither it nor ``--domain-scores`` is given.  ``-l/--bgclist`` is the reference's tab-separated
list (first column = BGC identifier, ``#`` comments; read_bgc_list, BGCtoolkit.py:331-358): it
selects the BGCs compared and fixes the row order of the matrices, the way the stacked figure
the flag was meant for takes its order from that file (BGCtoolkit.py:788-801, :818).

Inputs
  --fasta      a domain FASTA as ``BGCtoolkit.py --cbt-fasta`` writes it (save_fasta,
               BGCtoolkit.py:1416-1520): header "{protein.identifier}_{dom_ID}{n} ProteinId:.. GeneId:.."
               with protein.identifier = "{BGC}~L#+CDS#".  BGC = text before "~", domain
               type = the suffix after the last "_" without its trailing counter.
  --bgccase    a pickled BGCCollection (BGClib/BGClib.py:758-777); needs the real BGClib
               importable.  --alias-tsv takes BGClib/data/domains_alias.tsv (ID <tab> alias ...).
  --db-fasta / --db-bgccase   score the input against this database instead of against itself.
Outputs (next to --out, TSV like write_metadata's tables, BGCtoolkit.py:1045-1137)
  <out>.bgc_similarity.tsv   BGC x BGC similarity, header row/column = BGC identifiers
  <out>.bgc_distance.tsv     1 - similarity (the "distance calculation" of BGClib/Readme.md:16-18)
  <out>.bgc_best.tsv         the integer numerators best[a,b]
  <out>.domain_scores.npy + <out>.domain_index.tsv    (with --domain-scores)
  <out>.domain_stats.tsv     (with --domain-stats MIN_SCORE) one line per domain pair whose score reaches
                             MIN_SCORE: identities, alignment columns, coordinates, gaps — the input an
                             identity-based distance needs (selected and computed on the device)
"""
from __future__ import annotations

import argparse
import re
import sys
from pathlib import Path
from typing import Dict, List, Tuple

import numpy as np

from .extract import DomainBatch, DomainRecord, extract


def read_fasta(path) -> List[Tuple[str, str]]:
    """Same observable behaviour as ProteinCollection.fasta_load (BGClib/BGClib.py:1569-1621):
    the whole header line is the identifier, blank lines are skipped, a repeated header is
    dropped with a warning, sequences are joined, stripped of spaces and upper-cased."""
    records: Dict[str, str] = {}
    header, chunks = None, []

    def flush():
        if header is None:
            return
        if header == "":
            print("Warning: skipping headerless sequence", file=sys.stderr)
        elif header in records:
            print(f"Warning: {header} already in collection. Skipping...", file=sys.stderr)
        else:
            records[header] = "".join(chunks).replace(" ", "").upper()

    with open(path) as f:
        for line in f:
            line = line.strip()
            if line == "":
                continue
            if line[0] == ">":
                flush()
                header, chunks = line[1:], []
            else:
                chunks.append(line)
    flush()
    return list(records.items())


_SUFFIX = re.compile(r"^(?P<prot>.*)_(?P<type>.*?)(?P<num>\d+)$")


def parse_domain_header(header: str) -> Tuple[str, str, str]:
    """"{BGC}~L0+CDS9_KS1 ProteinId:.." -> (bgc, protein identifier, domain type)."""
    first = header.split(" ")[0]
    m = _SUFFIX.match(first)
    if m:
        prot, dtype = m.group("prot"), m.group("type")
    else:
        prot, dtype = first, "domain"
    bgc = prot.split("~")[0]
    return bgc, prot, dtype or "domain"


def batch_from_fasta(path) -> DomainBatch:
    seqs, recs, bgc_ids, bgc_of, type_names, type_of = [], [], [], [], [], []
    bpos, tpos = {}, {}
    for header, seq in read_fasta(path):
        if not seq:
            continue
        bgc, prot, dtype = parse_domain_header(header)
        if bgc not in bpos:
            bpos[bgc] = len(bgc_ids)
            bgc_ids.append(bgc)
        if dtype not in tpos:
            tpos[dtype] = len(type_names)
            type_names.append(dtype)
        recs.append(DomainRecord(len(seqs), bgc, prot, header, dtype, 0, len(seq), len(seq)))
        seqs.append(seq)
        bgc_of.append(bpos[bgc])
        type_of.append(tpos[dtype])
    return DomainBatch(seqs, recs, bgc_ids, np.asarray(bgc_of, np.int32), type_names, np.asarray(type_of, np.int32))


def read_alias_tsv(path) -> Dict[str, str]:
    """BGClib/data/domains_alias.tsv: "ID <tab> alias <tab> ..." (HMM_DB.read_domain_alias_file,
    BGClib/BGClib.py:432-451); '#' lines are comments."""
    alias = {}
    with open(path) as f:
        for line in f:
            if line.startswith("#") or not line.strip():
                continue
            parts = line.rstrip("\n").split("\t")
            if len(parts) >= 2:
                alias[parts[0]] = parts[1]
    return alias


def batch_from_bgccase(path, alias, cbp_only=True) -> DomainBatch:
    try:
        from BGClib import BGCCollection
    except Exception as e:   # the reference and its third-party imports must be installed for pickles
        raise SystemExit(f"--bgccase needs the BGClib package importable: {e}")
    col = BGCCollection.from_file(Path(path))
    return extract(col, alias=alias, cbp_only=cbp_only)


def read_bgc_list(path) -> List[str]:
    """First column of the reference's --bgclist file (read_bgc_list, BGCtoolkit.py:331-358: tab
    separated, '#' lines and blank lines skipped; the optional second column names a reference
    protein for the figure and is not used here).  Order kept, repeats dropped."""
    ids: List[str] = []
    with open(path) as f:
        for line in f:
            if line[0] == "#" or line.strip() == "":
                continue
            b = line.split("\t")[0].strip()
            if b and b not in ids:
                ids.append(b)
    if not ids:
        raise SystemExit("Error: filter BGC list given but the file is empty...")
    return ids


def apply_bgc_list(batch: DomainBatch, ids: List[str]) -> DomainBatch:
    known = set(batch.bgc_ids)
    for b in ids:
        if b not in known:       # the reference warns and goes on (BGCtoolkit.py:818)
            print(f"\t--bgclist used but {b} not found in BGC data", file=sys.stderr)
    return batch.select_bgcs([b for b in ids if b in known])


def write_matrix_tsv(path, matrix, row_ids, col_ids, fmt):
    with open(path, "w") as f:
        f.write("BGC\t" + "\t".join(col_ids) + "\n")
        for rid, row in zip(row_ids, matrix):
            f.write(rid + "\t" + "\t".join(fmt(v) for v in row) + "\n")


def main(argv=None):
    ap = argparse.ArgumentParser(prog="bgclib_b200.cli", description=__doc__.split("\n\n")[0])
    ap.add_argument("--fasta")
    ap.add_argument("--bgccase")
    ap.add_argument("--db-fasta")
    ap.add_argument("--db-bgccase")
    ap.add_argument("--alias-tsv")
    ap.add_argument("--comparative", default=False, action="store_true",
                    help="calculate protein similarity between each pair of BGCs (the switch "
                         "BGCtoolkit.py:157-161 sketches); the default action of this tool")
    ap.add_argument("-l", "--bgclist", help="tab-separated file, first column = BGC identifier: the BGCs to "
                                            "compare and their order in the matrices (BGCtoolkit.py:86, :331-358)")
    ap.add_argument("--out", required=True)
    ap.add_argument("--mode", default="local", choices=["local", "global"], help="for --domain-scores")
    ap.add_argument("--open-gap-score", type=int, default=-12)
    ap.add_argument("--extend-gap-score", type=int, default=-1)
    ap.add_argument("--all-types", action="store_true", help="compare domains of different types too")
    ap.add_argument("--all-proteins", action="store_true", help="--bgccase: not only core biosynthetic proteins")
    ap.add_argument("--domain-scores", action="store_true", help="also write the domain x domain score matrix")
    ap.add_argument("--domain-stats", type=int, default=None, metavar="MIN_SCORE",
                    help="also write alignment statistics (identities, columns, coordinates) of every domain "
                         "pair whose score reaches MIN_SCORE (not with a database)")
    args = ap.parse_args(argv)
    if bool(args.fasta) == bool(args.bgccase):
        ap.error("give exactly one of --fasta / --bgccase")
    alias = read_alias_tsv(args.alias_tsv) if args.alias_tsv else None

    def load(fasta, case):
        return batch_from_fasta(fasta) if fasta else batch_from_bgccase(case, alias, not args.all_proteins)

    batch = load(args.fasta, args.bgccase)
    if args.bgclist:
        batch = apply_bgc_list(batch, read_bgc_list(args.bgclist))
        if len(batch) == 0:
            raise SystemExit("Error: no BGC of --bgclist holds domain sequences")
    db = load(args.db_fasta, args.db_bgccase) if (args.db_fasta or args.db_bgccase) else None
    from . import api   # needs the GPU from here on
    out = Path(args.out)
    out.parent.mkdir(parents=True, exist_ok=True)
    gap = dict(open_gap_score=args.open_gap_score, extend_gap_score=args.extend_gap_score)
    res = api.bgc_similarity(batch, database=db, same_type_only=not args.all_types, **gap)
    cols = res.db_bgc_ids if db is not None else res.bgc_ids
    write_matrix_tsv(f"{out}.bgc_similarity.tsv", res.sim, res.bgc_ids, cols, lambda v: repr(float(v)))
    write_matrix_tsv(f"{out}.bgc_distance.tsv", 1.0 - res.sim, res.bgc_ids, cols, lambda v: repr(float(v)))
    write_matrix_tsv(f"{out}.bgc_best.tsv", res.best, res.bgc_ids, cols, lambda v: str(int(v)))
    if args.domain_scores:
        dps = api.domain_pair_scores(batch, database=db, mode=args.mode, same_type_only=not args.all_types, **gap)
        np.save(f"{out}.domain_scores.npy", dps.scores)
        with open(f"{out}.domain_index.tsv", "w") as f:
            f.write("axis\tindex\tBGC\tprotein\ttype\tlength\n")
            for r in dps.index:
                f.write(f"row\t{r.index}\t{r.bgc_id}\t{r.protein_identifier}\t{r.type_key}\t{r.length}\n")
            for r in (dps.db_index or []):
                f.write(f"col\t{r.index}\t{r.bgc_id}\t{r.protein_identifier}\t{r.type_key}\t{r.length}\n")
    if args.domain_stats is not None:
        if db is not None:
            ap.error("--domain-stats compares the input with itself; drop the --db-* option")
        st = api.domain_pair_stats(batch, min_score=args.domain_stats, mode=args.mode,
                                   same_type_only=not args.all_types, **gap)
        recs = st.index
        with open(f"{out}.domain_stats.tsv", "w") as f:
            f.write("domain_a\tdomain_b\tBGC_a\tBGC_b\ttype\t" + "\t".join(st.stats.dtype.names) + "\tidentity\n")
            ident = st.identity
            for k, (i, j) in enumerate(st.pairs.tolist()):
                a, b = recs[i], recs[j]
                f.write(f"{a.protein_identifier}:{a.ali_from}-{a.ali_to}\t{b.protein_identifier}:{b.ali_from}-{b.ali_to}\t"
                        f"{a.bgc_id}\t{b.bgc_id}\t{a.type_key}\t"
                        + "\t".join(str(int(st.stats[name][k])) for name in st.stats.dtype.names)
                        + f"\t{ident[k]:.6f}\n")
    print(f"{len(batch)} domains in {len(res.bgc_ids)} BGCs"
          + (f" vs {len(db)} domains in {len(res.db_bgc_ids)} BGCs" if db is not None else "")
          + f": {res.plan.get('total_pairs', 0)} domain pairs aligned -> {out}.bgc_similarity.tsv")
    return 0


if __name__ == "__main__":
    sys.exit(main())
