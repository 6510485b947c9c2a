#!/usr/bin/env python
"""The reference's own example data through the engine (needs one B200):

    python examples/af293_ks_demo.py [path/to/all_KS_domains.fasta]

Without an argument the 15 KS domains of BGClib's ``examples/all_KS_domains.fasta`` are taken
from the committed fixture (tests/golden/ks_domains.json, written by the reference's own
``ProteinCollection.fasta_load``).  Prints the 15 x 15 local score matrix, the 13 x 13 BGC
similarity (two BGCs carry two KS domains each) and, for the closest domain pairs, identities
and coordinates without traceback.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bgclib_b200  # noqa: E402
from bgclib_b200 import cli  # noqa: E402


def main():
    if len(sys.argv) > 1:
        batch = cli.batch_from_fasta(sys.argv[1])
    else:
        recs = json.loads((ROOT / "tests" / "golden" / "ks_domains.json").read_text())
        fa = ROOT / "gpurun_out" / "_demo_ks.fasta"
        fa.parent.mkdir(exist_ok=True)
        fa.write_text("".join(f">{r['identifier']}\n{r['sequence']}\n" for r in recs))
        batch = cli.batch_from_fasta(fa)
    np.set_printoptions(linewidth=200)
    scores = bgclib_b200.domain_pair_scores(batch, mode="local")
    print(f"{len(batch)} domains, {scores.plan['total_pairs']} pairs, {scores.plan['total_cells']} DP cells")
    print(scores.scores)
    sim = bgclib_b200.bgc_similarity(batch)
    print("BGCs:", sim.bgc_ids)
    print(np.round(sim.sim, 3))
    close = bgclib_b200.domain_pair_stats(batch, min_score=800)
    for (i, j), st in zip(close.pairs, close.stats):
        print(f"{batch.records[i].protein_identifier} x {batch.records[j].protein_identifier}: score {st['score']}, "
              f"{st['identities']}/{st['columns']} identical, query [{st['q_start']}:{st['q_end']}], "
              f"subject [{st['s_start']}:{st['s_end']}], {st['gaps']} gap columns")


if __name__ == "__main__":
    main()
