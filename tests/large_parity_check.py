#!/usr/bin/env python
"""Large-sample parity run (not collected by pytest; minutes of host CPU for the oracle):

    python tests/large_parity_check.py [--n 2500]

Full score matrices of an n-domain config-3 set (local and global) and of a mixed config-4 set
against the oracle on all host cores — millions of pairs instead of the thousands the pytest
suite samples — plus the fused BGC similarity of the mixed set.  One JSON line with the counts.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2500)
    a = ap.parse_args()
    import oracle
    import bgclib_b200
    from bgclib_b200 import synth
    out = {"n": a.n}
    ks = synth.ks_like(a.n, seed=99).sequences()
    for mode, cm in (("local", oracle.MODE_LOCAL), ("global", oracle.MODE_GLOBAL)):
        got = bgclib_b200.domain_pair_scores(ks, mode=mode).scores
        t0 = time.perf_counter()
        want = oracle.all_pairs(ks, mode=cm)
        out[f"ks_{mode}_pairs"] = int(a.n * (a.n + 1) // 2)
        out[f"ks_{mode}_mismatches"] = int((got != want).sum())
        out[f"ks_{mode}_oracle_s"] = round(time.perf_counter() - t0, 1)
    mixed = synth.mibig_scale(max(8, a.n // 24), seed=98)
    seqs = mixed.sequences()
    n = len(seqs)
    want = oracle.all_pairs(seqs)
    got = bgclib_b200.domain_pair_scores(seqs).scores
    out["mixed_n"], out["mixed_pairs"] = n, int(n * (n + 1) // 2)
    out["mixed_local_mismatches"] = int((got != want).sum())
    got_g = bgclib_b200.domain_pair_scores(seqs, mode="global").scores
    out["mixed_global_mismatches"] = int((got_g != oracle.all_pairs(seqs, mode=oracle.MODE_GLOBAL)).sum())
    eng = bgclib_b200.get_engine(0)
    import torch
    eng.set_scoring()
    eng.pack(seqs, type_of_seq=mixed.type_of_seq, bgc_of_seq=mixed.bgc_of_seq, n_bgc=mixed.n_bgc)
    eng.plan(same_type_only=True)
    rb = torch.zeros((n, mixed.n_bgc), dtype=torch.int32, device="cuda")
    ss = torch.zeros(n, dtype=torch.int32, device="cuda")
    eng.score(rowbest=rb, selfsc=ss)
    sim, best, selfsum = eng.reduce_bgc(rb, ss)
    o_sim, o_best, o_self, o_rb = oracle.bgc_similarity_np(want, mixed.bgc_of_seq, mixed.type_of_seq, mixed.n_bgc)
    out["mixed_bgcs"] = int(mixed.n_bgc)
    out["similarity_mismatches"] = int((best.cpu().numpy() != o_best).sum() + (sim.cpu().numpy() != o_sim).sum()
                                       + (rb.cpu().numpy() != o_rb).sum() + (selfsum.cpu().numpy() != o_self).sum())
    out["ok"] = all(v == 0 for k, v in out.items() if k.endswith("mismatches"))
    print(json.dumps(out))
    return 0 if out["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
