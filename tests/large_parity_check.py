#!/usr/bin/env python
"""Large-sample parity run (not collected by pytest; minutes of host CPU for the oracle):

    python tests/large_parity_check.py [--n 2500]

Full score matrices of an n-domain config-3 set (local and global) and of a mixed config-4 set
against the oracle on all host cores — millions of pairs instead of the thousands the pytest
suite samples — plus the fused BGC similarity of the mixed set, the length classes that run on
the round-2 kernel forms (one alignment per thread / per two lanes, the multi-pass sweep, its
32-bit form) and the statistics kernel on a large pair list.  One JSON line with the counts.
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
    # ---- the kernel forms of round 2, each on a set sized for seconds of oracle time ----------
    classes = [("short_20_45", 3000, 30, 8, 16, 45),          # G = 1 / G = 2
               ("short_40_130", 2000, 80, 25, 40, 130),       # G = 2 / 4 / 8
               ("long_1025_2700", 160, 1800, 500, 1025, 2700),   # multi-pass s16x2
               ("w32_2800_4200", 40, 3500, 400, 2800, 4200)]    # scores beyond int16: multi-pass s32
    for name, n_c, mean, sd, lo, hi in classes:
        seqs_c = synth.ks_like(n_c, seed=lo, mean=mean, sd=sd, lo=lo, hi=hi, family_size=10).sequences()
        for mode, cm in (("local", oracle.MODE_LOCAL), ("global", oracle.MODE_GLOBAL)):
            res = bgclib_b200.domain_pair_scores(seqs_c, mode=mode)
            out[f"{name}_{mode}_pairs"] = int(n_c * (n_c + 1) // 2)
            out[f"{name}_{mode}_mismatches"] = int((res.scores != oracle.all_pairs(seqs_c, mode=cm)).sum())
            out[f"{name}_{mode}_tiles_32bit"] = int(res.plan["n_fallback32"])
    # ---- K7 on a large pair list (plain, with score hints, selected on the device) ------------
    rng = np.random.default_rng(7)
    npairs = 60000
    pi = rng.integers(0, a.n, npairs).astype(np.int32)
    pj = rng.integers(0, a.n, npairs).astype(np.int32)
    for mode, cm in (("local", oracle.MODE_LOCAL), ("global", oracle.MODE_GLOBAL)):
        eng.set_scoring("BLOSUM62", -12, -1, mode)
        eng.pack(ks)
        want_st = oracle.stats_list(ks, pi, pj, mode=cm)
        for tag, known in (("plain", None), ("hinted", want_st["score"].astype(np.int32))):
            got_st = np.ascontiguousarray(eng.pair_stats(pi, pj, known).cpu().numpy()).view(oracle.STATS_DTYPE).reshape(-1)
            out[f"stats_{mode}_{tag}_mismatches"] = int(sum((got_st[k] != want_st[k]).sum() for k in want_st.dtype.names))
        out[f"stats_{mode}_pairs"] = npairs
    eng.set_scoring("BLOSUM62", -12, -1, "local")
    sub_n = 600
    eng.pack(ks[:sub_n])
    eng.plan()
    sc = torch.full((sub_n, sub_n), -2**31, dtype=torch.int32, device="cuda")
    eng.score(scores=sc)
    pairs_t, st_t = eng.pair_stats_above(sc, 150)
    sel = pairs_t.cpu().numpy()
    want_sel = np.argwhere(np.triu(oracle.all_pairs(ks[:sub_n]) >= 150, k=1))
    got_st = np.ascontiguousarray(st_t.cpu().numpy()).view(oracle.STATS_DTYPE).reshape(-1)
    want_st = oracle.stats_list(ks[:sub_n], want_sel[:, 0], want_sel[:, 1])
    out["stats_selected_pairs"] = int(len(want_sel))
    out["stats_selected_mismatches"] = int((sel.shape != want_sel.shape) or (sel != want_sel).sum() +
                                           sum((got_st[k] != want_st[k]).sum() for k in want_st.dtype.names))
    out["ok"] = all(v == 0 for k, v in out.items() if k.endswith("mismatches"))
    print(json.dumps(out))
    return 0 if out["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
