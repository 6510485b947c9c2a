#!/usr/bin/env python
"""Small query sets against a large resident database: one-shot packing (query + database packed
together on every call) vs the persistent database (bgca_pack_append: only the queries move).

    python tools/bench_query_db.py [--db-bgcs 2500] [--query-bgcs 10]
Engine-level calls on synthetic config-4 data, wall clock per call including every host step.
"""
from __future__ import annotations

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
    ap.add_argument("--db-bgcs", type=int, default=2500)
    ap.add_argument("--query-bgcs", type=int, default=10)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch
    from bgclib_b200 import synth
    from bgclib_b200.engine import Engine
    dbc = synth.mibig_scale(a.db_bgcs)
    qc_all = synth.mibig_scale(a.query_bgcs + 5, seed=777)
    nq = int(np.searchsorted(qc_all.bgc_of_seq, a.query_bgcs))            # domains of the first query BGCs
    qc = qc_all.subset(nq)
    d_seqs, q_seqs = dbc.sequences(), qc.sequences()
    Bd, Bq, nd = dbc.n_bgc, a.query_bgcs, len(dbc)
    dev = torch.device("cuda", 0)

    def reduce_slices(eng, rowbest, selfsc, q_first):
        sim, best, selfsum = eng.reduce_bgc(rowbest, selfsc)
        return (sim[:Bq, Bq:] if q_first else sim[Bd:, :Bd]).cpu().numpy()

    eng = Engine(0)
    eng.set_scoring()

    def one_shot():
        seqs = q_seqs + d_seqs
        types = np.concatenate([qc.type_of_seq, dbc.type_of_seq]).astype(np.int32)
        bgcs = np.concatenate([qc.bgc_of_seq, dbc.bgc_of_seq + Bq]).astype(np.int32)
        eng.pack(seqs, type_of_seq=types, bgc_of_seq=bgcs, n_bgc=Bq + Bd, n_query=nq)
        eng.plan(same_type_only=True, cross_only=True)
        rb = torch.zeros((nq + nd, Bq + Bd), dtype=torch.int32, device=dev)
        ss = torch.zeros((nq + nd,), dtype=torch.int32, device=dev)
        eng.score(rowbest=rb, selfsc=ss)
        return reduce_slices(eng, rb, ss, True)

    one_shot()
    t0 = time.perf_counter()
    for _ in range(a.reps):
        ref = one_shot()
    t_one = (time.perf_counter() - t0) / a.reps

    eng2 = Engine(0)
    eng2.set_scoring()
    t0 = time.perf_counter()
    eng2.pack(d_seqs, type_of_seq=dbc.type_of_seq, bgc_of_seq=dbc.bgc_of_seq, n_bgc=Bd)
    idx = torch.arange(nd, dtype=torch.int32, device=dev)
    db_self = eng2.score_pairs32(idx, idx)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0

    def persistent():
        eng2.pack_append(q_seqs, type_of_seq=qc.type_of_seq, bgc_of_seq=qc.bgc_of_seq + Bd, n_bgc_total=Bd + Bq)
        eng2.plan(same_type_only=True, cross_only=2)
        rb = torch.zeros((nd + nq, Bd + Bq), dtype=torch.int32, device=dev)
        ss = torch.zeros((nd + nq,), dtype=torch.int32, device=dev)
        eng2.score(rowbest=rb, selfsc=ss)
        ss[:nd] = db_self
        return reduce_slices(eng2, rb, ss, False)

    persistent()
    t0 = time.perf_counter()
    for _ in range(a.reps):
        got = persistent()
    t_per = (time.perf_counter() - t0) / a.reps
    print(json.dumps({"database_domains": nd, "database_bgcs": Bd, "query_domains": nq, "query_bgcs": Bq,
                      "one_shot_ms_per_call": 1e3 * t_one, "persistent_ms_per_call": 1e3 * t_per,
                      "database_build_ms": 1e3 * t_build, "dp_ms_last_call": eng2.last_dp_ms(),
                      "identical": bool((ref == got).all())}))


if __name__ == "__main__":
    main()
