#!/usr/bin/env python
"""K7 (alignment statistics without traceback) throughput on config-3 domains: GCUPS over a
random pair list, CUDA events on the launch stream (parity is tests/test_gpu_parity.py's job).

    python tools/bench_stats.py [--pairs 200000] [--mode local]
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=200_000)
    ap.add_argument("--mode", default="local")
    ap.add_argument("--n-seqs", type=int, default=10_000)
    ap.add_argument("--length", type=int, default=0, help="mean domain length (default: config 3, ~420)")
    ap.add_argument("--hinted", action="store_true", help="give the kernels the pairs' scores (bgca_pair_stats_hinted)")
    ap.add_argument("--above", type=int, default=None,
                    help="instead of a random list: score an --n-seqs collection all-vs-all and run the statistics "
                         "for every pair above this score, selected on the device (bgca_select_pairs)")
    args = ap.parse_args()
    import torch
    import bgclib_b200
    from bgclib_b200 import synth
    if args.length:        # another length class than config 3's (the kernels are instantiated per rows-per-lane class)
        sd = max(1.0, 0.035 * args.length)
        col = synth.ks_like(args.n_seqs, seed=args.length, mean=args.length, sd=sd,
                            lo=max(1, int(args.length - 4 * sd)), hi=int(args.length + 4 * sd))
    else:
        col = synth.ks_like(args.n_seqs)
    eng = bgclib_b200.get_engine(0)
    eng.set_scoring("BLOSUM62", -12, -1, args.mode)
    eng.pack_ascii(col.ascii, col.offsets)
    if args.above is not None:
        return bench_above(args, torch, eng, col)
    rng = np.random.default_rng(1)
    pi = rng.integers(0, len(col), args.pairs).astype(np.int32)
    pj = rng.integers(0, len(col), args.pairs).astype(np.int32)
    cells = int((col.lengths[pi].astype(np.int64) * col.lengths[pj]).sum())
    tpi, tpj = torch.from_numpy(pi).cuda(), torch.from_numpy(pj).cuda()
    known = None
    if args.hinted:                          # the scores the scoring kernels report for these pairs
        known = eng.pair_stats(tpi, tpj)[:, 0].contiguous()
    out = eng.pair_stats(tpi, tpj, known)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms, kms = [], []
    for _ in range(4):
        e0.record()
        out = eng.pair_stats(tpi, tpj, known)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        kms.append(eng.last_dp_ms())          # the K7 kernels alone (CUDA events inside the engine)
    t, kt = float(np.median(ms)), float(np.median(kms))
    import os
    print(json.dumps({"stage": "K7 alignment statistics over an explicit pair list",
                      "kernel": "gotoh_stats_kernel (v1)" if os.environ.get("BGCA_STATS_V1") else
                                ("gotoh_stats2_kernel<HINT>" if args.hinted and args.mode == "local" else "gotoh_stats2_kernel"),
                      "mode": args.mode, "pairs": args.pairs, "cells": cells, "kernel_ms": kt,
                      "kernel_gcups": cells / (kt * 1e-3) / 1e9, "call_ms": t,
                      "call_gcups": cells / (t * 1e-3) / 1e9, "pairs_per_s": args.pairs / (kt * 1e-3),
                      "note": "call = bgca_pair_stats incl. the host-side grouping of the list by query"}))


def bench_above(args, torch, eng, col):
    """All-vs-all scores, then statistics for every pair above a score, all on the device."""
    n = len(col)
    eng.plan()
    scores = torch.full((n, n), -2**31, dtype=torch.int32, device="cuda")
    eng.score(scores=scores)
    torch.cuda.synchronize()
    ms, kms = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        e0.record()
        pairs, out = eng.pair_stats_above(scores, args.above)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        kms.append(eng.last_dp_ms())
    L = torch.from_numpy(col.lengths.astype(np.int64)).cuda()
    cells = int((L[pairs[:, 0].long()] * L[pairs[:, 1].long()]).sum().item())
    t, kt = float(np.median(ms)), float(np.median(kms))
    print(json.dumps({"stage": "K7 for every pair above a score, selected on the device from the score matrix",
                      "kernel": "gotoh_stats2_kernel<HINT>" if args.mode == "local" else "gotoh_stats2_kernel",
                      "mode": args.mode, "n": n, "min_score": args.above, "pairs": int(pairs.shape[0]), "cells": cells,
                      "kernel_ms": kt, "kernel_gcups": cells / (kt * 1e-3) / 1e9, "call_ms": t,
                      "call_gcups": cells / (t * 1e-3) / 1e9, "pairs_per_s": int(pairs.shape[0]) / (kt * 1e-3),
                      "note": "call = select (count, scan, fill, work items) + K7, no host round trip of the list"}))


if __name__ == "__main__":
    main()
