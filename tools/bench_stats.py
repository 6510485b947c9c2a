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
    args = ap.parse_args()
    import torch
    import bgclib_b200
    from bgclib_b200 import synth
    col = synth.ks_like(args.n_seqs)
    eng = bgclib_b200.get_engine(0)
    eng.set_scoring("BLOSUM62", -12, -1, args.mode)
    eng.pack_ascii(col.ascii, col.offsets)
    rng = np.random.default_rng(1)
    pi = rng.integers(0, len(col), args.pairs).astype(np.int32)
    pj = rng.integers(0, len(col), args.pairs).astype(np.int32)
    cells = int((col.lengths[pi].astype(np.int64) * col.lengths[pj]).sum())
    tpi, tpj = torch.from_numpy(pi).cuda(), torch.from_numpy(pj).cuda()
    out = eng.pair_stats(tpi, tpj)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms, kms = [], []
    for _ in range(4):
        e0.record()
        out = eng.pair_stats(tpi, tpj)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        kms.append(eng.last_dp_ms())          # the K7 kernels alone (CUDA events inside the engine)
    t, kt = float(np.median(ms)), float(np.median(kms))
    import os
    print(json.dumps({"stage": "K7 alignment statistics over an explicit pair list",
                      "kernel": "gotoh_stats_kernel (v1)" if os.environ.get("BGCA_STATS_V1") else "gotoh_stats2_kernel",
                      "mode": args.mode, "pairs": args.pairs, "cells": cells, "kernel_ms": kt,
                      "kernel_gcups": cells / (kt * 1e-3) / 1e9, "call_ms": t,
                      "call_gcups": cells / (t * 1e-3) / 1e9, "pairs_per_s": args.pairs / (kt * 1e-3),
                      "note": "call = bgca_pair_stats incl. the host-side grouping of the list by query"}))


if __name__ == "__main__":
    main()
