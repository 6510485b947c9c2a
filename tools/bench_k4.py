#!/usr/bin/env python
"""K4 (gotoh_s32_kernel: warp per pair, 32-bit DPX, explicit pair lists) on random pairs of the
config-3 collection: throughput with CUDA events, and the command the ncu summary under profiles/
was taken with.

    python tools/bench_k4.py [--pairs 200000] [--mode local]
"""
import argparse, json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import bgclib_b200
from bgclib_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=200_000)
ap.add_argument("--n-seqs", type=int, default=10_000)
ap.add_argument("--mode", default="local")
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
col = synth.ks_like(a.n_seqs)
eng = bgclib_b200.get_engine(0)
eng.set_scoring("BLOSUM62", -12, -1, a.mode)
eng.pack_ascii(col.ascii, col.offsets)
rng = np.random.default_rng(1)
pi = rng.integers(0, len(col), a.pairs).astype(np.int32)
pj = rng.integers(0, len(col), a.pairs).astype(np.int32)
cells = int((col.lengths[pi].astype(np.int64) * col.lengths[pj]).sum())
tpi, tpj = torch.from_numpy(pi).cuda(), torch.from_numpy(pj).cuda()
out = eng.score_pairs32(tpi, tpj)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms = []
for _ in range(a.reps):
    e0.record()
    out = eng.score_pairs32(tpi, tpj)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
t = float(np.median(ms))
print(json.dumps({"stage": "K4 gotoh_s32_kernel over an explicit pair list", "mode": a.mode, "pairs": a.pairs,
                  "cells": cells, "ms": t, "gcups": cells / (t * 1e-3) / 1e9, "pairs_per_s": a.pairs / (t * 1e-3),
                  "checksum": int(out.to(torch.int64).sum().item())}))
