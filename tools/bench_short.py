#!/usr/bin/env python
"""Throughput on one homogeneous domain type (tuning aid): --mean/--sd/--lo/--hi lengths."""
import argparse, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bgclib_b200
from bgclib_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=20000)
ap.add_argument("--mean", type=float, default=70)
ap.add_argument("--sd", type=float, default=4)
ap.add_argument("--lo", type=int, default=55)
ap.add_argument("--hi", type=int, default=90)
ap.add_argument("--mode", default="local")
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
col = synth.ks_like(a.n, seed=7, mean=a.mean, sd=a.sd, lo=a.lo, hi=a.hi)
eng = bgclib_b200.get_engine(0)
eng.set_scoring("BLOSUM62", -12, -1, a.mode)
eng.pack_ascii(col.ascii, col.offsets)
plan = eng.plan()
n = len(col)
scores = torch.empty((n, n), dtype=torch.int32, device="cuda")
for _ in range(2):
    eng.score(scores=scores)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    eng.score(scores=scores)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(json.dumps({"n": n, "mean_len": float(col.lengths.mean()), "mode": a.mode, "ms": ms,
                  "gcups": plan["total_cells"] / ms / 1e6, "pairs_per_s": plan["total_pairs"] / ms * 1e3,
                  "variants": plan["n_variants"], "tiles": plan["n_tiles"]}))
