#!/usr/bin/env python
"""DESIGN.md's by-length table: GCUPS of the all-vs-all matrix path on homogeneous synthetic sets,
one JSON line per length class.  n is chosen so each class sweeps ~2e12 cells.

    python tools/bench_lengths.py [--classes 20,30,40,70,120,250,420,600,900,1100,1500,2048,3000]
"""
import argparse, json, math, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bgclib_b200
from bgclib_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--classes", default="20,30,40,70,120,250,420,600,900,1100,1500,2048,2600,4000")
ap.add_argument("--cells", type=float, default=2e12)
ap.add_argument("--mode", default="local")
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
eng = bgclib_b200.get_engine(0)
eng.set_scoring("BLOSUM62", -12, -1, a.mode)
for L in [int(x) for x in a.classes.split(",")]:
    n = int(min(40000, max(600, math.sqrt(2 * a.cells) / L)))
    sd = max(1.0, 0.035 * L)
    lo, hi = max(1, int(L - 4 * sd)), int(L + 4 * sd)
    if L in (1024, 2048):
        hi = L                      # the class that ends exactly on a pass boundary
    col = synth.ks_like(n, seed=L, mean=L, sd=sd, lo=lo, hi=hi)
    eng.pack_ascii(col.ascii, col.offsets)
    plan = eng.plan()
    scores = torch.empty((n, n), dtype=torch.int32, device="cuda")
    eng.score(scores=scores)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        eng.score(scores=scores)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"length": L, "n": n, "lo": lo, "hi": hi, "mode": a.mode, "ms": ms,
                      "gcups": plan["total_cells"] / ms / 1e6, "pairs_per_s": plan["total_pairs"] / ms * 1e3,
                      "variants": plan["n_variants"], "tiles": plan["n_tiles"],
                      "fallback32_tiles": plan["n_fallback32"]}), flush=True)
    del scores
