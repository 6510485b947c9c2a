#!/usr/bin/env python
"""Where the end-to-end host call spends its time: wall clock of bgca_score_all_host against the
device time of its DP kernels, with and without the pipelined row-block path (tuning aid)."""
import os, sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bgclib_b200
from bgclib_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
col = synth.ks_like(n, seed=20260103)
eng = bgclib_b200.get_engine(0)
eng.set_scoring()
a = torch.from_numpy(col.ascii.copy()).pin_memory().numpy()
out = torch.empty((n, n), dtype=torch.int32).pin_memory().numpy()
for mode in ("pipe", "nopipe", "pipe", "nopipe"):
    if mode == "nopipe":
        os.environ["BGCA_NO_PIPELINE"] = "1"
    else:
        os.environ.pop("BGCA_NO_PIPELINE", None)
    eng.score_all_host(a, col.offsets, out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(2):
        t0 = time.perf_counter()
        eng.score_all_host(a, col.offsets, out)
        ts.append(time.perf_counter() - t0)
    print(json.dumps({"mode": mode, "n": n, "call_ms": [round(1e3 * t, 2) for t in ts], "dp_ms": round(eng.last_dp_ms(), 2),
                      "launches": eng.last_launches()}), flush=True)
