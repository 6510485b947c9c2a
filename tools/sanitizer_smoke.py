#!/usr/bin/env python
"""One small invocation of every kernel family, for `compute-sanitizer --tool memcheck`
(profiles/r02_sanitizer_memcheck.log).  Sizes are tiny: the tool slows kernels ~50x."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import bgclib_b200  # noqa: E402
from bgclib_b200 import synth  # noqa: E402

rng = np.random.default_rng(0)
AA = "ARNDCQEGHILKMFPSTWYV"


def rnd(n):
    return "".join(AA[k] for k in rng.integers(0, 20, n))


eng = bgclib_b200.get_engine(0)
sets = {
    "G=1/2 (1100 seqs of 18..40)": [rnd(int(k)) for k in rng.integers(18, 41, 1100)],
    "G=4..32 (60 seqs of 60..900)": [rnd(int(k)) for k in rng.integers(60, 900, 60)],
    "multi-pass (6 seqs of 1100..2300)": [rnd(int(k)) for k in rng.integers(1100, 2300, 6)],
    "32-bit multi-pass (3 seqs of 2800..3100) + K4": [rnd(2800), rnd(3000), rnd(3100), "W" * 260, "WY" * 135],
}
for name, seqs in sets.items():
    for mode in ("local", "global"):
        r = bgclib_b200.domain_pair_scores(seqs, mode=mode)
        print(name, mode, "tiles", r.plan["n_tiles"], "variants", r.plan["n_variants"], "32-bit tiles", r.plan["n_fallback32"])
col = synth.mibig_scale(20, seed=4)
objs = synth.as_collection(col)
res = bgclib_b200.bgc_similarity(objs)                      # pack_spans + fused epilogue + K5
print("bgc_similarity", res.sim.shape)
db = bgclib_b200.PackedDatabase(objs)
q = synth.as_collection(synth.mibig_scale(3, seed=5))
print("persistent database", bgclib_b200.bgc_similarity(q, database=db).sim.shape)
seqs = col.sequences()[:150]
for mode in ("local", "global"):
    st = bgclib_b200.domain_pair_stats(seqs, min_score=60 if mode == "local" else -300, mode=mode)   # select + K7 (hints)
    st2 = bgclib_b200.domain_pair_stats(seqs, st.pairs[:500], mode=mode)                            # K7, explicit list
    print("stats", mode, len(st.pairs), bool((st.stats[:500] == st2.stats).all()))
long_st = bgclib_b200.domain_pair_stats([rnd(1500), rnd(1300), rnd(200)], np.array([[0, 1], [1, 2]]))     # K7 first form
print("stats beyond 1023 residues", long_st.stats["score"].tolist())
torch.cuda.synchronize()
print("sanitizer smoke done")
