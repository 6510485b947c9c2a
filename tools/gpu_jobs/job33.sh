#!/bin/bash
# K7 occupancy classes: e0 = committed launch bounds, e1 = one CTA/SM less for the 8/12-row classes of the hinted and global kernels, e2 = also for the 16/20-row classes
run() { for L in 240 370 500 620; do python tools/bench_stats.py --hinted --length $L --n-seqs 6000; python tools/bench_stats.py --mode global --length $L --n-seqs 6000; done; }
cp bgclib_b200/libbgca.so /tmp/libbgca_base.so
run > gpurun_out/r2n_k7_e0.jsonl 2>&1
for e in 1 2; do cp tools/_build/libbgca_e$e.so bgclib_b200/libbgca.so; run > gpurun_out/r2n_k7_e$e.jsonl 2>&1; done
cp /tmp/libbgca_base.so bgclib_b200/libbgca.so
python3 - <<'PY'
import json
for e in (0,1,2):
    f=f"gpurun_out/r2n_k7_e{e}.jsonl"
    print(e, [(json.loads(l)["mode"][0]+("h" if "HINT" in json.loads(l)["kernel"] else ""), round(json.loads(l)["kernel_gcups"])) for l in open(f) if l.startswith("{")])
PY
