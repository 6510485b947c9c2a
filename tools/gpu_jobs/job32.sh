#!/bin/bash
# K7 with hints: one CTA per SM less for the 8- and 12-row classes (no spills inside the step), A/B against the committed build
run() { for L in 240 370 420; do python tools/bench_stats.py --hinted --length $L --n-seqs 6000; done; python tools/bench_stats.py --above 1 --n-seqs 3000 --length 370; }
run > gpurun_out/r2m_k7_base.jsonl 2>&1
cp bgclib_b200/libbgca.so /tmp/libbgca_base.so; cp tools/_build/libbgca_st2hc.so bgclib_b200/libbgca.so
run > gpurun_out/r2m_k7_hintctas.jsonl 2>&1
python -m pytest tests -m gpu -x -q -k "stats" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest stats rc=$?"
cp /tmp/libbgca_base.so bgclib_b200/libbgca.so
python3 - <<'PY'
import json
for f in ("gpurun_out/r2m_k7_base.jsonl","gpurun_out/r2m_k7_hintctas.jsonl"):
    print(f, [round(json.loads(l)["kernel_gcups"]) for l in open(f) if l.startswith("{")])
PY
