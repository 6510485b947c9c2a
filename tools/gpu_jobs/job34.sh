#!/bin/bash
# last commit of the round: every GPU test, smoke(), K7 figures of the committed build
python -m pytest tests -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2o_pytest.log
python __graft_entry__.py smoke > gpurun_out/r2o_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2o_smoke.log
for m in "" "--hinted" "--mode global"; do python tools/bench_stats.py $m; done > gpurun_out/r2o_k7.jsonl 2>&1
python tools/bench_stats.py --above 1 --n-seqs 3000 >> gpurun_out/r2o_k7.jsonl 2>&1
python3 -c "
import json
print([round(json.loads(l)['kernel_gcups']) for l in open('gpurun_out/r2o_k7.jsonl') if l.startswith('{')])"
