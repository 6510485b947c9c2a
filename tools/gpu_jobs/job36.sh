#!/bin/bash
# multi-pass sweep instantiated per role of the pass (reader / writer of the parked row): tests, by-length of the multi-pass classes, local + global, and the 32-bit form
python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2q_pytest.log
timeout 200 python tools/bench_lengths.py --classes 420,1040,1100,1500,2048,2600,4000 > gpurun_out/r2q_by_length.jsonl 2>&1
timeout 100 python tools/bench_lengths.py --classes 1100,1500 --mode global > gpurun_out/r2q_by_length_global.jsonl 2>&1
python3 -c "
import json
for f in ('gpurun_out/r2q_by_length.jsonl','gpurun_out/r2q_by_length_global.jsonl'):
    print([(json.loads(l)['length'], round(json.loads(l)['gcups'])) for l in open(f) if l.startswith('{')])"
