#!/bin/bash
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest7.log
python tools/bench_stats.py --n-seqs 3000 --above 1 > gpurun_out/r2_k7_v2d.jsonl 2>&1
python tools/bench_stats.py --n-seqs 3000 --above 1 --mode global >> gpurun_out/r2_k7_v2d.jsonl 2>&1
cat gpurun_out/r2_k7_v2d.jsonl
python tools/bench_lengths.py --classes 1100,2048,2600,4000,6000 > gpurun_out/r2_lengths_v3_long.jsonl 2>&1; cat gpurun_out/r2_lengths_v3_long.jsonl
python tools/bench_lengths.py --classes 420,1500,4000 --mode global > gpurun_out/r2_lengths_v3_global.jsonl 2>&1; cat gpurun_out/r2_lengths_v3_global.jsonl
