#!/bin/bash
# K7 on long per-query lists with the last build (every pair of a 3 000-domain collection, selected on the device)
python tools/bench_stats.py --above 1 --n-seqs 3000 > gpurun_out/r2l_k7_long_lists.jsonl 2>&1
python tools/bench_stats.py --above 1 --n-seqs 3000 --mode global >> gpurun_out/r2l_k7_long_lists.jsonl 2>&1
cut -c1-330 gpurun_out/r2l_k7_long_lists.jsonl
