#!/bin/bash
# K7 with a uniform step counter (A/B against the committed build); K4 throughput and its ncu summary
for m in "" "--hinted" "--mode global"; do python tools/bench_stats.py $m; done > gpurun_out/r2h_k7_base.jsonl 2>&1
python tools/bench_stats.py --above 600 --n-seqs 3000 >> gpurun_out/r2h_k7_base.jsonl 2>&1
python tools/bench_k4.py > gpurun_out/r2h_k4.jsonl 2>&1; python tools/bench_k4.py --mode global >> gpurun_out/r2h_k4.jsonl 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gotoh_s32 -c 2 -o gpurun_out/prof_r2h_k4 python tools/bench_k4.py --reps 1 > gpurun_out/r2h_k4_ncu.log 2>&1; echo "ncu k4 rc=$?"
cp bgclib_b200/libbgca.so /tmp/libbgca_base.so
cp tools/_build/libbgca_st2redux.so bgclib_b200/libbgca.so
for m in "" "--hinted" "--mode global"; do python tools/bench_stats.py $m; done > gpurun_out/r2h_k7_redux.jsonl 2>&1
python tools/bench_stats.py --above 600 --n-seqs 3000 >> gpurun_out/r2h_k7_redux.jsonl 2>&1
python -m pytest tests -m gpu -x -q -k "stats" > gpurun_out/r2h_pytest_redux.log 2>&1; echo "pytest stats (redux) rc=$?"; tail -1 gpurun_out/r2h_pytest_redux.log
cp /tmp/libbgca_base.so bgclib_b200/libbgca.so
cut -c1-230 gpurun_out/r2h_k7_base.jsonl gpurun_out/r2h_k7_redux.jsonl gpurun_out/r2h_k4.jsonl
