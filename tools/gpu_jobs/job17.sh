#!/bin/bash
# ncu summaries of the remaining kernel forms: 32-bit multi-pass (replaces K4 on the length path), K7 with hints, G = 1
python tools/bench_lengths.py --classes 4000 --cells 3e11 --steps 1 > gpurun_out/plain_w32.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gotoh_pair16 -s 3 -c 3 -o gpurun_out/prof_r02_w32 \
    python tools/bench_lengths.py --classes 4000 --cells 3e11 --steps 1 > gpurun_out/ncu_w32.log 2>&1; echo "w32 rc=$?"
python tools/bench_stats.py --pairs 60000 --hinted > gpurun_out/plain_k7h.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gotoh_stats2 -s 3 -c 2 -o gpurun_out/prof_r02_k7_hint \
    python tools/bench_stats.py --pairs 60000 --hinted > gpurun_out/ncu_k7h.log 2>&1; echo "k7 rc=$?"
python tools/bench_lengths.py --classes 20 --cells 4e11 --steps 1 > gpurun_out/plain_g1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gotoh_pair16 -s 6 -c 4 -o gpurun_out/prof_r02_g1 \
    python tools/bench_lengths.py --classes 20 --cells 4e11 --steps 1 > gpurun_out/ncu_g1.log 2>&1; echo "g1 rc=$?"
python tools/bench_lengths.py --classes 1100 --cells 4e11 --steps 1 > gpurun_out/plain_mp.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gotoh_pair16 -s 4 -c 5 -o gpurun_out/prof_r02_mp \
    python tools/bench_lengths.py --classes 1100 --cells 4e11 --steps 1 > gpurun_out/ncu_mp.log 2>&1; echo "mp rc=$?"
