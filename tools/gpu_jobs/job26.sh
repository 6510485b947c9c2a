#!/bin/bash
# occupancy experiment: G = 16 local kernels as 320-thread CTAs x 2 (96 registers, 20 warps/SM) against the 256-thread build (16 warps/SM)
python tools/bench_lengths.py --classes 320,350,380,420 --steps 3 > gpurun_out/r2g_len_base.jsonl 2>&1
cp bgclib_b200/libbgca.so /tmp/libbgca_base.so
cp tools/_build/libbgca_exp320.so bgclib_b200/libbgca.so
python tools/bench_lengths.py --classes 320,350,380,420 --steps 3 > gpurun_out/r2g_len_exp320.jsonl 2>&1
python -m pytest tests -m gpu -x -q -k "parity or stress or config" > gpurun_out/r2g_pytest_exp320.log 2>&1; echo "pytest(exp320) rc=$?"; tail -2 gpurun_out/r2g_pytest_exp320.log
cp /tmp/libbgca_base.so bgclib_b200/libbgca.so
cat gpurun_out/r2g_len_base.jsonl gpurun_out/r2g_len_exp320.jsonl | cut -c1-140
