#!/bin/bash
# G = 1 without shuffles again, multi-pass loop trimmed (one-step prefetch of the parked row, predicated park store), three CTAs/SM for multi-pass K <= 18
python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2d_pytest.log
timeout 300 python tools/bench_lengths.py --classes 20,30,40,70,600,1040,1100,1150,1500,2048,2600 > gpurun_out/r2d_by_length.jsonl 2>&1
timeout 200 python tools/bench_lengths.py --classes 420,1100,1500 --mode global > gpurun_out/r2d_by_length_global.jsonl 2>&1
timeout 300 python tests/large_parity_check.py > gpurun_out/r2d_large_parity.json 2> gpurun_out/r2d_large_parity.err; echo "large parity rc=$?"; tail -c 600 gpurun_out/r2d_large_parity.json
