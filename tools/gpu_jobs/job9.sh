#!/bin/bash
# evidence: launch list + full ncu capture of the hot kernel on the bench command, large parity run
python bench.py --steps 2 --warmup 1 --no-fused --no-cpu-baseline --no-probe > gpurun_out/plain_r02.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 1 --no-fused --no-cpu-baseline --no-probe > gpurun_out/ncu_launches_r02.log 2>&1
echo "launch list rc=$?"; grep -c gotoh gpurun_out/launches_r02.csv
ncu --set full --clock-control none --import-source on -k regex:gotoh_pair16 -s 8 -c 4 -o gpurun_out/prof_r02_dp \
    python bench.py --steps 2 --warmup 1 --no-fused --no-cpu-baseline --no-probe --no-e2e > gpurun_out/ncu_full_r02.log 2>&1
echo "full capture rc=$?"
python tests/large_parity_check.py > gpurun_out/r2_large_parity.json 2> gpurun_out/r2_large_parity.err; echo "large parity rc=$?"; cat gpurun_out/r2_large_parity.json; tail -3 gpurun_out/r2_large_parity.err
