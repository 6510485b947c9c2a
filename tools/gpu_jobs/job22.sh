#!/bin/bash
# round 2, second session: hot loop with the trimmed step overhead — tests, bench line, by-length, launch list, full capture
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2c_pytest.log
timeout 300 python bench.py > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; echo "bench rc=$?"
timeout 300 python tools/bench_lengths.py --classes 20,30,40,50,70,120,250,420,600,900,1100,1500,2048,2600 > gpurun_out/r2c_by_length.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-fused --no-cpu-baseline --no-probe > gpurun_out/r2c_ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gotoh_pair16 -s 8 -c 4 -o gpurun_out/prof_r2c_dp \
    python bench.py --steps 2 --warmup 1 --no-fused --no-cpu-baseline --no-probe --no-e2e > gpurun_out/r2c_ncu_full.log 2>&1
echo "full capture rc=$?"
