#!/bin/bash
# build of this session (trimmed hot loop, G = 1 without shuffles, multi-pass: predicated park store, 3 CTAs/SM for K <= 18):
# tests, default bench line + reference arm, by-length tables
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2e_pytest.log
timeout 300 python bench.py > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference > gpurun_out/r2e_bench_reference_arm.json 2> gpurun_out/r2e_bench_reference_arm.err; echo "reference arm rc=$?"
timeout 400 python tools/bench_lengths.py --classes 20,30,40,50,70,120,250,420,600,900,1100,1500,2048,2600,4000,6000 > gpurun_out/r2e_by_length.jsonl 2>&1
timeout 200 python tools/bench_lengths.py --classes 420,1500,4000 --mode global > gpurun_out/r2e_by_length_global.jsonl 2>&1
