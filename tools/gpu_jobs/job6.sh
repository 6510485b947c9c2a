#!/bin/bash
# tests + K7 benches + bench line with api_e2e
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest6.log
python tools/bench_stats.py --hinted > gpurun_out/r2_k7_v2c.jsonl 2>&1
python tools/bench_stats.py >> gpurun_out/r2_k7_v2c.jsonl 2>&1
python tools/bench_stats.py --above 400 >> gpurun_out/r2_k7_v2c.jsonl 2>&1
python tools/bench_stats.py --above 400 --mode global >> gpurun_out/r2_k7_v2c.jsonl 2>&1
cat gpurun_out/r2_k7_v2c.jsonl
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --fused-steps 1 > gpurun_out/r2_bench2_n1.json 2> gpurun_out/r2_bench2_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench2_n1.json').read()); print(json.dumps({k:d[k] for k in ('value','e2e','fused','api_e2e','parity_sample')}, indent=1))"
tail -3 gpurun_out/r2_bench2_n1.err
