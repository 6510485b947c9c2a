#!/bin/bash
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-probe --fused-steps 1 --fused-config5 --no-e2e > gpurun_out/r2_bench_config5_n1.json 2> gpurun_out/r2_bench_config5_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_config5_n1.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench_config5_n1.json') if l.startswith('{')][-1]); print(json.dumps(d['fused_config5'])); print(d['fused']['ms_per_step'], d['api_e2e']['call_s'])"
