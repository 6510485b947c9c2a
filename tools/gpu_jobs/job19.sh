#!/bin/bash
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest14.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest14.log
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-probe --fused-steps 1 > gpurun_out/r2_bench6_n1.json 2> gpurun_out/r2_bench6_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench6_n1.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench6_n1.json') if l.startswith('{')][-1]); print(json.dumps(d['hbm_stages'])); print(json.dumps(d['api_e2e']))"
