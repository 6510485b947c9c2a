#!/bin/bash
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest9.log
python bench.py --steps 3 --warmup 2 --no-fused --no-cpu-baseline --no-probe > gpurun_out/r2_bench4_n1.json 2> gpurun_out/r2_bench4_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench4_n1.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench4_n1.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d['roofline']['traffic'], d['roofline']['traffic_over_algorithmic'])"
BGCA_NO_PIPELINE=1 python bench.py --steps 3 --warmup 2 --no-fused --no-cpu-baseline --no-probe > gpurun_out/r2_bench4_n1_nopipe.json 2>/dev/null
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench4_n1_nopipe.json') if l.startswith('{')][-1]); print('nopipe', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
