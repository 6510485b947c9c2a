#!/bin/bash
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest8.log
python bench.py --steps 3 --warmup 2 --no-fused --no-cpu-baseline --no-probe > gpurun_out/r2_bench3_n1.json 2> gpurun_out/r2_bench3_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench3_n1.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d['parity_sample'])"
BGCA_NO_STAGE16=1 python bench.py --steps 3 --warmup 2 --no-fused --no-cpu-baseline --no-probe > gpurun_out/r2_bench3_n1_nostage.json 2>/dev/null
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench3_n1_nostage.json') if l.startswith('{')][-1]); print('nostage', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
python bench.py --steps 1 --warmup 1 --no-fused --no-cpu-baseline --no-probe --no-e2e --no-parity-sample > gpurun_out/plain_r02b.log 2>&1 && \
ncu --set full --clock-control none -k regex:"gotoh_pair16|permute16" -s 8 -c 9 -o gpurun_out/prof_r02_dp_stage16 \
    python bench.py --steps 1 --warmup 1 --no-fused --no-cpu-baseline --no-probe --no-e2e --no-parity-sample > gpurun_out/ncu_full_r02b.log 2>&1
echo "ncu rc=$?"
