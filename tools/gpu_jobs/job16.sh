#!/bin/bash
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest12.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest12.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
( time python bench.py ) > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; echo "bench rc=$?"; tail -4 gpurun_out/r2_final_bench_n1.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err; echo "ref rc=$?"; tail -4 gpurun_out/r2_final_bench_ref.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_final_bench_n1.json') if l.startswith('{')][-1])
r=json.loads([l for l in open('gpurun_out/r2_final_bench_ref.json') if l.startswith('{')][-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['unit'], 'ref', r['value'], r['e2e']['unit'], 'cfg equal', d['config']==r['config'])
print('roofline', {k:d['roofline'][k] for k in ('frac','frac_at_observed_clock','frac_alu_pipe_nominal','frac_of_measured_dp_stream','traffic','traffic_over_algorithmic')})
print('stages', d['hbm_stages']['k1_pack']['frac'], d['hbm_stages']['k5_bgc_reduce']['frac'])
print('fused', d['fused']['gcups'], 'api', d['api_e2e']['gcups'], d['api_e2e']['extract_plus_pack_share'], 'parity', d['parity_sample']['pairs'], d['parity_sample']['mismatches'], 'launches', d['gpu_launches'])
"
