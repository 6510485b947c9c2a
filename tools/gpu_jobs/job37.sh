#!/bin/bash
# last commit of the round: default bench line, then the large parity run (every kernel form incl. the role-specialised multi-pass sweep)
timeout 200 python bench.py > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err; echo "bench rc=$?"
timeout 300 python tests/large_parity_check.py > gpurun_out/r2r_large_parity.json 2> gpurun_out/r2r_large_parity.err; echo "large parity rc=$?"
python3 -c "
import json
d=json.loads(open('gpurun_out/r2r_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['fused']['gcups'], d['parity_sample']['mismatches'], d['clocks']['reasons'])
d=json.loads(open('gpurun_out/r2r_large_parity.json').read().strip().splitlines()[-1])
print({k:v for k,v in d.items() if 'mismatch' in k and v} , d['ok'])"
