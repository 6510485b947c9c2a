#!/bin/bash
# 8 GPUs: multi-rank API parity under pytest, then the bench line at N = 8 (fused leg incl. configs[4])
python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/r2_pytest_mgpu.log 2>&1; echo "pytest multi-gpu rc=$?"; tail -3 gpurun_out/r2_pytest_mgpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --fused-steps 2 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench n8 rc=$?"
tail -c 2500 gpurun_out/r2_bench_n8.json; tail -5 gpurun_out/r2_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 3 --warmup 3 --fused-steps 2 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "bench n4 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 3 --warmup 3 --fused-steps 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"
python bench.py --steps 3 --warmup 3 --fused-steps 2 > gpurun_out/r2_bench_n1_8gpubox.json 2> gpurun_out/r2_bench_n1_8gpubox.err; echo "bench n1 rc=$?"
python -c "
import json
for n in (1,2,4,8):
    f = 'gpurun_out/r2_bench_n%d.json' % n if n > 1 else 'gpurun_out/r2_bench_n1_8gpubox.json'
    d = json.loads(open(f).read())
    print(n, round(d['value']), round(d['e2e']['value']), round(d['fused']['gcups']), d['fused']['ms_per_step'], d['parity_sample']['pairs'], d['parity_sample']['mismatches'], (d.get('fused_config5') or {}).get('ms_per_step'))
"
