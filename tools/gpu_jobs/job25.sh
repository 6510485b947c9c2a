#!/bin/bash
# 8 GPUs, build of this session: weak-scaling headline, fused strong-scaling leg (config 4) and one step of config 5, parity sample on every rank
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --fused-steps 2 > gpurun_out/r2f_bench_n8.json 2> gpurun_out/r2f_bench_n8.err; echo "bench n8 rc=$?"
tail -c 400 gpurun_out/r2f_bench_n8.json
