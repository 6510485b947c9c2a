#!/bin/bash
# four GPUs: the N = 4 bench line of the round's last build
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 3 --warmup 3 --fused-steps 2 > gpurun_out/r2k_bench_n4.json 2> gpurun_out/r2k_bench_n4.err; echo "bench n4 rc=$?"
