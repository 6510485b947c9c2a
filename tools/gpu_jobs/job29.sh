#!/bin/bash
# two GPUs: the multi-rank parity check collected by pytest (torchrun, 2 ranks, NCCL) on the round's last build, and the N = 2 bench line
python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/r2j_pytest_mgpu.log 2>&1; echo "pytest multi-gpu rc=$?"; tail -3 gpurun_out/r2j_pytest_mgpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 3 --warmup 3 --fused-steps 2 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err; echo "bench n2 rc=$?"
