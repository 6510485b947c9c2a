#!/bin/bash
# final check of the round's last commit on one GPU: every GPU test, smoke(), the default bench line
python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2i_pytest.log
python __graft_entry__.py smoke > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2i_smoke.log
timeout 300 python bench.py > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; echo "bench rc=$?"
python tools/sanitizer_smoke.py > gpurun_out/r2i_sanitizer_smoke.log 2>&1; echo "kernel family smoke rc=$?"
