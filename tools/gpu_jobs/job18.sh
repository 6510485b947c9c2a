#!/bin/bash
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest13.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python tools/bench_query_db.py > gpurun_out/r2_query_db.jsonl 2>&1; tail -4 gpurun_out/r2_query_db.jsonl | cut -c1-400
