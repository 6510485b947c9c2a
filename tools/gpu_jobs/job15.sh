#!/bin/bash
python -m pytest tests -m gpu -q -x -k "e2e" > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest11.log
python tools/e2e_probe.py 10000
