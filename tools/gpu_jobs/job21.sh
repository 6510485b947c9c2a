#!/bin/bash
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest15.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest15.log
