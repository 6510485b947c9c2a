#!/bin/bash
# memcheck over one small invocation of every kernel family
python tools/sanitizer_smoke.py > gpurun_out/sanitizer_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/sanitizer_plain.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitizer_smoke.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -8 gpurun_out/r02_sanitizer_memcheck.log
