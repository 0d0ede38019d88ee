#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2e_tests.log 2>&1
tail -n 5 gpurun_out/r2e_tests.log
PGX_TIMING=1 timeout 200 python tools/profile_run.py 100 2 2>&1 | grep -v "branch round" | tail -n 12
timeout 200 python tools/profile_run.py 100 4 2>&1 | tail -n 2
