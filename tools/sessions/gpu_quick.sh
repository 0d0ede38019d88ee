#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02h}
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
tail -n 2 gpurun_out/${TAG}_tests.log
bash tools/gpu_bench1.sh ${TAG}
PGX_TIMING=1 timeout 100 python tools/profile_run.py 100 3 2>&1 | grep -E "waited|by phase" | tail -n 2 | cut -c1-420
