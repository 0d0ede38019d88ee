#!/bin/bash
# round 2, session n: GPU suite (incl. the token route and the recycled arrays), e2e split, bench line
mkdir -p gpurun_out
TAG=${1:-r2n}
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
tail -n 4 gpurun_out/${TAG}_tests.log
PGX_TIMING=1 E2E_ITERS=4 timeout 200 python tools/e2e_probe.py > gpurun_out/${TAG}_e2e_probe.log 2>&1
grep -E "^e2e|session|download" gpurun_out/${TAG}_e2e_probe.log | tail -n 9
bash tools/gpu_bench1.sh ${TAG}
