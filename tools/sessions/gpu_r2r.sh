#!/bin/bash
# round 2, last session (build = r02q with sweep variant 1 as the compiled-in default): bench line, e2e split, launch list, whole-CLI comparison
mkdir -p gpurun_out
TAG=${1:-r02r}
bash tools/gpu_bench1.sh ${TAG}
PGX_TIMING=1 E2E_ITERS=3 timeout 100 python tools/e2e_probe.py > gpurun_out/${TAG}_e2e_probe.log 2>&1
grep -E "^e2e|session|download" gpurun_out/${TAG}_e2e_probe.log | tail -n 6
bash tools/gpu_launches.sh ${TAG} > /dev/null 2>&1; head -n 12 gpurun_out/${TAG}_launches.md
timeout 100 bash tools/cli_compare.sh 8 > gpurun_out/${TAG}_cli_compare.log 2>&1; tail -n 6 gpurun_out/${TAG}_cli_compare.log
