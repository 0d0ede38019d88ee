#!/bin/bash
# round 2, session o: GPU suite (all failures listed) + bench line
mkdir -p gpurun_out
TAG=${1:-r2o}
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
tail -n 6 gpurun_out/${TAG}_tests.log
bash tools/gpu_bench1.sh ${TAG}
