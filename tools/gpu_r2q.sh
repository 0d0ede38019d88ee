#!/bin/bash
# round 2, session q: the sweep kernel variants (PGX_SWEEP_VARIANT 0 / 1 / 2) -- GPU suite on the default, A/B timing of one device-resident
# pass per variant, campaigns on every variant, bench line, launch list, ncu --set full of the first sweep
mkdir -p gpurun_out
TAG=${1:-r02q}
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
tail -n 3 gpurun_out/${TAG}_tests.log
for V in 0 1 2; do
  echo "== PGX_SWEEP_VARIANT=$V" >> gpurun_out/${TAG}_ab.log
  PGX_SWEEP_VARIANT=$V timeout 120 python tools/profile_run.py 100 6 >> gpurun_out/${TAG}_ab.log 2>&1
done
cat gpurun_out/${TAG}_ab.log
bash tools/gpu_bench1.sh ${TAG}
PGX_SWEEP_VARIANT=0 timeout 60 python tools/fuzz_gpu.py 0 100 20 > gpurun_out/${TAG}_fuzz_v0.log 2>&1; tail -n 1 gpurun_out/${TAG}_fuzz_v0.log
PGX_SWEEP_VARIANT=1 timeout 60 python tools/fuzz_gpu.py 0 150 25 > gpurun_out/${TAG}_fuzz_v1.log 2>&1; tail -n 1 gpurun_out/${TAG}_fuzz_v1.log
PGX_SWEEP_VARIANT=2 timeout 60 python tools/fuzz_gpu.py 0 100 20 > gpurun_out/${TAG}_fuzz_v2.log 2>&1; tail -n 1 gpurun_out/${TAG}_fuzz_v2.log
bash tools/gpu_launches.sh ${TAG} > /dev/null 2>&1; head -n 12 gpurun_out/${TAG}_launches.md
bash tools/gpu_ncu_full.sh ${TAG} "k_sweep_tiled:0" 2>&1 | tail -2
