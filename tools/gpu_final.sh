#!/bin/bash
# One GPU-box session: tests, smoke, timing in both scan modes, the bench line, and the two ncu captures kept under profiles/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 2 -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
PGX_SCAN=3pass timeout 300 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_parity.py -m gpu -q -n 2 -p no:cacheprovider 2>&1 | tail -2
for m in lookback 3pass; do echo SCAN=$m; PGX_SCAN=$m python tools/stage_probe.py 100 2>&1 | tail -2; done
PGX_TIMING=1 timeout 300 python tools/profile_run.py 100 1 2>&1 | grep "branch round" | tee gpurun_out/branch_rounds.log | tail -16
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_final.csv python tools/profile_run.py 100 1 2>&1 | tail -1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_sweep_tiled -c 1 -o gpurun_out/sweep_final python tools/profile_run.py 100 1 2>&1 | tail -1
timeout 300 bash tools/cli_compare.sh 16 2>&1 | tail -8 | tee gpurun_out/cli_compare_final.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -2
