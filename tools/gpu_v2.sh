#!/bin/bash
# GPU session: validate the v2 sweep (PGX_SWEEP=v2) against the whole GPU suite, time it against the current kernel, whole-CLI comparison
set -x
mkdir -p gpurun_out
PGX_SWEEP=v2 timeout 900 python -m pytest tests -m gpu -q -n 2 -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/v2_tests.log
for v in top v2; do echo SWEEP=$v; PGX_SWEEP=$v timeout 300 python tools/profile_run.py 100 3 2>&1 | tail -1; done | tee gpurun_out/v2_timing.log
timeout 600 bash tools/cli_compare.sh 16 2>&1 | tail -12 | tee gpurun_out/cli_compare.log
PGX_SWEEP=v2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_sweep_v2 -c 1 -o gpurun_out/sweep_v2 python tools/profile_run.py 100 1 2>&1 | tail -1
