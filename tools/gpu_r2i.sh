#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2i_tests.log 2>&1
tail -n 3 gpurun_out/r2i_tests.log
timeout 200 python tools/profile_run.py 100 4 2>&1 | tail -n 2
bash tools/gpu_ncu_full.sh r2i "k_sweep_tiled:0" 2>&1 | tail -2
