#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 2500 --csv --log-file gpurun_out/r2c_launches_warm.csv python tools/profile_run.py 100 1 > gpurun_out/r2c_launches_warm.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r2c_launches_warm.csv > gpurun_out/r2c_launches_warm.md 2>&1
head -n 30 gpurun_out/r2c_launches_warm.md
bash tools/gpu_ncu_full.sh r2c 'k_arc_insert:5' 'k_arc_fold:5' 'k_branch_hits_fast:5' 'k_seg_filter:5'
