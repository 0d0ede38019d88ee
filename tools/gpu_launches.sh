#!/bin/bash
# launch list of one device-resident pass (ncu, per-launch durations; cold-cache and serialised: shares, not absolutes)
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_run.py 100 1 > gpurun_out/${TAG}_launches.log 2>&1
tail -n 2 gpurun_out/${TAG}_launches.log
python tools/ncu_summary.py launches gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.md 2>&1
head -n 45 gpurun_out/${TAG}_launches.md
