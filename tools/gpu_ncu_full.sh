#!/bin/bash
# ncu --set full of single launches: tools/gpu_ncu_full.sh TAG "pattern:skip" ...
mkdir -p gpurun_out
TAG=$1; shift
for spec in "$@"; do
  pat=${spec%%:*}; skip=${spec##*:}
  name=$(echo $pat | tr -c 'A-Za-z0-9_\n' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -f -o gpurun_out/${TAG}_${name} python tools/profile_run.py 100 1 > gpurun_out/${TAG}_${name}.log 2>&1
  tail -n 1 gpurun_out/${TAG}_${name}.log
done
ls -la gpurun_out/*.ncu-rep
