#!/bin/bash
# diagnostic session: is the stage-2 time bimodal per process, and does it depend on the sweep variant?
mkdir -p gpurun_out
{
nvidia-smi topo -m 2>&1 | head -8; nproc; (numactl -H 2>/dev/null | head -6) || lscpu | grep -i numa
for rep in 1 2 3; do for v in top v2 v2l; do echo "== SWEEP=$v rep $rep"; PGX_SWEEP=$v PGX_TIMING=1 timeout 300 python tools/profile_run.py 100 3 2>&1 | grep -E "gen_vtx|hits " | tail -4; done; done
} 2>&1 | tee gpurun_out/diag.log
PGX_SWEEP=v2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_v2.csv python tools/profile_run.py 100 1 2>&1 | tail -1
