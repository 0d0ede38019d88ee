#!/bin/bash
# diagnostic: where does bench.py's e2e leg spend its time on this box (PGX_TIMING split inside the bench process), next to the stand-alone probe
mkdir -p gpurun_out
TAG=${1:-r02s}
{ nproc; free -g | head -2; cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag; numactl --hardware 2>/dev/null | head -4; } > gpurun_out/${TAG}_host.txt 2>&1
PGX_TIMING=1 timeout 120 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
grep -E "session \+ upload|download|waited" gpurun_out/${TAG}_bench.err | tail -n 24
python -c "
import json
b=json.loads([l for l in open('gpurun_out/${TAG}_bench.json') if l.startswith('{')][-1]); print('bench e2e', b['e2e']['ms_per_step'], 'gfa', b['detail']['e2e_gfa_route']['ms_per_step'])"
PGX_TIMING=1 E2E_ITERS=3 timeout 60 python tools/e2e_probe.py 2>&1 | grep -E "^e2e|session|download" | tail -n 4
cat gpurun_out/${TAG}_host.txt
