#!/bin/bash
# session C: tiled sweep as the default (full GPU suite), its timing, and the effect of host-thread placement (NUMA)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 2 -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/c_tests.log
{
for rep in 1 2; do timeout 300 python tools/profile_run.py 100 3 2>&1 | tail -1; done
REMOTE=$(cat /sys/devices/system/node/node1/cpulist 2>/dev/null); echo "remote cpus: $REMOTE"
echo "== default (bound to the GPU-local CPUs)"; PGX_TIMING=1 timeout 300 python tools/e2e_probe.py 2>&1 | grep -E "e2e|download|upload|bound" | tail -7
echo "== PGX_NUMA=0 (scheduler's choice)"; PGX_NUMA=0 PGX_TIMING=1 timeout 300 python tools/e2e_probe.py 2>&1 | grep -E "e2e|download|upload" | tail -6
if [ -n "$REMOTE" ]; then
echo "== forced to the far socket"; PGX_NUMA_CPUS=$REMOTE PGX_TIMING=1 timeout 300 python tools/e2e_probe.py 2>&1 | grep -E "e2e|download|upload|bound" | tail -7
PGX_NUMA_CPUS=$REMOTE timeout 300 python tools/profile_run.py 100 3 2>&1 | tail -1
fi
} 2>&1 | tee gpurun_out/c_numa.log
