#!/bin/bash
# round 2, final single-GPU session: GPU suite, bench line, launch list, ncu --set full of the roofline kernel (+ DRAM traffic for bench.py)
mkdir -p gpurun_out
TAG=${1:-r02f}
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
tail -n 3 gpurun_out/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
bash tools/gpu_bench1.sh ${TAG}
bash tools/gpu_launches.sh ${TAG} > /dev/null 2>&1; head -n 28 gpurun_out/${TAG}_launches.md
bash tools/gpu_ncu_full.sh ${TAG} "k_sweep_tiled:0" "k_rs_scatter:0" 2>&1 | tail -3
