#!/bin/bash
# round 2, closing single-GPU session (build r02p: overflow groups inside k_arc_fold, occupancy beside k_arc_rank, segment count beside k_flag_vtx_fast):
# GPU suite, smoke, bench line, e2e split, launch list, per-kernel DRAM table, ncu --set full of the roofline kernel, campaigns, whole-CLI comparison
mkdir -p gpurun_out
TAG=${1:-r02p}
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
tail -n 3 gpurun_out/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
bash tools/gpu_bench1.sh ${TAG}
PGX_TIMING=1 E2E_ITERS=3 timeout 200 python tools/e2e_probe.py > gpurun_out/${TAG}_e2e_probe.log 2>&1
grep -E "^e2e|session|download" gpurun_out/${TAG}_e2e_probe.log | tail -n 6
bash tools/gpu_launches.sh ${TAG} > /dev/null 2>&1; head -n 24 gpurun_out/${TAG}_launches.md
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 130 --csv --log-file gpurun_out/${TAG}_dram_first130.csv python tools/profile_run.py 100 1 > gpurun_out/${TAG}_dram.log 2>&1
python tools/ncu_summary.py dram gpurun_out/${TAG}_dram_first130.csv > gpurun_out/${TAG}_dram_per_kernel.md 2>&1; head -n 12 gpurun_out/${TAG}_dram_per_kernel.md
bash tools/gpu_ncu_full.sh ${TAG} "k_sweep_tiled:0" 2>&1 | tail -2
timeout 120 python tools/fuzz_gpu.py 0 200 40 > gpurun_out/${TAG}_fuzz_gpu.log 2>&1; tail -n 1 gpurun_out/${TAG}_fuzz_gpu.log
timeout 120 python tools/fuzz_gpu.py 0 100 30 cmtie > gpurun_out/${TAG}_fuzz_cmtie_gpu.log 2>&1; tail -n 1 gpurun_out/${TAG}_fuzz_cmtie_gpu.log
timeout 120 python tools/fuzz_gpu.py 0 100 30 local > gpurun_out/${TAG}_fuzz_local_gpu.log 2>&1; tail -n 1 gpurun_out/${TAG}_fuzz_local_gpu.log
timeout 200 bash tools/cli_compare.sh 8 > gpurun_out/${TAG}_cli_compare.log 2>&1; tail -n 6 gpurun_out/${TAG}_cli_compare.log
