#!/bin/bash
# The round's last GPU-box session (13 GPU-minutes left): the evidence only this side can produce comes first --
# the two ncu captures of the current build -- then the bench line, the per-stage timing, the GPU tests and smoke().
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01b.csv python tools/profile_run.py 100 1 > gpurun_out/launches_r01b.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_sweep_tiled -c 1 -o gpurun_out/sweep_r01b python tools/profile_run.py 100 1 > gpurun_out/sweep_r01b.log 2>&1
timeout 400 python bench.py > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err
PGX_TIMING=1 timeout 120 python tools/stage_probe.py 100 > gpurun_out/stage_r01b.log 2>&1
PGX_TIMING=1 E2E_ITERS=3 timeout 120 python tools/e2e_probe.py > gpurun_out/e2e_r01b.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -n 2 -p no:cacheprovider > gpurun_out/tests_r01b.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01b.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_r01b.log 2>&1
tail -3 gpurun_out/tests_r01b.log gpurun_out/smoke_r01b.log gpurun_out/bench_r01b.json
