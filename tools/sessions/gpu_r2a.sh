#!/bin/bash
# round 2, session A: parity gate at the headline config, per-round trace of stage 2 (design input), one bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2a_tests.log 2>&1
tail -n 3 gpurun_out/r2a_tests.log
PGX_TIMING=1 timeout 200 python tools/profile_run.py 100 2 > gpurun_out/r2a_trace.log 2>&1
tail -n 45 gpurun_out/r2a_trace.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 1500 gpurun_out/r2a_bench.json; tail -n 5 gpurun_out/r2a_bench.err
nproc; free -g | head -2
