#!/bin/bash
# Final session of a round on the committed build: GPU tests, the bench line, and the ncu launch list of one pass.
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -n 2 -p no:cacheprovider > gpurun_out/tests_f.log 2>&1
timeout 200 python bench.py > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_f.csv python tools/profile_run.py 100 1 > gpurun_out/launches_f.log 2>&1
tail -n 2 gpurun_out/tests_f.log; tail -c 400 gpurun_out/bench_f.json; tail -n 1 gpurun_out/launches_f.log
