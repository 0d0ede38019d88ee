#!/bin/bash
# Short verification session: GPU test suite, smoke(), randomized GPU-vs-oracle campaign, CLI comparison, one bench line.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -n 2 -p no:cacheprovider > gpurun_out/tests_v.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v.log 2>&1
timeout 100 python tools/fuzz_gpu.py 177,179,225,251,463,485,502,574 0 80 > gpurun_out/fuzz_gpu_first_v.log 2>&1
timeout 120 python tools/fuzz_gpu.py 0 400 70 > gpurun_out/fuzz_gpu_v.log 2>&1
timeout 200 bash tools/cli_compare.sh 16 > gpurun_out/cli_compare_v.log 2>&1
timeout 200 python bench.py > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
tail -n 3 gpurun_out/tests_v.log; tail -n 1 gpurun_out/smoke_v.log; tail -n 1 gpurun_out/fuzz_gpu_first_v.log; tail -n 1 gpurun_out/fuzz_gpu_v.log; tail -n 6 gpurun_out/cli_compare_v.log; nproc
