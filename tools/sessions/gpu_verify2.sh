#!/bin/bash
# Verification of the tie-order replay on the GPU: test suite, the campaign cases where ordinals of cs-tied survivors decide
# ("local" mode of tools/fuzz_gpu.py), a slice of the plain campaign, per-stage timing.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -n 2 -p no:cacheprovider > gpurun_out/tests_v2.log 2>&1
timeout 60 python tools/fuzz_gpu.py 1006,2573,176 0 50 local > gpurun_out/fuzz_gpu_ord_v2.log 2>&1
timeout 60 python tools/fuzz_gpu.py 1000 400 35 local > gpurun_out/fuzz_gpu_local_v2.log 2>&1
timeout 60 python tools/fuzz_gpu.py 400 400 35 > gpurun_out/fuzz_gpu_v2.log 2>&1
timeout 100 python tools/stage_probe.py 100 > gpurun_out/stage_v2.log 2>&1
tail -n 3 gpurun_out/tests_v2.log; tail -n 1 gpurun_out/fuzz_gpu_ord_v2.log; tail -n 1 gpurun_out/fuzz_gpu_local_v2.log; tail -n 1 gpurun_out/fuzz_gpu_v2.log; tail -n 2 gpurun_out/stage_v2.log
