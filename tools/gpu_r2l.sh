#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2l_tests.log 2>&1
tail -n 3 gpurun_out/r2l_tests.log
timeout 300 python tools/fuzz_gpu.py 0 150 200 cmtie > gpurun_out/r2l_fuzz_cmtie.log 2>&1
grep -c " ok " gpurun_out/r2l_fuzz_cmtie.log; grep -E "MISMATCH" gpurun_out/r2l_fuzz_cmtie.log | cut -c1-200 | head -5; tail -n 1 gpurun_out/r2l_fuzz_cmtie.log
timeout 200 python tools/fuzz_gpu.py 0 300 120 > gpurun_out/r2l_fuzz.log 2>&1; tail -n 1 gpurun_out/r2l_fuzz.log
timeout 200 python tools/fuzz_gpu.py 0 300 100 x local > gpurun_out/r2l_fuzz_local.log 2>&1; tail -n 1 gpurun_out/r2l_fuzz_local.log
