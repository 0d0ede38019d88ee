#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2m_tests.log 2>&1
tail -n 3 gpurun_out/r2m_tests.log
timeout 400 python tools/fuzz_gpu.py 0 400 300 cmtie > gpurun_out/r2m_fuzz_cmtie.log 2>&1
grep -E "MISMATCH" gpurun_out/r2m_fuzz_cmtie.log | cut -c1-250 | head -5; tail -n 1 gpurun_out/r2m_fuzz_cmtie.log
PGX_ROUNDS=careful timeout 200 python tools/fuzz_gpu.py 0 200 100 > gpurun_out/r2m_fuzz_careful.log 2>&1; tail -n 1 gpurun_out/r2m_fuzz_careful.log
