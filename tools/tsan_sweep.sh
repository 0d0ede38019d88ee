#!/bin/bash
# TEST INFRASTRUCTURE: the tiled sweep kernels of pgx_kernels.cuh under ThreadSanitizer on the SIMT shim (tools/asan/sweep_race.cpp): a missing
# __syncthreads() between two shared-memory accesses of a CTA is a data race TSan reports.     usage: tools/tsan_sweep.sh [n_hits=1500] [seed=7]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
python -c "import sys; sys.path.insert(0, '$ROOT/tests'); from emu import build; print(build.build())" > /dev/null
mkdir -p /tmp/tsan && cd /tmp/tsan
g++ -std=c++17 -O1 -g -fsanitize=thread -fno-omit-frame-pointer -ffp-contract=off -pthread -I$ROOT/tests/emu -I$ROOT/tests/emu/_build -I$ROOT/include $ROOT/tools/asan/sweep_race.cpp -o sweep_race
TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4" ./sweep_race ${1:-1500} ${2:-7} 2> tsan.log | tee result.log
N=$(grep -c "WARNING: ThreadSanitizer" tsan.log || true)
echo "ThreadSanitizer reports: $N"
if [ "$N" != "0" ]; then grep -A12 "WARNING: ThreadSanitizer" tsan.log | head -60; echo "TSAN_SWEEP FAIL"; exit 1; fi
grep -q "SWEEP_RACE PASS" result.log && echo "TSAN_SWEEP PASS"
