#!/bin/bash
# round 2, session B: the sync-free rounds on the GPU -- parity suite, per-round trace, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2b_tests.log 2>&1
tail -n 15 gpurun_out/r2b_tests.log
PGX_TIMING=1 timeout 200 python tools/profile_run.py 100 2 > gpurun_out/r2b_trace.log 2>&1
tail -n 22 gpurun_out/r2b_trace.log
timeout 200 python tools/profile_run.py 100 3 2>&1 | tail -n 2
PGX_NO_DIRTY=1 timeout 200 python tools/profile_run.py 100 3 2>&1 | tail -n 1
PGX_ROUNDS=careful timeout 200 python tools/profile_run.py 100 3 2>&1 | tail -n 1
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
print({k:b[k] for k in ('value','ms_per_step','gpu_launches','detail','parity')}); print(b['e2e']['ms_per_step'], b['roofline']['frac'])
PY
tail -n 5 gpurun_out/r2b_bench.err
