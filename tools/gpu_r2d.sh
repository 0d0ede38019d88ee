#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2d_tests.log 2>&1
tail -n 8 gpurun_out/r2d_tests.log
timeout 200 python tools/profile_run.py 100 3 2>&1 | tail -n 2
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
print({k:b[k] for k in ('value','ms_per_step','gpu_launches','detail','parity')}); print(b['e2e']['ms_per_step'], b['roofline']['frac'])
PY
tail -n 5 gpurun_out/r2d_bench.err
bash tools/gpu_launches.sh r2d
