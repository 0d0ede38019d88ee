#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
b=json.loads([l for l in open('gpurun_out/${TAG}_bench.json') if l.startswith('{')][-1])
print({k:b[k] for k in ('value','ms_per_step','gpu_launches','detail','parity')}); print('e2e', b['e2e']['ms_per_step'], 'roofline', b['roofline']['frac'], b['roofline']['ms_per_launch'], 'cpu', b.get('cpu_baseline',{}).get('value'))
PY
tail -n 3 gpurun_out/${TAG}_bench.err
