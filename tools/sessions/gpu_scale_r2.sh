#!/bin/bash
# tools/gpu_scale_r2.sh N TAG [timing]: the bench line at N ranks (torchrun for N > 1); "timing" adds a second run with the device-clock phase breakdown
N=${1:-2}; TAG=${2:-r02g}
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
if [ "$N" = "1" ]; then RUN="python bench.py"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N"; fi
timeout 600 $RUN --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "bench rc=$?"
python - <<PY
import json
b=json.loads([l for l in open('gpurun_out/${TAG}_bench_n$N.json') if l.startswith('{')][-1])
print({k:b[k] for k in ('n_gpus','value','ms_per_step','gpu_launches','detail')}); print('parity', b['parity']['status'], 'e2e ms', b['e2e']['ms_per_step'], 'roofline', b['roofline']['frac'])
PY
if [ "${3:-}" = "timing" ]; then
  PGX_TIMING=1 timeout 600 $RUN --steps 3 --warmup 2 --no-parity --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_phases_n$N.err
  grep "rank 0 stage 2 by phase" gpurun_out/${TAG}_phases_n$N.err | tail -n 2 | cut -c1-700
  grep -E "rank 0.*waited|election:" gpurun_out/${TAG}_phases_n$N.err | tail -n 3
fi
