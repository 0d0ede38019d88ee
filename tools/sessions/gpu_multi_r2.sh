#!/bin/bash
# multi-GPU session: tools/gpu_multi_r2.sh N TAG  -- sharded parity against the oracle (tools/dist_check.py) + the bench line with its golden gate
N=${1:-2}; TAG=${2:-r2}
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/${TAG}_dist_check_n$N.log 2>&1
echo "dist_check rc=$?"; grep -E "^case|DIST_CHECK" gpurun_out/${TAG}_dist_check_n$N.log; tail -n 3 gpurun_out/${TAG}_dist_check_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "bench rc=$?"
python - <<PY
import json
try:
    b=json.loads([l for l in open('gpurun_out/${TAG}_bench_n$N.json') if l.startswith('{')][-1])
    print({k:b[k] for k in ('n_gpus','value','ms_per_step','gpu_launches','detail','parity')}); print('e2e ms', b['e2e']['ms_per_step'])
except Exception as e: print('no json', e)
PY
grep -E "rank 0 stage 2|redoing|FATAL|Error|error" gpurun_out/${TAG}_bench_n$N.err | tail -n 8
tail -n 4 gpurun_out/${TAG}_bench_n$N.err
if [ "${3:-}" = "c5" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/dist_c5.py 1000 > gpurun_out/${TAG}_dist_c5_n$N.log 2>&1
  echo "dist_c5 rc=$?"; grep -E "^config 5|DIST_C5" gpurun_out/${TAG}_dist_c5_n$N.log; tail -n 2 gpurun_out/${TAG}_dist_c5_n$N.log
fi
