#!/bin/bash
# 2-GPU session: the sharded path against the oracle (tools/dist_check.py under torchrun), then the bench line at N=2.
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dist_check.py > gpurun_out/dist_check_n2.log 2>&1
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
grep -E "case|DIST_CHECK" gpurun_out/dist_check_n2.log | tail -6
tail -c 600 gpurun_out/bench_n2.json
