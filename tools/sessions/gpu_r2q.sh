#!/bin/bash
# round 2, session q: the sweep kernel variants (PGX_SWEEP_VARIANT 0 .. 3).  A/B timing of device-resident passes per variant first; the
# fastest one is then the variant of everything else in the session: GPU suite (which also runs every variant against the oracle), bench line,
# campaigns, ncu --set full of the first sweep.
mkdir -p gpurun_out
TAG=${1:-r02q}
for V in 0 1 2 3; do
  echo "== PGX_SWEEP_VARIANT=$V" >> gpurun_out/${TAG}_ab.log
  PGX_SWEEP_VARIANT=$V timeout 120 python tools/profile_run.py 100 6 >> gpurun_out/${TAG}_ab.log 2>&1
done
BEST=$(python tools/pick_variant.py gpurun_out/${TAG}_ab.log 2> gpurun_out/${TAG}_ab_summary.txt)
cat gpurun_out/${TAG}_ab_summary.txt
echo "chosen variant: ${BEST:-1}" | tee -a gpurun_out/${TAG}_ab_summary.txt
export PGX_SWEEP_VARIANT=${BEST:-1}
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
tail -n 3 gpurun_out/${TAG}_tests.log
bash tools/gpu_bench1.sh ${TAG}
timeout 60 python tools/fuzz_gpu.py 0 150 25 > gpurun_out/${TAG}_fuzz_gpu.log 2>&1; tail -n 1 gpurun_out/${TAG}_fuzz_gpu.log
timeout 60 python tools/fuzz_gpu.py 0 100 15 cmtie > gpurun_out/${TAG}_fuzz_cmtie_gpu.log 2>&1; tail -n 1 gpurun_out/${TAG}_fuzz_cmtie_gpu.log
bash tools/gpu_ncu_full.sh ${TAG} "k_sweep_tiled:0" 2>&1 | tail -2
