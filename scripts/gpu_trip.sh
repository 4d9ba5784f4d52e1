#!/bin/bash
# parity tests (bounded) + run-length / debug-mode experiments
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for rl in 0 1 2 3 5; do
 for m in 0 3; do
  SPMVB_RUN_LOG2=$rl SPMVB_DEBUG_MODE=$m timeout 120 python bench.py --steps 100 --warmup 5 --variant 2 --no-cpu-baseline > gpurun_out/exp_rl${rl}_m$m.json 2> gpurun_out/exp_rl${rl}_m$m.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/exp_rl${rl}_m$m.json'))
    print('run_log2 $rl mode $m: ms/step %.4f kernel_ms %.4f (min %.4f) frac %.3f'%(d['ms_per_step'],d['roofline']['kernel_ms_avg'],d['roofline']['kernel_ms_min'],d['roofline']['frac']))
except Exception as e: print('rl $rl mode $m failed rc=$rc', e)
PY
 done
done
