#!/bin/bash
mkdir -p gpurun_out
for m in 1 2; do
  SPMVB_RUN_LOG2=3 SPMVB_DEBUG_MODE=$m timeout 120 python bench.py --steps 100 --warmup 5 --variant 2 --no-cpu-baseline > gpurun_out/exp_rl3_m$m.json 2> gpurun_out/exp_rl3_m$m.err; rc=$?
  python - <<PY
import json
d=json.load(open('gpurun_out/exp_rl3_m$m.json'))
print('run_log2 3 mode $m: ms/step %.4f kernel_ms %.4f (min %.4f) frac %.3f'%(d['ms_per_step'],d['roofline']['kernel_ms_avg'],d['roofline']['kernel_ms_min'],d['roofline']['frac']))
PY
done
export SPMVB_RUN_LOG2=3
bash scripts/gpu_ncu.sh ring_rl3 --variant 2
