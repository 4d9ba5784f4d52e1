#!/bin/bash
# kernel-time experiments: debug modes of the ring kernel (not parity-valid!): $@ = list of SPMVB_DEBUG_MODE values
mkdir -p gpurun_out
for m in "$@"; do
  SPMVB_DEBUG_MODE=$m timeout 120 python bench.py --steps 100 --warmup 5 --variant 2 --no-cpu-baseline > gpurun_out/exp_mode$m.json 2> gpurun_out/exp_mode$m.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/exp_mode$m.json'))
    print('mode $m (pf %d, bits %d): ms/step %.4f kernel_ms %.4f (min %.4f)'%($m>>8,$m&255,d['ms_per_step'],d['roofline']['kernel_ms_avg'],d['roofline']['kernel_ms_min']))
except Exception as e: print('mode $m failed rc=$rc', e)
PY
done
