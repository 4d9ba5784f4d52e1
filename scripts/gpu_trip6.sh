#!/bin/bash
# C4 experiments: "cu,occ_run_log2" specs at uniform scale 26, cdb 16384, OCC kernel, autotune off
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=, read cu rl <<< "$spec"
  SPMVB_NO_AUTOTUNE=1 SPMVB_OCC_RUN_LOG2=$rl timeout 600 python bench.py --steps 10 --warmup 3 --workload uniform --scale 26 --cols-div-blocks 16384 --variant $VARIANT --cu $cu --no-cpu-baseline --e2e-steps 1 > gpurun_out/c4_$spec.json 2> gpurun_out/c4_$spec.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/c4_$spec.json'))
    print('cu,rl=$spec: ms/step %.3f kernel_ms %.3f frac %.3f GF %.1f'%(d['ms_per_step'],d['roofline']['kernel_ms_avg'],d['roofline']['frac'],d['value']))
except Exception as e: print('$spec failed rc=$rc', e)
PY
done
