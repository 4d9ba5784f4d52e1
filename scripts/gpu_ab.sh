#!/bin/bash
# A/B timing of several builds of libspmvb.so: $@ = library names under spmv-fpga_b200/lib
mkdir -p gpurun_out
for l in "$@"; do
  SPMVB_LIB=$PWD/spmv-fpga_b200/lib/$l timeout 120 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-gpu-build > gpurun_out/ab_$l.json 2> gpurun_out/ab_$l.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ab_$l.json'))
    print('$l: ms/step %.4f kernel_ms %.4f (min %.4f) frac %.3f'%(d['ms_per_step'],d['roofline']['kernel_ms_avg'],d['roofline']['kernel_ms_min'],d['roofline']['frac']))
except Exception as e: print('$l failed rc=$rc', e)
PY
done
