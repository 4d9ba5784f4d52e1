#!/bin/bash
mkdir -p gpurun_out
if [ "$1" = "test" ]; then timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log; fi
for v in 0 8; do
  timeout 120 python bench.py --steps 200 --warmup 5 --variant $v --no-cpu-baseline > gpurun_out/bench_var$v.json 2> gpurun_out/bench_var$v.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_var$v.json'))
    print('variant $v: ms/step %.4f kernel_ms %.4f (min %.4f) frac %.3f GF %.1f e2e %.1f'%(d['ms_per_step'],d['roofline']['kernel_ms_avg'],d['roofline']['kernel_ms_min'],d['roofline']['frac'],d['value'],d['e2e']['value']))
except Exception as e: print('variant $v failed rc=$rc', e)
PY
done
bash scripts/gpu_ncu.sh xs --variant 8
