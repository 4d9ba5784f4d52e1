#!/bin/bash
# bench "workload,scale,cdb,variant,dtype" specs
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=, read wl sc cdb v dt <<< "$spec"
  SECONDS=0; timeout 600 python bench.py --steps 30 --warmup 3 --workload $wl --scale $sc --cols-div-blocks $cdb --variant $v --dtype $dt --no-cpu-baseline > gpurun_out/bench_$spec.json 2> gpurun_out/bench_$spec.err; rc=$?
  echo "wall ${SECONDS}s"; tail -2 gpurun_out/bench_$spec.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$spec.json'))
    print('$spec: nnz %d pairs %d ms/step %.4f kernel_ms %.4f frac %.3f GF %.1f setup %s'%(d['config']['nnz'],d['config']['pairs'],d['ms_per_step'],d['roofline']['kernel_ms_avg'],d['roofline']['frac'],d['value'],{k:round(v,2) for k,v in d['setup_s'].items()}))
except Exception as e: print('$spec failed rc=$rc', e)
PY
done
