#!/bin/bash
# bench a list of "variant[:ENV=VAL...]" specs on config 2
mkdir -p gpurun_out
for spec in "$@"; do
  v=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ':' ' '); fi
  env $envs timeout 120 python bench.py --steps 200 --warmup 5 --variant $v --no-cpu-baseline > gpurun_out/bench_$spec.json 2> gpurun_out/bench_$spec.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$spec.json'))
    print('$spec: ms/step %.4f kernel_ms %.4f (min %.4f) frac %.3f GF %.1f'%(d['ms_per_step'],d['roofline']['kernel_ms_avg'],d['roofline']['kernel_ms_min'],d['roofline']['frac'],d['value']))
except Exception as e: print('$spec failed rc=$rc', e)
PY
done
