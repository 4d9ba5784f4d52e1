#!/bin/bash
# round 2, GPU session 4: suite + default bench with the cross-layout autotune, sanitizer, small cases for every XS config
mkdir -p gpurun_out
for cfg in 0 1 2; do
  SANITIZE_OPTS=xs_config=$cfg timeout 300 python scripts/sanitize_case.py > gpurun_out/s4_smallcases_cfg$cfg.log 2>&1; rc=$?; echo "small cases xs_config=$cfg exit $rc"; tail -1 gpurun_out/s4_smallcases_cfg$cfg.log
  if [ $rc -ne 0 ]; then echo "ABORT"; tail -30 gpurun_out/s4_smallcases_cfg$cfg.log; exit 1; fi
done
( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 ) > gpurun_out/s4_pytest_gpu.log 2>&1; echo "gpu suite exit $?"; tail -16 gpurun_out/s4_pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/s4_bench_default.json 2> gpurun_out/s4_bench_default.err; echo "bench exit $?"; tail -3 gpurun_out/s4_bench_default.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s4_bench_default.json'))
print('C4 ms %.3f GF %.1f frac %.3f e2e %.1f GF (%.2f ms) dev %s' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['engine']['device_layout']))
for k,v in d.get('also',{}).items():
    print(k, v if 'error' in v else 'ms %.4f GF %.1f frac %.3f e2e %.1f (%.3f ms) %s dev %s' % (v['ms_per_step'], v['value'], v['roofline']['frac'], v['e2e']['value'], v['e2e']['ms_per_step'], v['roofline']['kernel'], v['engine']['device_layout']))
PY
bash scripts/gpu_sanitize.sh memcheck
