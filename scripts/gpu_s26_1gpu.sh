#!/bin/bash
# round 2, GPU session 26 (1 GPU, the last seconds): the scale-24 twin of the default workload on ONE GPU, to set beside the
# 2-GPU end-to-end call of session 23
mkdir -p gpurun_out
timeout 70 python bench.py --scale 24 --steps 10 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/s26_uniform24_1gpu.json 2> gpurun_out/s26_uniform24_1gpu.err; echo "exit $?"
python - <<PY
import json
d = json.loads(open('gpurun_out/s26_uniform24_1gpu.json').read().strip().splitlines()[-1])
print('ms/step %.4f e2e %s' % (d['ms_per_step'], d['e2e']))
PY
