#!/bin/bash
# round 2, GPU session 23 (2 GPUs, short): the multi-GPU end-to-end call after the timer fix (scale-24 twin of the default workload)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --scale 24 --steps 10 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/s23_uniform24_2gpu.json 2> gpurun_out/s23_uniform24_2gpu.err; echo "exit $?"
python - <<PY
import json
d = json.loads(open('gpurun_out/s23_uniform24_2gpu.json').read().strip().splitlines()[-1])
print('ms/step %.4f e2e %s check %s' % (d['ms_per_step'], d['e2e'], d['check']))
PY
tail -5 gpurun_out/s23_uniform24_2gpu.err
