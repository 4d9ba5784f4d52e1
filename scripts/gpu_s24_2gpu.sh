#!/bin/bash
# round 2, GPU session 24 (2 GPUs, the last GPU seconds of the round): default workload at N = 2 with the multi-GPU
# end-to-end call, then the group tests for as long as the budget lasts
mkdir -p gpurun_out
timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/s24_uniform_2gpu.json 2> gpurun_out/s24_uniform_2gpu.err; echo "exit $?"
python - <<PY
import json
d = json.loads(open('gpurun_out/s24_uniform_2gpu.json').read().strip().splitlines()[-1])
print('ms/step %.4f e2e %s check %s' % (d['ms_per_step'], d['e2e'], d['check']))
PY
timeout 45 python -m pytest tests/test_gpu_group.py -x -q -m gpu > gpurun_out/s24_pytest_group.log 2>&1; echo "group tests exit $?"; tail -3 gpurun_out/s24_pytest_group.log
