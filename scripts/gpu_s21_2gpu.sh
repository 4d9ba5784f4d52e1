#!/bin/bash
# round 2, GPU session 21 (2 GPUs): x replicated over NVLink in the multi-GPU end-to-end call; group tests; default bench
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_group.py -x -q -m gpu -rs ) > gpurun_out/s21_pytest_group.log 2>&1; echo "group tests exit $?"; tail -6 gpurun_out/s21_pytest_group.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/s21_uniform_2gpu.json 2> gpurun_out/s21_uniform_2gpu.err; rc=$?
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/s21_uniform_2gpu.json').read().strip().splitlines()[-1])
    print('uniform N=2: ms/step %.4f GF %.1f frac %.3f e2e %s check %s' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e'], d['check']))
except Exception as e:
    print('uniform failed rc=$rc', e); print(open('gpurun_out/s21_uniform_2gpu.err').read()[-3000:])
PY
