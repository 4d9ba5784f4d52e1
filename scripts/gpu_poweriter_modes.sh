#!/bin/bash
# $1 = N GPUs: group tests (all exchange modes), then the power iteration of BASELINE configs[4] with each exchange
N=$1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_group.py -x -q -m gpu ) > gpurun_out/pm${N}_pytest_group.log 2>&1; echo "group tests exit $?"; tail -4 gpurun_out/pm${N}_pytest_group.log
for mode in 2 1 0; do
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$mode bench.py --gpus $N --workload poweriter --steps 100 --warmup 5 --exchange $mode ) > gpurun_out/pm${N}_poweriter_x$mode.json 2> gpurun_out/pm${N}_poweriter_x$mode.err; rc=$?
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/pm${N}_poweriter_x$mode.json'))
    print('N=$N exchange $mode (%s): ms/iter %.4f GF %.1f wall %.4f check %s' % (d['engine']['exchange_mode'], d['ms_per_step'], d['value'], d['engine']['wall_ms_per_step'], d['check']))
except Exception as e:
    print('N=$N exchange $mode failed rc=$rc', e); print(open('gpurun_out/pm${N}_poweriter_x$mode.err').read()[-2000:])
PY
done
