#!/bin/bash
# $1 = N GPUs: the power iteration with exchange modes 2 and 0, with the phase times of the last iteration
N=$1
mkdir -p gpurun_out
for mode in 2 0; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$mode bench.py --gpus $N --workload poweriter --steps 100 --warmup 5 --exchange $mode > gpurun_out/pp${N}_poweriter_x$mode.json 2> gpurun_out/pp${N}_poweriter_x$mode.err; rc=$?
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/pp${N}_poweriter_x$mode.json'))
    print('N=$N exchange $mode: ms/iter %.4f phases %s' % (d['ms_per_step'], d['engine']['last_iteration_phase_ms_max_over_ranks']))
except Exception as e:
    print('N=$N exchange $mode failed rc=$rc', e); print(open('gpurun_out/pp${N}_poweriter_x$mode.err').read()[-2000:])
PY
done
