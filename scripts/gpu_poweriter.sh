#!/bin/bash
# $1 = N GPUs, $2 = scale: both exchange modes, same norm?
N=$1; S=$2
mkdir -p gpurun_out
for mode in chunks broadcast; do
  SPMVB_EXCHANGE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload poweriter --dtype f32 --scale $S --steps 30 --warmup 3 > gpurun_out/pi_${N}_s${S}_$mode.json 2> gpurun_out/pi_${N}_s${S}_$mode.err; echo "$mode exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/pi_${N}_s${S}_$mode.json')); print('N=$N s$S $mode ms/iter %.4f GF %.1f last_norm %.9g bounds %s'%(d['ms_per_step'], d['value'], d['last_norm'], d['config']['row_bounds']))"
done
