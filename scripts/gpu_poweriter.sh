#!/bin/bash
# $1 = N GPUs.  Power iteration: both exchange modes at a small scale (same norm?), then BASELINE config 5.
N=$1
mkdir -p gpurun_out
for mode in chunks broadcast; do
  SPMVB_EXCHANGE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload poweriter --dtype f32 --scale 20 --steps 20 --warmup 3 > gpurun_out/pi_${N}_s20_$mode.json 2> gpurun_out/pi_${N}_s20_$mode.err; echo "s20 $mode exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/pi_${N}_s20_$mode.json')); print('$mode s20 ms/iter %.4f last_norm %.9g'%(d['ms_per_step'], d['config'].get('last_norm', d.get('last_norm', float('nan')))))"
done
for mode in ${MODES24:-chunks}; do
  SPMVB_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload poweriter --dtype f32 --scale 24 --steps 30 --warmup 3 > gpurun_out/pi_${N}_s24_$mode.json 2> gpurun_out/pi_${N}_s24_$mode.err; echo "s24 $mode exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/pi_${N}_s24_$mode.json')); print('$mode s24 ms/iter %.4f GF %.1f last_norm %.9g'%(d['ms_per_step'], d['value'], d['config'].get('last_norm', d.get('last_norm', float('nan')))))"
done
