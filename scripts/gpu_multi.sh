#!/bin/bash
# $1 = N GPUs; rest: bench args.  Runs under torchrun and prints a one-line summary.
N=$1; shift
mkdir -p gpurun_out
tag=$(echo "$@" | tr ' /' '__')
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/mg_${N}_$tag.json 2> gpurun_out/mg_${N}_$tag.err; rc=$?
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/mg_${N}_$tag.json'))
    r=d.get('roofline',{})
    print('N=$N $@: nnz %d ms/step %.4f GF %.1f eff_GBs %.0f kernel_ms %s frac %s e2e %s variant %s'%(d['config']['nnz'],d['ms_per_step'],d['value'],d['effective_gbs'],r.get('kernel_ms_avg'),r.get('frac'),d.get('e2e',{}).get('value'),d['config'].get('variant')))
except Exception as e:
    print('N=$N $@ failed rc=$rc', e); import subprocess; print(open('gpurun_out/mg_${N}_$tag.err').read()[-1500:])
PY
