#!/bin/bash
# round 2, GPU session 19 (8 GPUs): power iteration (configs[4]) with the measured kernel choice per shard, and the default
# bench (configs[3], strong scaling) as the driver will run it
N=8
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --workload poweriter --steps 100 --warmup 5 > gpurun_out/s19_poweriter_8gpu.json 2> gpurun_out/s19_poweriter_8gpu.err; rc=$?
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/s19_poweriter_8gpu.json').read().strip().splitlines()[-1])
    print('poweriter N=8: ms/iter %.4f phases %s check %s' % (d['ms_per_step'], d['engine']['last_iteration_phase_ms_max_over_ranks'], d.get('check')))
except Exception as e:
    print('poweriter failed rc=$rc', e); print(open('gpurun_out/s19_poweriter_8gpu.err').read()[-2000:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/s19_uniform_8gpu.json 2> gpurun_out/s19_uniform_8gpu.err; rc=$?
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/s19_uniform_8gpu.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('uniform N=8: ms/step %.4f GF %.1f frac(rank0) %.3f aggregate_frac %.3f e2e %.1f variant %s check %s' % (d['ms_per_step'], d['value'], r['frac'], r['aggregate_frac'], d['e2e']['value'], d['engine']['variant'], d['check']['max_err_over_tolerance']))
except Exception as e:
    print('uniform failed rc=$rc', e); print(open('gpurun_out/s19_uniform_8gpu.err').read()[-2000:])
PY
