#!/bin/bash
# $1 = N GPUs: the multi-GPU group tests (one process driving N GPUs), then bench.py under torchrun (one process per GPU):
# the default workload (strong scaling of the 1 B-nnz matrix) and the power iteration (BASELINE configs[4])
N=$1; shift
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/mg${N}_topo.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_group.py -x -q -m gpu -rs ) > gpurun_out/mg${N}_pytest_group.log 2>&1; echo "group tests exit $?"; tail -8 gpurun_out/mg${N}_pytest_group.log
run() {  # name, bench args...
  name=$1; shift
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" ) > gpurun_out/mg${N}_$name.json 2> gpurun_out/mg${N}_$name.err; rc=$?
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/mg${N}_$name.json'))
    r = d.get('roofline', {})
    print('N=$N $name: nnz %s ms/step %.4f GF %.1f eff_GBs %.0f frac(rank0) %.3f aggregate_frac %.3f e2e %.1f check %s' % (
        d['config']['nnz'], d['ms_per_step'], d['value'], d['effective_gbs'], r.get('frac', 0), r.get('aggregate_frac', 0),
        d.get('e2e', {}).get('value', 0), d.get('check')))
    print('   engine', json.dumps(d.get('engine'))[:600])
except Exception as e:
    print('N=$N $name failed rc=$rc', e); print(open('gpurun_out/mg${N}_$name.err').read()[-2500:])
PY
}
run uniform --steps 20 --warmup 3
run poweriter --workload poweriter --steps 100 --warmup 5
run rmat --workload rmat --steps 20 --warmup 3
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload poweriter --scale 22 --steps 5 --warmup 3 > gpurun_out/mg${N}_nccl_debug.json 2> gpurun_out/mg${N}_nccl_debug.err; echo "nccl debug run exit $?"
grep -E "NVLS|Channel|via P2P|NET/|Broadcast|AllReduce" gpurun_out/mg${N}_nccl_debug.err | head -30 > gpurun_out/mg${N}_nccl_debug_excerpt.txt; wc -l gpurun_out/mg${N}_nccl_debug_excerpt.txt
