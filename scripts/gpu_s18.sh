#!/bin/bash
# round 2, GPU session 18: kernels timed at engine creation for irregular matrices without a device layout (fp32 shards)
mkdir -p gpurun_out
timeout 900 python scripts/exp_partition.py 24 f32 8 0 > gpurun_out/s18_partition_f32.jsonl 2> gpurun_out/s18_partition_f32.err; echo "partition f32 exit $?"; grep "^w=" gpurun_out/s18_partition_f32.err
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "private or cu_major or power or upload or small_column" ) > gpurun_out/s18_pytest.log 2>&1; echo "parity subset exit $?"; tail -3 gpurun_out/s18_pytest.log
timeout 600 python bench.py --workload poweriter --steps 100 > gpurun_out/s18_bench_poweriter_1gpu.json 2> gpurun_out/s18_bench_poweriter_1gpu.err; echo "poweriter 1 GPU exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/s18_bench_poweriter_1gpu.json').read().strip().splitlines()[-1]); print('poweriter 1 GPU ms/iter', d['ms_per_step'])"
