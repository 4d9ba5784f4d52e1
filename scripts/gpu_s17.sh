#!/bin/bash
# round 2, GPU session 17: the rest of the GPU suite after the fixed assertion; cost of the 8 row shards of the R-MAT matrix
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_layout_build.py tests/test_gpu_cg.py -x -q ) > gpurun_out/s17_pytest_rest.log 2>&1; echo "layout-build + cg tests exit $?"; tail -4 gpurun_out/s17_pytest_rest.log
timeout 900 python scripts/exp_partition.py 24 f32 8 0 4 8 16 > gpurun_out/s17_partition_f32.jsonl 2> gpurun_out/s17_partition_f32.err; echo "partition f32 exit $?"; grep "^w=" gpurun_out/s17_partition_f32.err
timeout 600 python scripts/exp_partition.py 24 f64 8 0 8 > gpurun_out/s17_partition_f64.jsonl 2> gpurun_out/s17_partition_f64.err; echo "partition f64 exit $?"; grep "^w=" gpurun_out/s17_partition_f64.err
