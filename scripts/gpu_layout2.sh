#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_layout_build.py -x -q -s > gpurun_out/pytest_layout_gpu.log 2>&1; echo "layout tests exit $?"
tail -12 gpurun_out/pytest_layout_gpu.log
timeout 300 python scripts/bench_layout_build.py --reps 4 > gpurun_out/layout_build.json 2> gpurun_out/layout_build.err; echo "plain exit $?"
cat gpurun_out/layout_build.json
timeout 300 python scripts/bench_layout_build.py --workload rmat --scale 22 --cdb 16384 --cu 8 > gpurun_out/layout_build_rmat.json 2>> gpurun_out/layout_build.err; echo "rmat exit $?"
cat gpurun_out/layout_build_rmat.json
