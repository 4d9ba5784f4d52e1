#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_layout_build.py > gpurun_out/layout_build.json 2> gpurun_out/layout_build.err; echo "plain exit $?"
cat gpurun_out/layout_build.json
timeout 300 python scripts/bench_layout_build.py --workload rmat --scale 22 --cdb 16384 --cu 8 > gpurun_out/layout_build_rmat.json 2>> gpurun_out/layout_build.err; echo "rmat exit $?"
cat gpurun_out/layout_build_rmat.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/layout_launches.csv python scripts/bench_layout_build.py --reps 2 --check 0 > gpurun_out/ncu_layout.log 2>&1; echo "ncu exit $?"
