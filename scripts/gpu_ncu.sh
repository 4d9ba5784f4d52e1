#!/bin/bash
# ncu --set full of the SpMV kernel for given bench args: $1 name, rest args
name=$1; shift
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-build --e2e-steps 1 "$@" > gpurun_out/plain_$name.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_ -s 3 -c 1 -f -o gpurun_out/prof_$name python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-build --e2e-steps 1 "$@" > gpurun_out/ncu_$name.log 2>&1
echo "ncu $name exit $?"
