#!/bin/bash
# ncu --set full capture of the SpMV kernel for the given bench args ($1 = output name, rest = bench args)
name=$1; shift
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/plain_$name.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_ -s 4 -c 2 -f -o gpurun_out/prof_$name python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_$name.log 2>&1
echo "ncu $name exit $?"
