#!/bin/bash
mkdir -p gpurun_out
SPMVB_BUILD_TRACE=1 timeout 300 python scripts/bench_layout_build.py --reps 3 > gpurun_out/layout_build_trace.json 2> gpurun_out/layout_build_trace.err; echo "trace exit $?"
grep "layout build" gpurun_out/layout_build_trace.err | tail -44
