#!/bin/bash
# round 2, GPU session 10: fast path for chunks of single-entry rows
mkdir -p gpurun_out
SPMVB_LIB=$PWD/spmv-fpga_b200/lib/libspmvb_check.so timeout 400 python scripts/sanitize_case.py > gpurun_out/s10_boundscheck.log 2>&1; rc=$?; echo "bounds-checked small cases exit $rc"; tail -2 gpurun_out/s10_boundscheck.log
if [ $rc -ne 0 ]; then echo "ABORT"; tail -30 gpurun_out/s10_boundscheck.log; exit 1; fi
timeout 900 python scripts/exp_options.py uniform 26 f64 "" "tile_mb=32" "diag_flags=32" > gpurun_out/s10_exp_uniform26.jsonl 2> gpurun_out/s10_exp_uniform26.err; echo "exp uniform26 exit $?"; grep -v "^generated\|Warning\|err = " gpurun_out/s10_exp_uniform26.err | tail -6
timeout 900 python scripts/exp_options.py rmat 24 f64 "" "variant=8" > gpurun_out/s10_exp_rmat24.jsonl 2> gpurun_out/s10_exp_rmat24.err; echo "exp rmat exit $?"; grep -v "^generated" gpurun_out/s10_exp_rmat24.err | tail -4
timeout 600 python scripts/exp_options.py rmat 24 f32 "" > gpurun_out/s10_exp_rmat24_f32.jsonl 2> gpurun_out/s10_exp_rmat24_f32.err; echo "exp rmat f32 exit $?"; grep -v "^generated" gpurun_out/s10_exp_rmat24_f32.err | tail -2
timeout 600 python scripts/exp_options.py laplacian 22 f64 "" > gpurun_out/s10_exp_lap.jsonl 2> gpurun_out/s10_exp_lap.err; echo "exp lap exit $?"; grep -v "^generated" gpurun_out/s10_exp_lap.err | tail -2
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/s10_pytest_gpu.log 2>&1; echo "gpu suite exit $?"; tail -5 gpurun_out/s10_pytest_gpu.log
