#!/bin/bash
# round 2, GPU session 7: the continuous x-window kernel (xs_config=3)
mkdir -p gpurun_out
SANITIZE_OPTS=xs_config=3 timeout 240 python scripts/sanitize_case.py > gpurun_out/s7_smallcases_cfg3.log 2>&1; rc=$?; echo "small cases xs_config=3 exit $rc"; tail -3 gpurun_out/s7_smallcases_cfg3.log
if [ $rc -ne 0 ]; then echo "ABORT"; tail -30 gpurun_out/s7_smallcases_cfg3.log; exit 1; fi
SPMVB_LIB=$PWD/spmv-fpga_b200/lib/libspmvb_check.so SANITIZE_OPTS=xs_config=3 timeout 600 python scripts/sanitize_case.py > gpurun_out/s7_boundscheck_cfg3.log 2>&1; echo "bounds-checked small cases xs_config=3 exit $?"; tail -2 gpurun_out/s7_boundscheck_cfg3.log
timeout 900 python scripts/exp_options.py uniform 26 f64 "xs_config=3" "xs_config=3,tile_mb=32" "xs_config=3,tile_mb=16" "xs_config=3,diag_flags=16" "xs_config=3,diag_flags=32" "diag_flags=32" > gpurun_out/s7_exp_uniform26.jsonl 2> gpurun_out/s7_exp_uniform26.err; echo "exp uniform26 exit $?"; grep -v "^generated\|Warning\|err = " gpurun_out/s7_exp_uniform26.err | tail -8
timeout 900 python scripts/exp_options.py rmat 24 f64 "xs_config=3,variant=8" "xs_config=3,variant=8,dev_tiles=1" > gpurun_out/s7_exp_rmat24.jsonl 2> gpurun_out/s7_exp_rmat24.err; echo "exp rmat exit $?"; grep -v "^generated" gpurun_out/s7_exp_rmat24.err | tail -4
timeout 600 python scripts/exp_options.py laplacian 22 f64 "xs_config=3,variant=8" > gpurun_out/s7_exp_lap.jsonl 2> gpurun_out/s7_exp_lap.err; echo "exp lap exit $?"; grep -v "^generated" gpurun_out/s7_exp_lap.err | tail -2
