#!/bin/bash
# round 2, GPU session 12: wide kernel with the row ids requested next to the x gathers; where its time goes
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k wide ) > gpurun_out/s12_pytest_wide.log 2>&1; rc=$?; echo "wide parity tests exit $rc"; tail -4 gpurun_out/s12_pytest_wide.log
if [ $rc -ne 0 ]; then echo "ABORT"; tail -40 gpurun_out/s12_pytest_wide.log; exit 1; fi
timeout 900 python scripts/exp_options.py uniform 26 f64 "variant=9" "variant=9,diag_flags=32" "variant=9,diag_flags=64" "variant=9,diag_flags=96" "variant=9,wide_range_log2=22" "variant=9,wide_range_log2=21" > gpurun_out/s12_exp_uniform26.jsonl 2> gpurun_out/s12_exp_uniform26.err; echo "exp uniform26 exit $?"; grep -v "^generated\|Warning" gpurun_out/s12_exp_uniform26.err | tail -8
timeout 900 python scripts/exp_options.py rmat 24 f64 "variant=9" "variant=9,diag_flags=32" "variant=9,diag_flags=64" > gpurun_out/s12_exp_rmat24.jsonl 2> gpurun_out/s12_exp_rmat24.err; echo "exp rmat exit $?"; grep -v "^generated" gpurun_out/s12_exp_rmat24.err | tail -6
timeout 600 python scripts/exp_options.py uniform 24 f64 "variant=9" > gpurun_out/s12_plain_u24.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_wide -s 6 -c 1 -f -o gpurun_out/s12_prof_uniform24_wide python scripts/exp_options.py uniform 24 f64 "variant=9" > gpurun_out/s12_ncu_u24.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/s12_plain_u24.log
ls -la gpurun_out/*.ncu-rep
