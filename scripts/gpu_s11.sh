#!/bin/bash
# round 2, GPU session 11: the wide image (2^23-column blocks, x gathered from L2, coalesced y updates) and its kernel
mkdir -p gpurun_out
SPMVB_LIB=$PWD/spmv-fpga_b200/lib/libspmvb_check.so timeout 500 python scripts/sanitize_case.py > gpurun_out/s11_boundscheck.log 2>&1; rc=$?; echo "bounds-checked small cases exit $rc"; tail -3 gpurun_out/s11_boundscheck.log
if [ $rc -ne 0 ]; then echo "ABORT"; tail -30 gpurun_out/s11_boundscheck.log; exit 1; fi
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k wide ) > gpurun_out/s11_pytest_wide.log 2>&1; rc=$?; echo "wide parity tests exit $rc"; tail -4 gpurun_out/s11_pytest_wide.log
if [ $rc -ne 0 ]; then echo "ABORT"; tail -40 gpurun_out/s11_pytest_wide.log; exit 1; fi
timeout 900 python scripts/exp_options.py uniform 26 f64 "variant=9" "variant=9,wide_hints=0" "variant=9,wide_hints=1" "variant=9,wide_range_log2=22" "variant=9,wide_range_log2=22,wide_hints=0" > gpurun_out/s11_exp_uniform26.jsonl 2> gpurun_out/s11_exp_uniform26.err; echo "exp uniform26 exit $?"; grep -v "^generated\|Warning" gpurun_out/s11_exp_uniform26.err | tail -8
timeout 900 python scripts/exp_options.py rmat 24 f64 "" "variant=9,wide_hints=0" "variant=9,wide_hints=3" "variant=9,wide_range_log2=22" > gpurun_out/s11_exp_rmat24.jsonl 2> gpurun_out/s11_exp_rmat24.err; echo "exp rmat exit $?"; grep -v "^generated" gpurun_out/s11_exp_rmat24.err | tail -6
timeout 600 python scripts/exp_options.py rmat 24 f32 "" "variant=9,wide_hints=3" > gpurun_out/s11_exp_rmat24_f32.jsonl 2> gpurun_out/s11_exp_rmat24_f32.err; echo "exp rmat f32 exit $?"; grep -v "^generated" gpurun_out/s11_exp_rmat24_f32.err | tail -4
