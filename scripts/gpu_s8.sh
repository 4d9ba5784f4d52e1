#!/bin/bash
# round 2, GPU session 8: the continuous x-window kernel without the fence in leave()
mkdir -p gpurun_out
SANITIZE_OPTS=xs_config=3 timeout 240 python scripts/sanitize_case.py > gpurun_out/s8_smallcases_cfg3.log 2>&1; rc=$?; echo "small cases xs_config=3 exit $rc"; tail -2 gpurun_out/s8_smallcases_cfg3.log
if [ $rc -ne 0 ]; then echo "ABORT"; tail -30 gpurun_out/s8_smallcases_cfg3.log; exit 1; fi
timeout 900 python scripts/exp_options.py uniform 26 f64 "xs_config=3" "xs_config=3,tile_mb=32" "xs_config=3,tile_mb=16" "xs_config=3,xs_run_log2=2" > gpurun_out/s8_exp_uniform26.jsonl 2> gpurun_out/s8_exp_uniform26.err; echo "exp uniform26 exit $?"; grep -v "^generated\|Warning\|err = " gpurun_out/s8_exp_uniform26.err | tail -8
timeout 900 python scripts/exp_options.py rmat 24 f64 "xs_config=3,variant=8" "xs_config=3,variant=8,dev_tiles=1" > gpurun_out/s8_exp_rmat24.jsonl 2> gpurun_out/s8_exp_rmat24.err; echo "exp rmat exit $?"; grep -v "^generated" gpurun_out/s8_exp_rmat24.err | tail -4
timeout 600 python scripts/exp_options.py laplacian 22 f64 "xs_config=3,variant=8" > gpurun_out/s8_exp_lap.jsonl 2> gpurun_out/s8_exp_lap.err; echo "exp lap exit $?"; grep -v "^generated" gpurun_out/s8_exp_lap.err | tail -2
