#!/bin/bash
# round 2, GPU session 2: random-access ceilings, tile size / warps A/B on the uniform matrix, the suite again
mkdir -p gpurun_out
spmv-fpga_b200/lib/access_probe > gpurun_out/s2_access_probe.txt 2>&1; echo "probe exit $?"; cat gpurun_out/s2_access_probe.txt
timeout 300 python scripts/sanitize_case.py > gpurun_out/s2_smallcases.log 2>&1; rc=$?; echo "small cases exit $rc"; tail -2 gpurun_out/s2_smallcases.log
if [ $rc -ne 0 ]; then echo "ABORT: small cases failed"; exit 1; fi
timeout 900 python scripts/exp_options.py uniform 26 f64 "" "tile_mb=48" "tile_mb=64" "tile_mb=24" > gpurun_out/s2_exp_uniform26.jsonl 2> gpurun_out/s2_exp_uniform26.err; echo "exp uniform26 exit $?"; grep -v "^generated" gpurun_out/s2_exp_uniform26.err | tail -8
timeout 900 python scripts/exp_options.py rmat 24 f64 "" "variant=7" "dev_tiles=1" "dev_tiles=1,variant=7" "dev_tiles=1,dev_cdb=32768,variant=7" > gpurun_out/s2_exp_rmat24.jsonl 2> gpurun_out/s2_exp_rmat24.err; echo "exp rmat exit $?"; grep -v "^generated" gpurun_out/s2_exp_rmat24.err | tail -8
timeout 600 python scripts/exp_options.py rmat 24 f32 "" "variant=7" "dev_tiles=1" "dev_tiles=1,variant=7" > gpurun_out/s2_exp_rmat24_f32.jsonl 2> gpurun_out/s2_exp_rmat24_f32.err; echo "exp rmat f32 exit $?"; grep -v "^generated" gpurun_out/s2_exp_rmat24_f32.err | tail -8
timeout 600 python scripts/exp_options.py laplacian 22 f64 "" "variant=8" > gpurun_out/s2_exp_lap.jsonl 2> gpurun_out/s2_exp_lap.err; echo "exp lap exit $?"; grep -v "^generated" gpurun_out/s2_exp_lap.err | tail -4
( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 ) > gpurun_out/s2_pytest_gpu.log 2>&1; echo "gpu suite exit $?"; tail -18 gpurun_out/s2_pytest_gpu.log
