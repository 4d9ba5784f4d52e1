#!/bin/bash
# round 2, GPU session 9: row ids staged through the ring slot (coalesced loads)
mkdir -p gpurun_out
for cfg in 0 1 2; do
  SPMVB_LIB=$PWD/spmv-fpga_b200/lib/libspmvb_check.so SANITIZE_OPTS=xs_config=$cfg timeout 400 python scripts/sanitize_case.py > gpurun_out/s9_boundscheck_cfg$cfg.log 2>&1; rc=$?; echo "bounds-checked small cases xs_config=$cfg exit $rc"; tail -2 gpurun_out/s9_boundscheck_cfg$cfg.log
  if [ $rc -ne 0 ]; then echo "ABORT"; tail -30 gpurun_out/s9_boundscheck_cfg$cfg.log; exit 1; fi
done
timeout 900 python scripts/exp_options.py uniform 26 f64 "" "stage_ids=0" "tile_mb=32" > gpurun_out/s9_exp_uniform26.jsonl 2> gpurun_out/s9_exp_uniform26.err; echo "exp uniform26 exit $?"; grep -v "^generated\|Warning\|err = " gpurun_out/s9_exp_uniform26.err | tail -6
timeout 900 python scripts/exp_options.py rmat 24 f64 "" "stage_ids=0" "variant=8" "variant=8,stage_ids=0" "dev_tiles=1,dev_cdb=32768,variant=7" "dev_tiles=1,dev_cdb=32768,variant=7,stage_ids=0" > gpurun_out/s9_exp_rmat24.jsonl 2> gpurun_out/s9_exp_rmat24.err; echo "exp rmat exit $?"; grep -v "^generated" gpurun_out/s9_exp_rmat24.err | tail -8
timeout 600 python scripts/exp_options.py rmat 24 f32 "" "stage_ids=0" > gpurun_out/s9_exp_rmat24_f32.jsonl 2> gpurun_out/s9_exp_rmat24_f32.err; echo "exp rmat f32 exit $?"; grep -v "^generated" gpurun_out/s9_exp_rmat24_f32.err | tail -4
timeout 600 python scripts/exp_options.py laplacian 22 f64 "" > gpurun_out/s9_exp_lap.jsonl 2> gpurun_out/s9_exp_lap.err; echo "exp lap exit $?"; grep -v "^generated" gpurun_out/s9_exp_lap.err | tail -2
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/s9_pytest_gpu.log 2>&1; echo "gpu suite exit $?"; tail -5 gpurun_out/s9_pytest_gpu.log
