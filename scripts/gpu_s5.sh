#!/bin/bash
# round 2, GPU session 5: bounds-checked kernels on the small cases, per-tile launches with a persisting L2 window
mkdir -p gpurun_out
for cfg in 0 1 2; do
  SPMVB_LIB=$PWD/spmv-fpga_b200/lib/libspmvb_check.so SANITIZE_OPTS=xs_config=$cfg timeout 600 python scripts/sanitize_case.py > gpurun_out/s5_boundscheck_cfg$cfg.log 2>&1; echo "bounds-checked small cases xs_config=$cfg exit $?"; tail -2 gpurun_out/s5_boundscheck_cfg$cfg.log
done
SPMVB_LIB=$PWD/spmv-fpga_b200/lib/libspmvb_check.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "engine_private or cu_major or small_column or kat6x6 or ragged or onerow or longrow" > gpurun_out/s5_boundscheck_pytest.log 2>&1; echo "bounds-checked parity subset exit $?"; tail -3 gpurun_out/s5_boundscheck_pytest.log
timeout 900 python scripts/exp_options.py uniform 26 f64 "tile_launch=1" "tile_launch=1,l2_persist_mb=79" "tile_launch=1,l2_persist_mb=79,tile_mb=32" "tile_launch=1,l2_persist_mb=79,tile_mb=48" "tile_launch=1,l2_persist_mb=79,tile_mb=64" "l2_persist_mb=79" > gpurun_out/s5_exp_uniform26.jsonl 2> gpurun_out/s5_exp_uniform26.err; echo "exp uniform26 exit $?"; grep -v "^generated" gpurun_out/s5_exp_uniform26.err | tail -12
