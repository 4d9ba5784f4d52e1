#!/bin/bash
# round 2, GPU session 3: x-window kernel configurations, L2 set-aside, tile sizes
mkdir -p gpurun_out
timeout 300 python scripts/sanitize_case.py > gpurun_out/s3_smallcases.log 2>&1; rc=$?; echo "small cases exit $rc"; tail -2 gpurun_out/s3_smallcases.log
if [ $rc -ne 0 ]; then echo "ABORT: small cases failed"; tail -30 gpurun_out/s3_smallcases.log; exit 1; fi
for cfg in 1 2; do
  timeout 300 env SPMVB_XS_CONFIG=$cfg python -c "
import sys; sys.path.insert(0,'spmv-fpga_b200'); import spmvb; spmvb.set_option('xs_config', $cfg); exec(open('scripts/sanitize_case.py').read())" > gpurun_out/s3_smallcases_cfg$cfg.log 2>&1; echo "small cases xs_config=$cfg exit $?"; tail -1 gpurun_out/s3_smallcases_cfg$cfg.log
done
timeout 900 python scripts/exp_options.py uniform 26 f64 "tile_mb=24" "tile_mb=24,xs_config=1" "tile_mb=24,xs_config=2" "tile_mb=16,xs_config=1" "tile_mb=32,l2_persist_mb=64" "tile_mb=48,l2_persist_mb=96" "tile_mb=20" > gpurun_out/s3_exp_uniform26.jsonl 2> gpurun_out/s3_exp_uniform26.err; echo "exp uniform26 exit $?"; grep -v "^generated" gpurun_out/s3_exp_uniform26.err | tail -12
timeout 900 python scripts/exp_options.py rmat 24 f64 "" "xs_config=1" "xs_config=2" "xs_config=1,dev_tiles=1" "xs_config=2,dev_tiles=1" "xs_config=2,dev_tiles=8" > gpurun_out/s3_exp_rmat24.jsonl 2> gpurun_out/s3_exp_rmat24.err; echo "exp rmat exit $?"; grep -v "^generated" gpurun_out/s3_exp_rmat24.err | tail -8
timeout 600 python scripts/exp_options.py rmat 24 f32 "variant=8" "variant=8,xs_config=1" "variant=8,xs_config=2" > gpurun_out/s3_exp_rmat24_f32.jsonl 2> gpurun_out/s3_exp_rmat24_f32.err; echo "exp rmat f32 exit $?"; grep -v "^generated" gpurun_out/s3_exp_rmat24_f32.err | tail -8
timeout 600 python scripts/exp_options.py laplacian 22 f64 "" "variant=8" "variant=8,xs_config=1" > gpurun_out/s3_exp_lap.jsonl 2> gpurun_out/s3_exp_lap.err; echo "exp lap exit $?"; grep -v "^generated" gpurun_out/s3_exp_lap.err | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_xs -s 4 -c 1 -f -o gpurun_out/s3_prof_lap_xs python scripts/exp_options.py laplacian 22 f64 "variant=8" > gpurun_out/s3_ncu_lap.log 2>&1; echo "ncu lap exit $?"
