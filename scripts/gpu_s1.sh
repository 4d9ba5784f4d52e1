#!/bin/bash
# round 2, GPU session 1: the GPU suite, the default bench + reference arm, option A/Bs, first ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/s1_box.txt; nproc >> gpurun_out/s1_box.txt; free -g >> gpurun_out/s1_box.txt
# the new kernel paths on small cases first, with a short leash: a hang here must not take the box down with it
timeout 300 python scripts/sanitize_case.py > gpurun_out/s1_smallcases.log 2>&1; rc=$?; echo "small cases exit $rc"; tail -4 gpurun_out/s1_smallcases.log
if [ $rc -ne 0 ]; then echo "ABORT: small cases failed"; exit 1; fi
( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=12 ) > gpurun_out/s1_pytest_gpu.log 2>&1; echo "gpu suite exit $?"; tail -25 gpurun_out/s1_pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/s1_bench_default.json 2> gpurun_out/s1_bench_default.err; echo "bench exit $?"; tail -3 gpurun_out/s1_bench_default.err; head -c 3000 gpurun_out/s1_bench_default.json
( time timeout 600 python bench.py --impl reference --steps 10 --warmup 2 ) > gpurun_out/s1_bench_ref.json 2> gpurun_out/s1_bench_ref.err; echo "ref exit $?"; head -c 1500 gpurun_out/s1_bench_ref.json
timeout 900 python scripts/exp_options.py uniform 25 f64 "" "tile_mb=16" "tile_mb=64" "xs_rowids=0" "variant=7" > gpurun_out/s1_exp_uniform25.jsonl 2> gpurun_out/s1_exp_uniform25.err; echo "exp uniform exit $?"; grep -v "^generated" gpurun_out/s1_exp_uniform25.err | tail -8
timeout 900 python scripts/exp_options.py rmat 24 f64 "" "variant=7" "xs_rowids=0" "dev_tiles=4" "dev_cdb=32768,variant=7" > gpurun_out/s1_exp_rmat24.jsonl 2> gpurun_out/s1_exp_rmat24.err; echo "exp rmat exit $?"; grep -v "^generated" gpurun_out/s1_exp_rmat24.err | tail -8
timeout 600 python scripts/exp_options.py laplacian 22 f64 "" "variant=8" > gpurun_out/s1_exp_lap.jsonl 2> gpurun_out/s1_exp_lap.err; echo "exp lap exit $?"; grep -v "^generated" gpurun_out/s1_exp_lap.err | tail -4
# ncu: XS kernel on uniform / R-MAT scale 22 (full set with source), one launch each
for w in uniform rmat; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_xs -s 4 -c 1 -f -o gpurun_out/s1_prof_${w}22 python scripts/exp_options.py $w 22 f64 "variant=8" > gpurun_out/s1_ncu_${w}22.log 2>&1; echo "ncu $w exit $?"
done
