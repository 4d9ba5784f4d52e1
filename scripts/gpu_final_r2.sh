#!/bin/bash
# round 2, final evidence on one GPU: smoke, default bench + reference arm, launch list and ncu --set full captures of the
# shipped kernels on BASELINE configs[3] / [2] / [1] (each only after the same command exited 0 without ncu)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/f_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/f_bench_default.json 2> gpurun_out/f_bench_default.err; echo "bench exit $?"
( time timeout 600 python bench.py --impl reference ) > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err; echo "reference arm exit $?"
timeout 600 python bench.py --workload band --steps 20 > gpurun_out/f_bench_band.json 2> gpurun_out/f_bench_band.err; echo "band exit $?"
timeout 600 python bench.py --workload band --impl reference --steps 20 > gpurun_out/f_bench_band_reference.json 2> gpurun_out/f_bench_band_reference.err; echo "band reference exit $?"
timeout 600 python bench.py --workload poweriter --steps 100 > gpurun_out/f_bench_poweriter_1gpu.json 2> gpurun_out/f_bench_poweriter_1gpu.err; echo "poweriter exit $?"
timeout 600 python bench.py --workload laplacian --gpu-build --steps 200 > gpurun_out/f_bench_laplacian.json 2> gpurun_out/f_bench_laplacian.err; echo "laplacian exit $?"
SMALL="--steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 900 python bench.py $SMALL --also "" > gpurun_out/f_plain_uniform.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/f_launches_default.csv python bench.py $SMALL --also "" > gpurun_out/f_ncu_launches.log 2>&1; echo "launch list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:spmv_xs -s 6 -c 1 -f -o gpurun_out/f_prof_uniform1B_xs python bench.py $SMALL --also "" > gpurun_out/f_ncu_uniform.log 2>&1; echo "ncu uniform exit $?"
timeout 600 python bench.py $SMALL --workload rmat > gpurun_out/f_plain_rmat.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_occ -s 10 -c 1 -f -o gpurun_out/f_prof_rmat24_occ python bench.py $SMALL --workload rmat > gpurun_out/f_ncu_rmat.log 2>&1; echo "ncu rmat exit $?"
timeout 600 python bench.py $SMALL --workload laplacian > gpurun_out/f_plain_lap.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_occ -s 4 -c 1 -f -o gpurun_out/f_prof_laplacian_occ python bench.py $SMALL --workload laplacian > gpurun_out/f_ncu_lap.log 2>&1; echo "ncu lap exit $?"
ls -la gpurun_out/f_*.ncu-rep
