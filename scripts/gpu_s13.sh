#!/bin/bash
# round 2, GPU session 13: sliced-ELLPACK image for regular matrices (BASELINE configs[1]) + its end-to-end pipeline
mkdir -p gpurun_out
SPMVB_LIB=$PWD/spmv-fpga_b200/lib/libspmvb_check.so timeout 600 python scripts/sanitize_case.py > gpurun_out/s13_boundscheck.log 2>&1; rc=$?; echo "bounds-checked small cases exit $rc"; tail -3 gpurun_out/s13_boundscheck.log
if [ $rc -ne 0 ]; then echo "ABORT"; tail -30 gpurun_out/s13_boundscheck.log; exit 1; fi
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "ell or wide or config2" ) > gpurun_out/s13_pytest.log 2>&1; rc=$?; echo "ell / wide parity tests exit $rc"; tail -4 gpurun_out/s13_pytest.log
if [ $rc -ne 0 ]; then echo "ABORT"; tail -40 gpurun_out/s13_pytest.log; exit 1; fi
timeout 600 python scripts/exp_options.py laplacian 22 f64 "" "variant=7" "variant=10" > gpurun_out/s13_exp_lap.jsonl 2> gpurun_out/s13_exp_lap.err; echo "exp lap exit $?"; grep -v "^generated" gpurun_out/s13_exp_lap.err | tail -4
timeout 600 python scripts/exp_options.py laplacian 22 f32 "variant=7" "variant=10" > gpurun_out/s13_exp_lap_f32.jsonl 2> gpurun_out/s13_exp_lap_f32.err; echo "exp lap f32 exit $?"; grep -v "^generated" gpurun_out/s13_exp_lap_f32.err | tail -3
timeout 600 python bench.py --workload laplacian --steps 200 > gpurun_out/s13_bench_laplacian.json 2> gpurun_out/s13_bench_laplacian.err; echo "bench laplacian exit $?"; tail -3 gpurun_out/s13_bench_laplacian.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/s13_bench_laplacian.json").read().strip().splitlines()[-1])
print("laplacian: variant", d["engine"]["variant"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms_avg"], "e2e", d["e2e"], "tuned", d["engine"]["device_layout"]["tuned_us"])
P
SMALL="--steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_ell -s 4 -c 1 -f -o gpurun_out/s13_prof_laplacian_ell python bench.py $SMALL --workload laplacian > gpurun_out/s13_ncu_lap.log 2>&1; echo "ncu lap exit $?"
ls -la gpurun_out/s13*.ncu-rep
