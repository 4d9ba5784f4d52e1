#!/bin/bash
# round 2, GPU session 14: software-pipelined ELL kernel; where the end-to-end time of configs[1] goes
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "ell or config2" ) > gpurun_out/s14_pytest.log 2>&1; rc=$?; echo "ell parity tests exit $rc"; tail -4 gpurun_out/s14_pytest.log
if [ $rc -ne 0 ]; then echo "ABORT"; tail -40 gpurun_out/s14_pytest.log; exit 1; fi
timeout 600 python scripts/exp_options.py laplacian 22 f64 "variant=7" "variant=10" > gpurun_out/s14_exp_lap.jsonl 2> gpurun_out/s14_exp_lap.err; echo "exp lap exit $?"; grep -v "^generated" gpurun_out/s14_exp_lap.err | tail -4
timeout 600 python scripts/exp_options.py laplacian 22 f32 "variant=7" "variant=10" > gpurun_out/s14_exp_lap_f32.jsonl 2> gpurun_out/s14_exp_lap_f32.err; echo "exp lap f32 exit $?"; grep -v "^generated" gpurun_out/s14_exp_lap_f32.err | tail -3
timeout 600 python scripts/exp_e2e.py > gpurun_out/s14_exp_e2e.jsonl 2> gpurun_out/s14_exp_e2e.err; echo "exp e2e exit $?"; cat gpurun_out/s14_exp_e2e.jsonl | cut -c1-200; tail -3 gpurun_out/s14_exp_e2e.err
