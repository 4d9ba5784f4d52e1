#!/bin/bash
# round 2, GPU session 6: where does the x-window kernel's time go on the 1 B-nnz matrix (diagnostic flags: wrong results)
mkdir -p gpurun_out
timeout 900 python scripts/exp_options.py uniform 26 f64 "" "diag_flags=16" "diag_flags=32" "diag_flags=48" > gpurun_out/s6_exp_uniform26_diag.jsonl 2> gpurun_out/s6_exp_uniform26_diag.err; echo "exp exit $?"; grep -v "^generated" gpurun_out/s6_exp_uniform26_diag.err | tail -6
