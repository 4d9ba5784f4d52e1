#!/bin/bash
# round 2, GPU session 15: the wide image on R-MAT with narrower column blocks (x ranges of 1-16 MB)
mkdir -p gpurun_out
timeout 900 python scripts/exp_options.py rmat 24 f64 "variant=7" "wide=1,variant=9,wide_range_log2=17" "wide=1,variant=9,wide_range_log2=18" "wide=1,variant=9,wide_range_log2=19" "wide=1,variant=9,wide_range_log2=20" "wide=1,variant=9,wide_range_log2=21" "wide=1,variant=9,wide_range_log2=19,wide_hints=0" > gpurun_out/s15_exp_rmat24.jsonl 2> gpurun_out/s15_exp_rmat24.err; echo "exp rmat exit $?"; grep -v "^generated" gpurun_out/s15_exp_rmat24.err | tail -8
timeout 600 python scripts/exp_options.py rmat 24 f32 "wide=1,variant=9,wide_range_log2=18" "wide=1,variant=9,wide_range_log2=20" > gpurun_out/s15_exp_rmat24_f32.jsonl 2> gpurun_out/s15_exp_rmat24_f32.err; echo "exp rmat f32 exit $?"; grep -v "^generated" gpurun_out/s15_exp_rmat24_f32.err | tail -4
python - <<'P'
import json
for f in ("gpurun_out/s15_exp_rmat24.jsonl", "gpurun_out/s15_exp_rmat24_f32.jsonl"):
    for l in open(f):
        d = json.loads(l)
        print(d["options"], d["kernel_ms"], "pairs", d["device_layout"]["pairs"], "chunks", d["device_layout"]["chunks"], "zero_rows", d["device_layout"]["zero_rows"])
P
