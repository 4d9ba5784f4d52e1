#!/bin/bash
# refresh of the committed bench line (default config) + single-GPU power iteration through the host driver
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/final_bench.json'))
print('ms/step %.4f GF %.1f frac %.3f e2e %.1f cpu %.2f gpu_build %s'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['setup_s'].get('gpu_layout_build')))"
timeout 200 python bench.py --workload poweriter --dtype f32 --scale 20 --steps 20 --warmup 3 > gpurun_out/pi_1_s20.json 2> gpurun_out/pi_1_s20.err; echo "poweriter exit $?"
cat gpurun_out/pi_1_s20.json | cut -c1-400
