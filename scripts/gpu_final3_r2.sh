#!/bin/bash
# round 2, last single-GPU evidence (tree after the multi-GPU end-to-end call): smoke, the whole GPU suite, default bench,
# reference arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/h_smoke.log
( time timeout 600 python -m pytest tests -x -q -m gpu ) > gpurun_out/h_pytest_gpu.log 2>&1; echo "gpu suite exit $?"; tail -4 gpurun_out/h_pytest_gpu.log
( time timeout 300 python bench.py ) > gpurun_out/h_bench_default.json 2> gpurun_out/h_bench_default.err; echo "bench exit $?"
( time timeout 120 python bench.py --impl reference ) > gpurun_out/h_bench_reference.json 2> gpurun_out/h_bench_reference.err; echo "reference arm exit $?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/h_bench_default.json").read().strip().splitlines()[-1])
print("variant", d["engine"]["variant"], "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e %.1f" % d["e2e"]["value"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "check", d["check"]["max_err_over_tolerance"])
for k, v in (d.get("also") or {}).items():
    print("   also", k, "variant", v["engine"]["variant"], "ms/step %.4f" % v["ms_per_step"], "frac %.3f" % v["roofline"]["frac"], "e2e %.1f" % v["e2e"]["value"])
r = json.loads(open("gpurun_out/h_bench_reference.json").read().strip().splitlines()[-1])
print("reference", r["value"], r["unit"], r["cpu_baseline"]["cores"])
P
