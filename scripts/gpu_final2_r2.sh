#!/bin/bash
# round 2, final evidence on one GPU after the ELL image: smoke, the whole GPU suite, default bench + reference arm, the
# Laplacian with the GPU layout builder, launch list of the default bench, ncu --set full of the ELL kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/g_smoke.log
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/g_pytest_gpu.log 2>&1; echo "gpu suite exit $?"; tail -4 gpurun_out/g_pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/g_bench_default.json 2> gpurun_out/g_bench_default.err; echo "bench exit $?"
( time timeout 600 python bench.py --impl reference ) > gpurun_out/g_bench_reference.json 2> gpurun_out/g_bench_reference.err; echo "reference arm exit $?"
timeout 600 python bench.py --workload laplacian --gpu-build --steps 200 > gpurun_out/g_bench_laplacian.json 2> gpurun_out/g_bench_laplacian.err; echo "laplacian exit $?"
python - <<'P'
import json
for f in ("g_bench_default", "g_bench_laplacian"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "variant", d["engine"]["variant"], "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e %.1f" % d["e2e"]["value"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
        for k, v in (d.get("also") or {}).items():
            print("   also", k, "variant", v["engine"]["variant"], "ms/step %.4f" % v["ms_per_step"], "frac %.3f" % v["roofline"]["frac"], "e2e %.1f" % v["e2e"]["value"])
    except Exception as e:
        print(f, "failed", e)
P
SMALL="--steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 600 python bench.py $SMALL --workload laplacian > gpurun_out/g_plain_lap.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_ell -s 4 -c 1 -f -o gpurun_out/g_prof_laplacian_ell python bench.py $SMALL --workload laplacian > gpurun_out/g_ncu_lap.log 2>&1; echo "ncu lap exit $?"
timeout 900 python bench.py $SMALL --also "" > gpurun_out/g_plain_uniform.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/g_launches_default.csv python bench.py $SMALL --also "" > gpurun_out/g_ncu_launches.log 2>&1; echo "launch list exit $?"
ls -la gpurun_out/g_*.ncu-rep
