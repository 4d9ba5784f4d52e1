#!/bin/bash
# round-end evidence: launch list + full ncu capture of the default bench, default bench line, reference arm
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref exit $?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-build --e2e-steps 2 > gpurun_out/plain_launches.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-build --e2e-steps 2 > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"

timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-build --e2e-steps 1 > gpurun_out/plain_full.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_occ -s 4 -c 1 -f -o gpurun_out/final_prof python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-build --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
python -c "import __graft_entry__ as g; g.smoke()"
