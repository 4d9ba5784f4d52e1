#!/bin/bash
# first GPU trip: parity tests, bench for both kernel variants, launch list
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
(nproc; free -g; lscpu | head -20) > gpurun_out/host.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo "bench v2 exit $?"
timeout 300 python bench.py --steps 100 --warmup 5 --variant 1 --no-cpu-baseline > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err; echo "bench v1 exit $?"
timeout 300 python bench.py --steps 100 --warmup 5 --flush-l2 --no-cpu-baseline > gpurun_out/bench_v2_flush.json 2> gpurun_out/bench_v2_flush.err; echo "bench flush exit $?"
cat gpurun_out/bench_v2.json gpurun_out/bench_v1.json gpurun_out/bench_v2_flush.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu exit $?"
