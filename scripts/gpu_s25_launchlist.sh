#!/bin/bash
# round 2, GPU session 25 (1 GPU, short): the drop-in executables after the driver was restructured; ncu launch list of the power iteration (configs[4] on one GPU), 3 + 3 iterations
mkdir -p gpurun_out
( time timeout 60 python -m pytest tests/test_gpu_dropin.py -x -q -m gpu ) > gpurun_out/s25_pytest_dropin.log 2>&1; echo "drop-in tests exit $?"; tail -4 gpurun_out/s25_pytest_dropin.log
timeout 80 python bench.py --workload poweriter --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s25_plain_poweriter.log 2>&1; echo "plain exit $?"
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s25_launches_poweriter.csv python bench.py --workload poweriter --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s25_ncu_poweriter.log 2>&1; echo "launch list exit $?"
tail -12 gpurun_out/s25_launches_poweriter.csv | cut -c 1-200
