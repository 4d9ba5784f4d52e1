#!/bin/bash
# round 2, GPU session 16: the whole GPU suite with the wide and ELL images in, smoke
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s16_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/s16_smoke.log
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/s16_pytest_gpu.log 2>&1; echo "gpu suite exit $?"; tail -6 gpurun_out/s16_pytest_gpu.log
