#!/bin/bash
# new tests first (cheap); the rest of the GPU suite only if they pass
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_cg.py tests/test_gpu_dropin.py -x -q -s > gpurun_out/pytest_new.log 2>&1; rc=$?
echo "new tests exit $rc"; tail -15 gpurun_out/pytest_new.log
if [ $rc -eq 0 ]; then
  timeout 400 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_cg.py --deselect tests/test_gpu_dropin.py > gpurun_out/pytest_gpu.log 2>&1; echo "gpu suite exit $?"
  tail -4 gpurun_out/pytest_gpu.log
fi
