#!/bin/bash
# GPU layout builder: its own tests first (short timeout), then the whole GPU suite
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_layout_build.py -x -q -s > gpurun_out/pytest_layout_gpu.log 2>&1; echo "layout tests exit $?"
tail -25 gpurun_out/pytest_layout_gpu.log
timeout 600 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_layout_build.py > gpurun_out/pytest_gpu.log 2>&1; echo "gpu suite exit $?"
tail -5 gpurun_out/pytest_gpu.log
