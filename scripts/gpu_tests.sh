#!/bin/bash
# the whole GPU suite through the C ABI (parity, drop-in executables, GPU layout builder, iterated callers)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "gpu suite exit $?"
tail -6 gpurun_out/pytest_gpu.log
