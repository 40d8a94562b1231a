#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"
tail -n 8 gpurun_out/pytest_gpu.log
timeout 200 python tests/perf_unet.py > gpurun_out/perf.log 2>&1
echo "perf rc=$?"
tail -n 12 gpurun_out/perf.log
