#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -W ignore -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"
grep -E "elastic|passed|failed|Error|error|assert" gpurun_out/pytest_gpu.log | tail -n 25
