#!/bin/bash
# full GPU validation: parity tests through the C ABI, smoke(), a short bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -W ignore -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 25 gpurun_out/pytest_gpu.log | cut -c1-300
