#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "##### perf"
timeout 300 python tests/perf_unet.py 2>&1 | tail -6
echo "##### perf with MPU_DGRAD_MN=0"
MPU_DGRAD_MN=0 timeout 300 python tests/perf_unet.py 2>&1 | tail -5
