#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"
tail -n 15 gpurun_out/pytest_gpu.log
echo "#### overlap off"
MPU_OVERLAP=0 timeout 200 python tests/perf_unet.py 2>&1 | tail -n 6
echo "#### overlap on"
timeout 200 python tests/perf_unet.py 2>&1 | tail -n 6
echo "#### bench"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1800
