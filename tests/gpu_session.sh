#!/bin/bash
# full GPU validation used during the round: parity tests through the C ABI, smoke(), a short bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -W ignore -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 25 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-600
