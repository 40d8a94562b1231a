#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"
tail -n 15 gpurun_out/pytest_gpu.log
timeout 200 python tests/perf_unet.py > gpurun_out/perf.log 2>&1
echo "perf rc=$?"
tail -n 7 gpurun_out/perf.log
echo "##### launch list"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1g.csv python tests/perf_unet.py --ncu --warmup 2 > gpurun_out/ncu_run.log 2>&1
python tests/agg_launches.py gpurun_out/launches_r1g.csv | head -24
