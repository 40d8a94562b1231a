#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -W ignore -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 200 python tests/perf_unet.py 2>&1 | tail -n 5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_quick.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['extras'])
PY
