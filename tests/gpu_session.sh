#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tests/bringup_gemm.py 2>&1 | grep -E "^===|perf|CTA0|SUMMARY|Error|error|bad" | grep -v "PASS$"
echo "##### unet bring-up"
python tests/bringup_unet.py --dim 64 --batch 4 2>&1 | grep -v "^   [a-z0-9]*_[0-9] " | cut -c1-160 > gpurun_out/bringup_unet.log
grep -E "probs|loss|worst|adam|moving|Error|error" gpurun_out/bringup_unet.log
echo "##### perf"
python tests/perf_unet.py 2>&1 | tail -8
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1d.csv python tests/perf_unet.py --ncu --warmup 2 > gpurun_out/ncu_run.log 2>&1
python tests/agg_launches.py gpurun_out/launches_r1d.csv | head -24
