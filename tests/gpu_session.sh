#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tests/bringup_gemm.py --only wgrad_small,wgrad_split,wgrad_c96,wgrad_n256,wperf_L0,wperf_L1,wperf_L2,wperf_L3,wperf_L4 2>&1 | grep -E "^===|perf|SUMMARY|Error|error|bad|max_err"
echo "##### unet"
python -m pytest tests/test_gpu_unet.py -x -q 2>&1 | tail -3
echo "##### perf"
python tests/perf_unet.py 2>&1 | tail -6
