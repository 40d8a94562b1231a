#!/bin/bash
# Helper run on the GPU box: tests of the U-Net path + GEMM perf table + one bench line into gpurun_out/$1
out=gpurun_out/${1:-run}
mkdir -p $out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet_baseline.py tests/test_gpu_unet.py -q -m gpu -x 2>&1 | tail -15 > $out/tests.log
python tests/perf_gemm.py > $out/perf_gemm.txt 2>&1
python bench.py --no-cpu-baseline --no-extras > $out/bench.json 2> $out/bench.err
tail -5 $out/tests.log
cat $out/bench.json
