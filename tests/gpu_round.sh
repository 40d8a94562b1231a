#!/bin/bash
# Helper run on the GPU box: tests of the U-Net path + GEMM perf table + one bench line into gpurun_out/$1
out=gpurun_out/${1:-run}
mkdir -p $out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet_baseline.py tests/test_gpu_unet.py -q -m gpu -x 2>&1 | tail -15 > $out/tests.log
python tests/perf_gemm.py > $out/perf_gemm.txt 2>&1
MPU_FWD_WIDE=0 python tests/perf_gemm.py perf_L2 perf_L3 perf_L4 > $out/perf_gemm_nowide.txt 2>&1
python bench.py --no-cpu-baseline --no-extras > $out/bench.json 2> $out/bench.err
MPU_FWD_WIDE=0 python bench.py --no-cpu-baseline --no-extras > $out/bench_nowide.json 2> $out/bench_nowide.err
bash tests/gpu_fwd_ablation.sh ${1:-run} > /dev/null 2>&1
tail -5 $out/tests.log
grep -h "perf" $out/perf_gemm.txt | head -20
echo "--- nowide"; grep -h "perf" $out/perf_gemm_nowide.txt
python - <<PY
import json
for f in ("bench.json", "bench_nowide.json"):
    try:
        d = json.loads(open("$out/" + f).read().strip().split("\n")[-1])
        print(f, d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["kernels"])
    except Exception as e:
        print(f, "failed", e)
PY
