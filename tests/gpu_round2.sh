#!/bin/bash
# GPU box: U-Net tests, then the train bench with the epilogue reductions enabled from level 0 / 1 / 2 / 3 / never
out=gpurun_out/${1:-run}
mkdir -p $out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet_baseline.py tests/test_gpu_unet.py tests/test_gpu_variants.py -q -m gpu 2>&1 | tail -25 > $out/tests.log
for lv in 0 1 2 3 9; do
  MPU_EPI_RED_LEVEL=$lv python bench.py --no-cpu-baseline --no-extras > $out/bench_red$lv.json 2> $out/bench_red$lv.err
done
python tests/perf_gemm.py > $out/perf_gemm.txt 2>&1
tail -25 $out/tests.log
python - <<PY
import json
for lv in (0, 1, 2, 3, 9):
    try:
        d = json.loads(open("$out/bench_red%d.json" % lv).read().strip().split("\n")[-1])
        print("red level", lv, "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
              "frac", round(d["roofline"]["frac"], 4), d["roofline"]["kernels"])
    except Exception as e:
        print(lv, "failed", e)
PY
grep -h "perf" $out/perf_gemm.txt
