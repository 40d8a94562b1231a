#!/bin/bash
# GPU box: tests, then cluster slab-multicast (MPU_FWD_CL2 = 0 never | 2 policy | 1 whenever possible) perf + bench
out=gpurun_out/${1:-run}
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet_baseline.py tests/test_gpu_unet.py tests/test_gpu_variants.py -q -m gpu -x 2>&1 | tail -25 > $out/tests.log
tail -6 $out/tests.log
for m in 0 2 1; do
  echo "=== MPU_FWD_CL2=$m"
  MPU_FWD_CL2=$m timeout 300 python tests/perf_gemm.py perf_L1 perf_L3 perf_L4 perf_L0cat 2>&1 | tee $out/perf_cl2_$m.txt | grep -E "perf|CTA0"
  MPU_FWD_CL2=$m timeout 300 python bench.py --no-cpu-baseline --no-extras > $out/bench_cl2_$m.json 2> $out/bench_cl2_$m.err
done
python - <<PY
import json
for m in (0, 2, 1):
    try:
        d = json.loads(open("$out/bench_cl2_%d.json" % m).read().strip().split("\n")[-1])
        print("cl2", m, "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
              "frac", round(d["roofline"]["frac"], 4), d["roofline"]["kernels"])
    except Exception as e:
        print(m, "failed", e)
PY
