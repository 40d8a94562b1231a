#!/bin/bash
# GPU box: fusion-training tests, then the train_fusion bench with the persistent epoch kernel on / off and a block sweep
out=gpurun_out/${1:-fusion}
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -12 > $out/tests.log; tail -5 $out/tests.log
for cfg in "1 0" "0 0" "1 148" "1 296" "1 444" "1 74"; do
  set -- $cfg
  MPU_FUSION_EPOCH_PERSISTENT=$1 MPU_FUSION_EPOCH_BLOCKS=$2 timeout 300 python bench.py --workload train_fusion --no-cpu-baseline > $out/fusion_p$1_b$2.json 2> $out/fusion_p$1_b$2.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/fusion_p$1_b$2.json").read().strip().split("\n")[-1])
    print("persistent=$1 blocks=$2", "GB/s", round(d["value"], 1), "ms/epoch", round(d["ms_per_step"], 3), "kernel frac", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"])
except Exception as e:
    print("persistent=$1 blocks=$2 failed", e); print(open("$out/fusion_p$1_b$2.err").read()[-600:])
PY
done
