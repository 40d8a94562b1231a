#!/bin/bash
# conv1 bias gradients inside the wgrad kernel (default) vs the separate column-sum pass (MPU_BIAS_COLSUM=1): tests + A/B
out=gpurun_out/${1:-cs_ab}
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet_baseline.py tests/test_gpu_unet.py tests/test_gpu_variants.py -q -m gpu -x 2>&1 | tail -6
for v in 1 0 1 0; do
  MPU_BIAS_COLSUM=$v timeout 300 python bench.py --no-cpu-baseline --steps 30 > $out/bench_cs$v.json 2>$out/bench.err
  python -c "
import json;d=json.loads(open('$out/bench_cs$v.json').read().strip().split(chr(10))[-1]);print('colsum_pass=$v',round(d['value'],1),round(d['ms_per_step'],3),round(d['roofline']['frac'],4),d['clocks']['sm_mhz'])"
done
