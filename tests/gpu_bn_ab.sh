#!/bin/bash
# BN coefficients / parameter gradients inside the apply kernels (default) vs separate single-block launches
# (MPU_BN_SEPARATE=1): parity tests + same-box A/B
out=gpurun_out/${1:-bn_ab}
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_unet_baseline.py tests/test_gpu_unet.py tests/test_gpu_variants.py tests/test_gpu_pins.py -q -m gpu -x 2>&1 | tail -4
for v in 1 0 1 0; do
  MPU_BN_SEPARATE=$v timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 30 > $out/bench_sep$v.json 2>$out/bench.err
  python -c "
import json;d=json.loads(open('$out/bench_sep$v.json').read().strip().split(chr(10))[-1]);print('bn_separate=$v',round(d['value'],1),round(d['ms_per_step'],3),round(d['roofline']['frac'],4),d['clocks']['sm_mhz'],d['gpu_launches'])"
done
