#!/bin/bash
# weight-sharing clusters (MPU_FWD_CLW: 0 never, 1 single-channel-tile convs, 2 also instead of slab sharing on 3-tap-slab levels)
out=gpurun_out/${1:-clw_ab}
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet_baseline.py tests/test_gpu_unet.py tests/test_gpu_variants.py -q -m gpu -x 2>&1 | tail -4
(for v in 0 1 2; do echo "=== MPU_FWD_CLW=$v"; MPU_FWD_CLW=$v timeout 200 python tests/perf_gemm.py perf_L0 perf_L0cat perf_L1 2>&1 | grep -E "perf|CTA0"; done) | tee $out/perf.txt
for v in 0 1 2 0 1; do
  MPU_FWD_CLW=$v timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 30 > $out/bench_clw$v.json 2>$out/bench.err
  python -c "
import json;d=json.loads(open('$out/bench_clw$v.json').read().strip().split(chr(10))[-1]);print('clw=$v',round(d['value'],1),round(d['ms_per_step'],3),round(d['roofline']['frac'],4),d['clocks']['sm_mhz'])"
done
