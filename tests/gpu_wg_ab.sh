#!/bin/bash
# same-box A/B of the wgrad split rules (MPU_WG_OLD_SPLITS=1: round 1's "about two CTAs per SM") + kernel / parity tests
out=gpurun_out/${1:-wg_ab}
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet_baseline.py -q -m gpu -x 2>&1 | tail -4
(echo "=== old split rule (MPU_WG_OLD_SPLITS=1), new drain"
 MPU_WG_OLD_SPLITS=1 timeout 200 python tests/perf_gemm.py wperf_L0 wperf_L1 wperf_L2 wperf_L3 wperf_L4
 echo "=== wave-aware splits"
 timeout 200 python tests/perf_gemm.py wperf_L0 wperf_L1 wperf_L2 wperf_L3 wperf_L4) > $out/perf.txt 2>&1
cat $out/perf.txt
for v in 1 0; do
  MPU_WG_OLD_SPLITS=$v timeout 300 python bench.py --no-cpu-baseline > $out/bench_old$v.json 2>$out/bench.err
  python -c "
import json;d=json.loads(open('$out/bench_old$v.json').read().strip().split(chr(10))[-1]);print('old_splits=$v',d['value'],d['ms_per_step'],d['roofline']['frac'],d['clocks'])"
done
