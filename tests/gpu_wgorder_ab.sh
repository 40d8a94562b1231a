#!/bin/bash
# wgrad launch order: partial-tile CTAs interleaved by K position (default) vs a separate region at the end (MPU_WG_NO_ORDER=1)
out=gpurun_out/${1:-wgorder}
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet_baseline.py tests/test_gpu_unet.py tests/test_gpu_variants.py -q -m gpu -x 2>&1 | tail -3
for v in 1 0; do
  MPU_WG_NO_ORDER=$v timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:wgrad --csv --log-file $out/wgrad_traffic_noorder$v.csv python tests/perf_unet.py --ncu > $out/ncu$v.log 2>&1
  echo "MPU_WG_NO_ORDER=$v"; python tests/launch_summary.py $out/wgrad_traffic_noorder$v.csv | grep -E "wgrad|total"
done
for v in 1 0 1 0; do
  MPU_WG_NO_ORDER=$v timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 30 > $out/bench_noorder$v.json 2>$out/bench.err
  python -c "
import json;d=json.loads(open('$out/bench_noorder$v.json').read().strip().split(chr(10))[-1]);print('no_order=$v',round(d['value'],1),round(d['ms_per_step'],3),round(d['roofline']['frac'],4),d['clocks']['sm_mhz'])"
done
