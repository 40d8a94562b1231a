#!/bin/bash
# same-box A/B: BatchNorm-backward sums / bias column sums in the GEMM epilogue from level L on (MPU_EPI_RED_LEVEL)
out=gpurun_out/${1:-red_ab}
mkdir -p $out
for lv in 99 2 3 4 99; do
  MPU_EPI_RED_LEVEL=$lv timeout 300 python bench.py --no-cpu-baseline --steps 30 > $out/bench_red$lv.json 2>$out/bench.err
  python -c "
import json;d=json.loads(open('$out/bench_red$lv.json').read().strip().split(chr(10))[-1]);print('red_level=$lv',round(d['value'],1),round(d['ms_per_step'],3),round(d['roofline']['frac'],4),d['clocks']['sm_mhz'])"
done
