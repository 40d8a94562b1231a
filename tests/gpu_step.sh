#!/bin/bash
# GPU box: kernel + network parity tests, ncu launch list of one train step with DRAM bytes, bench line
out=gpurun_out/${1:-step}
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet_baseline.py tests/test_gpu_unet.py tests/test_gpu_variants.py -q -m gpu -x 2>&1 | tail -6
timeout 800 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $out/launches_traffic.csv python tests/perf_unet.py --ncu > $out/ncu.log 2>&1
python tests/launch_summary.py $out/launches_traffic.csv | tee $out/summary.txt
timeout 300 python bench.py --no-cpu-baseline > $out/bench.json 2>$out/bench.err
python -c "
import json;d=json.loads(open('$out/bench.json').read().strip().split(chr(10))[-1]);print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['clocks'], 'e2e', d['e2e']['value'])"
