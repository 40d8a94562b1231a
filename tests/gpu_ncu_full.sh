#!/bin/bash
# ncu --set full captures (one launch each) of the GEMM kernels at the levels that matter; summaries via tests/ncu_summary.py
out=gpurun_out/${1:-ncu_full}
mkdir -p $out
for c in perf_L0 perf_L0cat perf_L1 perf_L2 wperf_L0 wperf_L1 wperf_L2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:mtgemm --launch-skip 4 --launch-count 1 \
    -f -o $out/$c python tests/perf_gemm.py $c > $out/$c.log 2>&1
done
python tests/ncu_summary.py $out/*.ncu-rep > $out/summary.txt 2>&1
grep -E "^==|gpu__time_duration|utchmma|dram__bytes|lts__t_sector_hit|sm__cycles_elapsed" $out/summary.txt
