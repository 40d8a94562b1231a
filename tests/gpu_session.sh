#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_unet.py -x -q 2>&1 | tail -3
echo "##### bench N=1"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_n1.json; cat gpurun_out/bench_n1.json
echo "##### launch list"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1e.csv python tests/perf_unet.py --ncu --warmup 2 > gpurun_out/ncu_run.log 2>&1
python tests/agg_launches.py gpurun_out/launches_r1e.csv | head -22
echo "##### ncu full fwd L3 / L0 / wgrad L1"
for c in perf_L3 perf_L0 wperf_L1; do
  k=mtgemm_fwd; [[ $c == w* ]] && k=mtgemm_wgrad
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof5_$c python tests/bringup_gemm.py --case $c > gpurun_out/ncu_$c.log 2>&1
done
ls gpurun_out/prof5_*
