#!/bin/bash
# GPU box, one B200: whole GPU suite, smoke(), full bench line + reference arm, ncu launch list with DRAM bytes of one train step
out=gpurun_out/${1:-final}
mkdir -p $out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > $out/pytest_gpu.log
tail -4 $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
timeout 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
timeout 600 python bench.py --workload predict --no-cpu-baseline > $out/bench_predict.json 2> $out/bench_predict.err
timeout 600 python bench.py --workload train_fusion --no-cpu-baseline > $out/bench_fusion.json 2> $out/bench_fusion.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file $out/launches_traffic.csv python tests/perf_unet.py --ncu > $out/ncu.log 2>&1
python tests/ncu_traffic.py $out/launches_traffic.csv $out/gemm_traffic.json > $out/traffic_summary.txt 2>&1
python tests/launch_summary.py $out/launches_traffic.csv > $out/launch_summary.txt 2>&1
python - <<PY
import json
for f in ("bench_n1", "bench_reference", "bench_predict", "bench_fusion"):
    try:
        d = json.loads(open("$out/%s.json" % f).read().strip().split("\n")[-1])
        print(f, {k: d.get(k) for k in ("metric", "value", "unit", "ms_per_step", "impl")}, "e2e", d.get("e2e", {}).get("value"),
              "frac", (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "failed", e)
PY
cat $out/traffic_summary.txt
# ncu --set full of the BatchNorm / elementwise kernels of one train step (summaries only; ~7 minutes: only with a 2nd argument)
[ -n "$2" ] || exit 0
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"bn_|head_kernel|conv_first|adam" \
  -f -o $out/elementwise python tests/perf_unet.py --ncu > $out/ncu_elementwise.log 2>&1
python tests/ncu_summary.py $out/elementwise.ncu-rep > $out/elementwise_summary.txt 2>&1
rm -f $out/elementwise.ncu-rep
