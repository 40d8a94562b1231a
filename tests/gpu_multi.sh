#!/bin/bash
# GPU box with >= 2 GPUs (gpurun --gpus 2): data-parallel correctness on hardware, the reference-graph pins, and the
# bench lines at N = 2 (train step with the gradient all-reduce; fusion training with the in-kernel peer exchange)
out=gpurun_out/${1:-multi}
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_dp.py tests/test_gpu_pins.py -q -m gpu -x 2>&1 | tail -6 | tee $out/pytest_dp.log
for wl in train train_fusion; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --workload $wl --no-cpu-baseline > $out/bench_${wl}_n2.json 2> $out/bench_${wl}_n2.err
  python -c "
import json;d=json.loads(open('$out/bench_${wl}_n2.json').read().strip().split(chr(10))[-1]);print('$wl N=2',d['value'],d['unit'],d['ms_per_step'],d.get('scaling'),(d.get('roofline') or {}).get('frac'), d.get('extras',{}).get('train_fusion',{}).get('exchange') if isinstance(d.get('extras'),dict) else None)"
done
