#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_kernels.py -x -q -k "volume or mapping or map_fuse" 2>&1 | tail -3
echo "##### bench N=2"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -2 | cut -c1-900
echo "##### bench N=1 (extras)"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['e2e']['value'], json.dumps(d['extras']))"
