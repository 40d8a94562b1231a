#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_n4.log 2>&1
echo "rc=$?"
tail -n 1 gpurun_out/bench_n4.log | cut -c1-400
