#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "##### smoke"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "##### bench"
python bench.py --steps 10 --warmup 3 2>&1 | tail -3
