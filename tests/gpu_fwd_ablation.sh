#!/bin/bash
# Ablation of mtgemm_fwd_kernel on the GPU box: which supply limits the MMA thread?  MPU_FWD_DEBUG bits (profiling
# build of the kernel only): 1 = epilogue drains nothing, 2 = no slab (activation) TMA loads, 4 = no weight TMA loads.
out=gpurun_out/${1:-ablate}
mkdir -p $out
for f in 0 1 2 4 6 7; do
  echo "# MPU_FWD_DEBUG=$f" >> $out/fwd_ablation.txt
  MPU_FWD_DEBUG=$f python tests/perf_gemm.py perf_L0 perf_L1 perf_L2 perf_L3 perf_L4 >> $out/fwd_ablation.txt 2>&1
done
cat $out/fwd_ablation.txt
