#!/bin/bash
out=gpurun_out/${1:-dbg}
mkdir -p $out
python -m pytest "tests/test_gpu_unet_baseline.py::test_train_step_layer_local_residuals_and_gradients[small_cf0125_64]" -q -m gpu -x 2>&1 | grep -v "^param grad\|^fwd \|^bwd " | tail -60 > $out/small_default.log
MPU_FWD_WIDE=0 python -m pytest "tests/test_gpu_unet_baseline.py::test_train_step_layer_local_residuals_and_gradients[small_cf0125_64]" -q -m gpu -x 2>&1 | grep -v "^param grad\|^fwd \|^bwd " | tail -30 > $out/small_nowide.log
python -m pytest tests/test_gpu_unet.py tests/test_gpu_variants.py -q -m gpu 2>&1 | tail -8 > $out/unet_variants.log
cat $out/small_default.log; echo ----; cat $out/small_nowide.log; echo ----; cat $out/unet_variants.log
