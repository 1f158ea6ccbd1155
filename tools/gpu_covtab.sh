#!/bin/bash
# A/B of the covariance table size / polynomial degree (TWXI_COV_TAB builds next to the baseline library).
mkdir -p gpurun_out
for v in base t64 t32 t16 base t64 t32 t16; do
  TWXI_LIB=$PWD/topowx_b200/libtwxi_$v.so TWXI_KED_CFG=$v python tools/time_tile_c5.py 3 3 2>&1 | tail -1
done | tee gpurun_out/covtab.log
for v in t32 t16; do
  TWXI_LIB=$PWD/topowx_b200/libtwxi_$v.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -m gpu -k "krig or full_tile or station" 2>&1 | tail -1
done | tee -a gpurun_out/covtab.log
