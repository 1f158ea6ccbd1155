#!/bin/bash
# A/B of developer builds of the library on three full-land C5 tiles: tools/gpu_ab.sh <suffix> [<suffix> ...]
# (topowx_b200/libtwxi<suffix>.so; "" = the product library), each timed twice, then the kriging parity tests on each.
mkdir -p gpurun_out; : > gpurun_out/ab.log
for r in 1 2; do for v in "$@"; do
  TWXI_LIB=$PWD/topowx_b200/libtwxi$v.so TWXI_KED_CFG="lib$v" python tools/time_tile_c5.py 3 3 2>&1 | tail -1
done; done | tee -a gpurun_out/ab.log
for v in "$@"; do
  TWXI_LIB=$PWD/topowx_b200/libtwxi$v.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -m gpu -k "krig or full_tile or station or gwr or chunk" 2>&1 | tail -1
done | tee -a gpurun_out/ab.log
