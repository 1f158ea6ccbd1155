#!/bin/bash
# round 2, session h: two problems per CTA (ked2_kernel, TWXI_KED_PAIRS = smallest size class that uses it)
mkdir -p gpurun_out
TWXI_KED_PAIRS=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -4
for pp in 0 1 6 9 10 12; do
  echo "== TWXI_KED_PAIRS=$pp"; TWXI_KED_CFG=pairs$pp TWXI_KED_PAIRS=$pp timeout 300 python tools/time_tile.py 3 2>&1 | tail -1
done 2>&1 | tee gpurun_out/time_r02h.log
