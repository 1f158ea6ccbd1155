#!/bin/bash
# ncu --set full of the one-warp KED kernel, size classes NB = 10 and NB = 14 (launch order is NB = 19..1)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ked_warp_kernel' --launch-skip 5 --launch-count 5 -o gpurun_out/kw_j -f python tools/prof_chunk.py 100 100 1 > gpurun_out/ncu_kw_j.log 2>&1; tail -2 gpurun_out/ncu_kw_j.log
