#!/bin/bash
# ncu --set full of one launch of the right-looking kriging kernel: $1 = launches to skip (order: NB = 12, 11, 10, ...)
mkdir -p gpurun_out
TWXI_KED_RL=${2:-1} timeout 600 ncu --set full --clock-control none --import-source on -k regex:ked_rl_kernel --launch-skip ${1:-2} --launch-count 1 -o gpurun_out/rl_full -f python tools/prof_chunk.py 250 250 1 > gpurun_out/ncu_rl_full.log 2>&1; tail -2 gpurun_out/ncu_rl_full.log
