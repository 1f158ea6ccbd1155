#!/bin/bash
# warp-role rotation (TWXI_KED_ROT) x kernel (TWXI_KED_RL): stage timing of the benchmark tile
for rot in 0 1; do for rl in 0 1; do
  echo "rot=$rot rl=$rl"; TWXI_KED_ROT=$rot TWXI_KED_RL=$rl timeout 200 python tools/time_tile.py 3 2>&1 | tail -1
done; done
TWXI_KED_ROT=1 TWXI_KED_RL=0 timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
TWXI_KED_ROT=1 TWXI_KED_RL=1 timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
