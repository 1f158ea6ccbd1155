#!/bin/bash
# KED v3.1: parity tests then worker-count sweep on the benchmark tile
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for cfg in "13,13" "99,99" "1,99" "1,1" "9,13" "9,99" "11,14"; do
  TWXI_KED_CFG=$cfg timeout 300 python tools/time_tile.py 2 2>&1 | tail -1
done
