#!/bin/bash
# occupancy sweep of the one-warp KED kernel: krig stage time of the benchmark tile vs resident CTAs per SM
for o in 2 4 6 8 12; do
  echo -n "maxocc $o: "; TWXI_KW_MAXOCC=$o TWXI_KED_VAR=666666666666777788888 timeout 300 python tools/time_tile.py 2 2>&1 | tail -1
done
