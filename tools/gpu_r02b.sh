#!/bin/bash
# round 2, session b: which ingredient of the prologue-free kriging kernel costs time (benchmark tile, stage ms)
for lib in libtwxi_base.so libtwxi.so libtwxi_e1.so libtwxi_e2.so; do
  echo "== $lib"; TWXI_LIB=topowx_b200/$lib timeout 300 python tools/time_tile.py 3 2>&1 | tail -1
done
for cfg in 000000003334443355555 000000033334443355555 111111113334443355555 000000002334444455555 333333333334443355555; do
  echo "== libtwxi.so KED_VAR=$cfg"; TWXI_KED_CFG=$cfg TWXI_KED_VAR=$cfg timeout 300 python tools/time_tile.py 3 2>&1 | tail -1
done
