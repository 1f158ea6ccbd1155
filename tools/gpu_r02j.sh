#!/bin/bash
# round 2, session j2: more resident CTAs / warps for the heavy size classes NB 12-14 (variants 6, 7)
mkdir -p gpurun_out
for cfg in 000000002334445555555 000000002336445555555 000000002336775555555 000000002334775555555 000000002366445555555 000000002337445555555; do
  TWXI_KED_CFG=$cfg TWXI_KED_VAR=$cfg timeout 300 python tools/time_tile_c5.py 2 3 2>&1 | tail -1
done | tee gpurun_out/kedvar_c5_r02j2.log
