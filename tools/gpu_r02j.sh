#!/bin/bash
# round 2, session j: launch-variant choice per size class on the C5 tiles (the round-1 table was tuned on the configs[1] tile)
mkdir -p gpurun_out
for cfg in 000000002334443355555 000000002334445555555 000000002334444455555 000000002335555555555 000000002344443355555 000000002334455555555 000000002444445555555 000000003334445555555; do
  TWXI_KED_CFG=$cfg TWXI_KED_VAR=$cfg timeout 300 python tools/time_tile_c5.py 2 3 2>&1 | tail -1
done | tee gpurun_out/kedvar_c5_r02j.log
