#!/bin/bash
# round 2, session l: launch-variant string re-checked after the instruction diet of the kriging kernel
mkdir -p gpurun_out
for cfg in 000000002334445555555 000000002233445555555 000000002223445555555 000000002233345555555; do
  TWXI_KED_CFG=$cfg TWXI_KED_VAR=$cfg timeout 200 python tools/time_tile_c5.py 2 3 2>&1 | tail -1
done | tee -a gpurun_out/kedvar_c5_r02l.log
