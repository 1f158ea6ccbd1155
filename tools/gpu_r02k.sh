#!/bin/bash
# round 2, session k: FEWER resident CTAs with a looser register cap (variants 6 = <3,5,128>, 7 = <3,4,128>, 8 = <5,3,192>,
# 9 = <7,2,255>): does the compiler's rematerialisation under the 80-register cap cost more than a resident CTA is worth?
mkdir -p gpurun_out
for cfg in 000000002334445555555 000000002664445555555 000000002774445555555 000000002338885555555 000000002334449999999 000000006334445555555 000000002364445555555 000000002334845555555; do
  TWXI_KED_CFG=$cfg TWXI_KED_VAR=$cfg timeout 300 python tools/time_tile_c5.py 2 3 2>&1 | tail -1
done | tee gpurun_out/kedvar_c5_r02k.log
