#!/bin/bash
# round 2, session e: the step drivers end to end (21 variogram xval, 22 variogram parameters, 23 GWR xval, 24 LOO, 25 gridded
# with rasters on disk and netCDF tiles)
mkdir -p gpurun_out
{
echo "== step21"; timeout 600 python scripts/step21_xval_norm_nnghs.py --nxval 500 2>&1 | tail -14
echo "== step22"; timeout 600 python scripts/step22_build_krig_params.py 2>&1 | tail -2
echo "== step23"; timeout 600 python scripts/step23_xval_anom_nnghs.py 2>&1 | tail -4
echo "== step24"; timeout 600 python scripts/step24_xval_interp.py 2>&1 | tail -3
echo "== step25 nc"; timeout 900 python scripts/step25_interp_tair.py --out /tmp/twx_nc --format nc 2>&1 | tail -2; ls -la /tmp/twx_nc/*/ | head -8
echo "== step25 raw"; timeout 900 python scripts/step25_interp_tair.py --out /tmp/twx_raw --format raw --rasters /tmp/twx_nc/rasters 2>&1 | tail -2
} 2>&1 | tee gpurun_out/steps_r02e.log
