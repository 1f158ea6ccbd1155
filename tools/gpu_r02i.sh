#!/bin/bash
# round 2, session i: resident CTAs per SM of gwr_kernel (launch bounds 3 / 4 / 5 / 6 x 256 threads), C5 tile stage times
for lib in libtwxi_g3.so libtwxi.so libtwxi_g5.so libtwxi_g6.so; do
  echo "== $lib"; TWXI_LIB=topowx_b200/$lib timeout 600 python - <<'PY'
import sys, os, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from topowx_b200 import db, _lib, synth
from topowx_b200.context import TwxiContext, interp_chunk
f, tiler, tiles, nall = bench.c5_tile_list(4)
da = bench.c5_stations(f)
ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in da]
w = torch.from_numpy(synth.make_wrk_chk_grid(f, tiles[0][1], tiles[0][2], 250, 250)).cuda()
lib = _lib.lib; out = None; tot = np.zeros(5)
for i in range(4):
    lib.twxi_set_stage_timing(1)
    out = interp_chunk(ctx[0], ctx[1], w, out=out); torch.cuda.synchronize()
    s5 = (C.c_float * 5)(); lib.twxi_get_stage_ms(s5)
    if i: tot += np.array(list(s5))
print("stage_ms", np.round(tot / 3, 3).tolist())
PY
done 2>&1 | grep "==\|stage_ms" | tee gpurun_out/gwr_minb_r02i.log
