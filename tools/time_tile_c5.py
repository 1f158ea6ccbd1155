"""Stage timing of full-land tiles of the C5 list (10 000 stations) through twxi_interp_chunk (device-resident buffers).
usage: python tools/time_tile_c5.py [reps] [ntiles]   -> mean stage_ms per tile (knn, nngh_params, krig, gwr_daily, fixer_quantise)"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from topowx_b200 import db, _lib, synth
from topowx_b200.context import TwxiContext, interp_chunk
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 3
f, tiler, tiles, nall = bench.c5_tile_list(64)
da = bench.c5_stations(f)
ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in da]
ws = [torch.from_numpy(synth.make_wrk_chk_grid(f, t[1], t[2], 250, 250)).cuda() for t in tiles[:nt]]
lib = _lib.lib
out = None
tot = np.zeros(5); ked = 0.0; cnt = 0
for r in range(reps + 1):
    for w in ws:
        lib.twxi_set_stage_timing(1)
        out = interp_chunk(ctx[0], ctx[1], w, out=out)
        torch.cuda.synchronize()
        s5 = (C.c_float * 5)(); lib.twxi_get_stage_ms(s5)
        kk = C.c_float(); lib.twxi_get_ked_kernel_ms(C.byref(kk))
        if r: tot += np.array(list(s5)); ked += kk.value; cnt += 1
print(os.environ.get("TWXI_KED_CFG", "default"), "stage_ms", np.round(tot / cnt, 3).tolist(), "sum", round(float(tot.sum() / cnt), 3),
      "ked_kernel only", round(ked / cnt, 3))
